cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err || tail -20 gpurun_out/r02_final_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_ref.json 2> gpurun_out/r02_final_ref.err || tail -5 gpurun_out/r02_final_ref.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-sdrf --no-dense --no-clocks --no-cuda-flavour"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'paper_(group|light)' -s 6 -c 2 -o gpurun_out/r02_final_prof $CMD > gpurun_out/ncu2.log 2>&1
CMD2="python bench.py --workload squirrel-dense --steps 2 --warmup 2"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none -k regex:'tc_support' -s 2 -c 2 -o gpurun_out/r02_final_tc $CMD2 > gpurun_out/ncu3.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_final_bench.json'))
print("ms/step", round(d['ms_per_step'],4), "Medges/s", round(d['value']/1e6,1), d['phase_ms_rank0'], "e2e ms", round(d['e2e']['ms_per_step'],3), "parity", d['parity_spot_check_vs_c_oracle'], d['e2e_parity_spot_check_vs_c_oracle'], "clocks", d['clocks']['sm_mhz'], d['clocks']['reasons'])
print("cpu", d.get('cpu_baseline',{}).get('value'), "cuda_flavour ms", d['cuda_flavour']['ms_per_step'], "dense", d['dense']['ms_per_step'], d['dense']['roofline']['frac'])
s=d['sdrf']; print("sdrf", s['iters_per_s'], s['speedup_vs_cpu'], s.get('squirrel',{}).get('iters_per_s'), s['directed']['iters_per_s'], {k:v['iters_per_s'] for k,v in s['classical'].items()})
r=json.load(open('gpurun_out/r02_final_ref.json')); print("ref", r['value'], r['config']==d['config'])
PY
