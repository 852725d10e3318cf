cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err || tail -20 gpurun_out/r02_bench_full.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_full.json'))
print("ms/step", round(d['ms_per_step'],4), "Medges/s", round(d['value']/1e6,1), d['phase_ms_rank0'], "e2e ms", round(d['e2e']['ms_per_step'],3), d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'], "parity", d['parity_spot_check_vs_c_oracle'], d['e2e_parity_spot_check_vs_c_oracle'])
print("clocks", d['clocks'] and {k:d['clocks'][k] for k in ('sm_mhz','sm_max_mhz','reasons')})
print("cpu", d.get('cpu_baseline',{}).get('value'), "py", d.get('python_speed_baseline',{}).get('value'))
print("dense", json.dumps(d.get('dense'))[:600])
s=d.get('sdrf',{})
print("sdrf", {k:s.get(k) for k in ('iters_per_s','e2e_iters_per_s','speedup_vs_cpu','prefix_matches_cpu','speedup_vs_reference_numba')}, s.get('cpu_baseline'), s.get('squirrel'))
PY
