# quick iteration job: small-graph parity of the paper-flavour kernels (both modes), bench (arxiv spot check inside), per-kernel counters
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_bfc.py -x -q -k "paper and not arxiv and not full_size" 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-sdrf > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err || tail -5 gpurun_out/r02_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench.json'))
print("ms/step", round(d['ms_per_step'],4), "Medges/s", round(d['value']/1e6,1), d['config']['phase_ms_rank0'], "e2e ms", round(d['e2e']['ms_per_step'],3), "parity", d['config']['parity_spot_check_vs_c_oracle'])
PY
if [ "$1" != "nocounters" ]; then
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-sdrf --no-clocks"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'paper_(group|light)' -s 6 -c 2 --csv --log-file gpurun_out/r02_launches.csv $CMD > /dev/null 2>&1
grep -v "^==" gpurun_out/r02_launches.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:30], r['Metric Name'][:40], r['Metric Value'])
"
fi
