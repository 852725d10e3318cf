# parity of the paper-flavour kernels in both membership modes, then the default bench without CPU/SDRF legs
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_bfc.py -x -q -k "paper or arxiv or squirrel or shard" 2>&1 | tail -15 > gpurun_out/r02_test.log
cat gpurun_out/r02_test.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-sdrf > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench.json'))
print(d['ms_per_step'], d['value'], d['config']['phase_ms_rank0'], d['e2e']['ms_per_step'], d['config']['parity_spot_check_vs_c_oracle'])
PY
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-sdrf --no-clocks"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'paper_' -s 9 -c 3 --csv --log-file gpurun_out/r02_launches.csv $CMD > /dev/null 2>&1
grep -v "^==" gpurun_out/r02_launches.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:40], r['Metric Name'], r['Metric Value'])
"
