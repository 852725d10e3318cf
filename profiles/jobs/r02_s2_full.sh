cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-sdrf --no-dense > gpurun_out/r02_s3_bench.json 2> gpurun_out/r02_s3_bench.err || tail -20 gpurun_out/r02_s3_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s3_bench.json'))
print("ms/step", round(d['ms_per_step'],4), "Medges/s", round(d['value']/1e6,1), d['phase_ms_rank0'], "e2e ms", round(d['e2e']['ms_per_step'],3), "parity", d['parity_spot_check_vs_c_oracle'], d['e2e_parity_spot_check_vs_c_oracle'])
print({k:v for k,v in d['roofline'].items() if k in ('frac','frac_issue','issue_slots_ms','traffic','edge_kernels_ms')})
PY
