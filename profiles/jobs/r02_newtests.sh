cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bfc.py -x -q -k "contiguous or host_end or sharded or golden_integer or global_table or squirrel_shape_every or strided" 2>&1 | tail -15
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-sdrf --no-clocks > gpurun_out/v.json 2> gpurun_out/v.err || tail -5 gpurun_out/v.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/v.json'))
print("ms/step", round(d['ms_per_step'],4), d['config']['phase_ms_rank0'], "e2e", d['e2e']['ms_per_step'], "parity", d['config']['parity_spot_check_vs_c_oracle'])
PY
