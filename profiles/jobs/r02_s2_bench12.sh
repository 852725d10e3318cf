cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_s4_bench.json 2> gpurun_out/r02_s4_bench.err || tail -20 gpurun_out/r02_s4_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s4_bench.json'))
print("ms/step", round(d['ms_per_step'],4), d['phase_ms_rank0'], "e2e ms", round(d['e2e']['ms_per_step'],3))
print("cuda_flavour", json.dumps(d.get('cuda_flavour'))[:700])
s=d.get('sdrf',{})
print("sdrf", {k:s.get(k) for k in ('iters_per_s','e2e_iters_per_s','speedup_vs_cpu','prefix_matches_cpu')})
print("classical", json.dumps(s.get('classical'))[:1500])
print("directed", json.dumps(s.get('directed'))[:800])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 2 --steps 20 --warmup 3 --no-sdrf > gpurun_out/r02_s4_bench_2gpu.json 2> gpurun_out/r02_s4_bench_2gpu.err || tail -20 gpurun_out/r02_s4_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s4_bench_2gpu.json'))
print("2gpu ms/step", round(d['ms_per_step'],4), d['phase_ms_rank0'], "e2e ms", round(d['e2e']['ms_per_step'],3), d['per_rank'])
print("cuda_flavour", json.dumps(d.get('cuda_flavour'))[:500])
PY
