cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/r02_default_bench.json 2> gpurun_out/r02_default_bench.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_default_bench.json'))
print("default bench: ms/step", round(d['ms_per_step'],4), "e2e", round(d['e2e']['ms_per_step'],3), "keys", sorted(d.keys()))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29799 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_default_bench_2gpu.json 2> gpurun_out/r02_default_bench_2gpu.err || tail -5 gpurun_out/r02_default_bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02_default_bench_2gpu.json')); print('2gpu default: ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), 'sdrf' in d, 'cuda_flavour' in d)"
