cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29741 tests/multi_gpu_worker.py 2>&1 | tail -15
for ex in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 20 --warmup 3 --no-sdrf --exchange $ex > gpurun_out/r02_bench_2gpu_$ex.json 2> gpurun_out/r02_bench_2gpu_$ex.err || tail -20 gpurun_out/r02_bench_2gpu_$ex.err
python - $ex <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/r02_bench_2gpu_{sys.argv[1]}.json'))
print(sys.argv[1], d['exchange'], "ms/step", round(d['ms_per_step'],4), "Medges/s", round(d['value']/1e6,1), d['phase_ms_rank0'], "e2e ms", round(d['e2e']['ms_per_step'],3), "parity", d['parity_spot_check_vs_c_oracle'], d['e2e_parity_spot_check_vs_c_oracle'], "mean/median", d['step_ms_mean_over_median'])
print(d['per_rank']); print(d['step_ms'])
PY
done
