cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-sdrf --no-dense --no-cpu > gpurun_out/r02_scale_1.json 2> gpurun_out/r02_scale_1.err || tail -5 gpurun_out/r02_scale_1.err
for N in 2 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 2976$N bench.py --gpus $N --steps 20 --warmup 3 --no-sdrf --no-dense --no-cpu > gpurun_out/r02_scale_$N.json 2> gpurun_out/r02_scale_$N.err || tail -5 gpurun_out/r02_scale_$N.err
done
python - <<'PY'
import json
base=None
for N in (1,2,4,8):
    try:
        d=json.load(open(f'gpurun_out/r02_scale_{N}.json'))
    except Exception as e:
        print(N,"failed",e); continue
    if N==1: base=d['ms_per_step']; be=d['e2e']['ms_per_step']
    cf=d.get('cuda_flavour') or {}
    print(N, "ms/step", round(d['ms_per_step'],4), "x", round(base/d['ms_per_step'],2), d['phase_ms_rank0'], "e2e", round(d['e2e']['ms_per_step'],3), "mean/median", d['step_ms_mean_over_median'], "edge ms per rank", [r['edge_kernels_ms'] for r in d['per_rank']], "| cuda flavour ms", round(cf.get('ms_per_step',0),4), cf.get('all_ranks_bit_identical_to_per_entry_kernels'), "parity", d['parity_spot_check_vs_c_oracle'])
PY
