cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_bfc.py -x -q -k "host_end or two_gpus" 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-sdrf --no-dense --no-cuda-flavour > gpurun_out/r02_e2e1.json 2>gpurun_out/r02_e2e1.err || tail -5 gpurun_out/r02_e2e1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29755 bench.py --gpus 2 --steps 10 --warmup 3 --no-sdrf --no-cuda-flavour > gpurun_out/r02_e2e2.json 2>gpurun_out/r02_e2e2.err || tail -5 gpurun_out/r02_e2e2.err
python -c "
import json
for f in ('r02_e2e1','r02_e2e2'):
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'], d['e2e_parity_spot_check_vs_c_oracle'])"
