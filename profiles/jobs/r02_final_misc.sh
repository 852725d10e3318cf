cd $GRAFT_REPO_ROOT
DCR_PAPER_MODE=hashed timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-sdrf --no-dense --no-cuda-flavour > gpurun_out/r02_final_hashed.json 2> gpurun_out/r02_final_hashed.err || tail -5 gpurun_out/r02_final_hashed.err
python -c "
import json; d=json.load(open('gpurun_out/r02_final_hashed.json')); print('hashed mode ms/step', d['ms_per_step'], d['parity_spot_check_vs_c_oracle'])"
timeout 300 python profiles/range_scaling_probe.py > gpurun_out/r02_final_range_probe.txt 2>&1; cat gpurun_out/r02_final_range_probe.txt | cut -c1-220
timeout 300 python profiles/cuda_flavour_probe.py 2>&1 | tail -2
