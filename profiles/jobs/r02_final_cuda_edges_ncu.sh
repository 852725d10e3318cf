cd $GRAFT_REPO_ROOT
timeout 200 python profiles/cuda_flavour_probe.py 2>&1 | tail -2
ncu --set full --clock-control none -k regex:'edges_(support|closing)' -s 8 -c 4 -o gpurun_out/r02_cuda_edges python profiles/cuda_flavour_probe.py > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/r02_cuda_edges.ncu-rep
