set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -2
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-sdrf --no-clocks"
python bench.py --steps 10 --warmup 3 --no-cpu --no-sdrf > gpurun_out/r02_base_bench.json 2> gpurun_out/r02_base_bench.err
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'paper_(group|light)' -s 6 -c 2 -o gpurun_out/r02_base_prof $CMD > gpurun_out/ncu.log 2>&1
ls -la gpurun_out/
ncu -i gpurun_out/r02_base_prof.ncu-rep --page source --csv > gpurun_out/r02_base_source.csv 2>/dev/null
ls -la gpurun_out/
