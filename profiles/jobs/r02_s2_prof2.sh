cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-sdrf --no-dense --no-clocks --no-cuda-flavour"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'paper_(group|light)' -s 6 -c 2 -o gpurun_out/r02b_prof $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/ | grep r02b
