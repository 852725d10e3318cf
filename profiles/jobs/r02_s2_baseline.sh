cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_s2_tests.log; cat gpurun_out/r02_s2_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_s2_bench.json 2> gpurun_out/r02_s2_bench.err || tail -20 gpurun_out/r02_s2_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-sdrf --no-dense --no-clocks"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_s2_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'paper_(group|light)' -s 6 -c 2 -o gpurun_out/r02_s2_prof $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/ | tail -8
