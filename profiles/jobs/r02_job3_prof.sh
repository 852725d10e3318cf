cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bfc.py -x -q -k "paper or arxiv or squirrel or shard" 2>&1 | tail -5
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-sdrf --no-clocks"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'paper_(group|light)' -s 6 -c 2 -o gpurun_out/r02_v2_prof $CMD > gpurun_out/ncu.log 2>&1
ls -la gpurun_out/
