cd $GRAFT_REPO_ROOT
python profiles/sanitizer_driver.py 2>&1 | tail -2
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python profiles/sanitizer_driver.py > gpurun_out/r02_memcheck.log 2>&1; tail -6 gpurun_out/r02_memcheck.log
