cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_classical.py tests/test_gpu_sdrf.py tests/test_gpu_directed.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -15
