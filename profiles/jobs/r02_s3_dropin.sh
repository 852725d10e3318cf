cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_classical.py -x -q 2>&1 | tail -3
