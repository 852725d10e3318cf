cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bfc.py -x -q -k "edge_centric or two_gpus" 2>&1 | tail -15
timeout 300 python profiles/cuda_flavour_probe.py 2>&1 | tail -5
