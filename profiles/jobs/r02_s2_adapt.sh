cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bfc.py -x -q -k "paper or arxiv or squirrel_shape or contiguous" 2>&1 | tail -3
timeout 300 python profiles/range_scaling_probe.py 2>&1 | tail -4 | cut -c1-170
