cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_ref_kernels.py -x -q -k "full_length or squirrel_shape_sequence" --durations=3 2>&1 | tail -12
