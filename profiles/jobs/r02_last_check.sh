cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_ref_kernels.py tests/test_gpu_sdrf.py -x -q -k "not full_length and not squirrel_shape_sequence" 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
