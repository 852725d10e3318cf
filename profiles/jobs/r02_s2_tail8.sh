cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bfc.py -x -q -k "paper or arxiv or squirrel_shape or contiguous" 2>&1 | tail -3
DCR_LIB_PATH=$PWD/build/libdcr_onlylight.so PROBE_WORLD=8 timeout 300 python profiles/range_tail_probe.py 2>&1 | grep "rank [07]" | cut -c1-60
DCR_LIB_PATH=$PWD/build/libdcr_onlylight.so PROBE_WORLD=1 timeout 300 python profiles/range_tail_probe.py 2>&1 | grep "rank 0" | cut -c1-60
timeout 300 python profiles/range_scaling_probe.py 2>&1 | tail -5
