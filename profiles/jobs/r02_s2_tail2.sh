cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bfc.py -x -q -k "paper or arxiv or squirrel_shape or contiguous" 2>&1 | tail -4
DCR_LIB_PATH=$PWD/build/libdcr_trace.so timeout 300 python profiles/range_tail_probe.py 2>&1 | tail -20
DCR_LIB_PATH=$PWD/build/libdcr_trace.so PROBE_WORLD=1 timeout 300 python profiles/range_tail_probe.py 2>&1 | tail -3
DCR_LIB_PATH=$PWD/build/libdcr_onlygroup.so PROBE_WORLD=1 timeout 300 python profiles/range_tail_probe.py 2>&1 | tail -2
timeout 300 python profiles/range_scaling_probe.py 2>&1 | tail -5
