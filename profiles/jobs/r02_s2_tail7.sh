cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bfc.py -x -q -k "arxiv or contiguous" 2>&1 | tail -3
DCR_LIB_PATH=$PWD/build/libdcr_trace.so PROBE_WORLD=8 timeout 300 python profiles/range_tail_probe.py 2>&1 | grep -A1 "rank [07]" | tail -6
timeout 300 python profiles/range_scaling_probe.py 2>&1 | tail -5
