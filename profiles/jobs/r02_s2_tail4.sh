cd $GRAFT_REPO_ROOT
DCR_LIB_PATH=$PWD/build/libdcr_trace.so PROBE_WORLD=1 timeout 300 python profiles/range_tail_probe.py 2>&1 | tail -4
DCR_LIB_PATH=$PWD/build/libdcr_trace.so PROBE_WORLD=8 timeout 300 python profiles/range_tail_probe.py 2>&1 | grep -A1 "rank 7" | tail -3
