cd $GRAFT_REPO_ROOT
DCR_LIB_PATH=$PWD/build/libdcr_trace.so timeout 300 python profiles/paper_cta_timeline.py 2>&1 | tail -24
mv build/libdcr_trace.so build/trace.so.skip
bash profiles/jobs/r02_counters.sh
