cd $GRAFT_REPO_ROOT
DCR_LIB_PATH=$PWD/build/libdcr_sdrfprof.so timeout 300 python profiles/sdrf_phase_driver.py 2>&1 | head -11
