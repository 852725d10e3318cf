cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_sdrf.py tests/test_gpu_directed.py tests/test_gpu_classical.py tests/test_gpu_ref_kernels.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -4
DCR_LIB_PATH=$PWD/build/libdcr_sdrfprof.so timeout 300 python profiles/sdrf_phase_driver.py 2>&1 | head -10
