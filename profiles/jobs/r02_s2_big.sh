cd $GRAFT_REPO_ROOT
for v in base2 big1k big512 big256; do
echo == $v; DCR_LIB_PATH=$PWD/build/libdcr_$v.so PROBE_WORLDS=1,8 timeout 300 python profiles/range_scaling_probe.py 2>&1 | tail -2 | cut -c1-150
done
DCR_LIB_PATH=$PWD/build/libdcr_base2.so timeout 600 python -m pytest tests/test_gpu_bfc.py -x -q -k "arxiv or contiguous or squirrel_shape" 2>&1 | tail -2
