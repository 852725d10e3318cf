cd $GRAFT_REPO_ROOT
DCR_LIB_PATH=$PWD/build/libdcr_trace.so timeout 300 python profiles/range_tail_probe.py 2>&1 | tail -20
for v in trace coop8k; do
echo == $v; DCR_LIB_PATH=$PWD/build/libdcr_$v.so PROBE_WORLDS=1,8 timeout 300 python profiles/range_scaling_probe.py 2>&1 | tail -2
done
