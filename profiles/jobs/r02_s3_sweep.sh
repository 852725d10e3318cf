cd $GRAFT_REPO_ROOT
for v in base grab2 grab8 run32 run128 cb20 cb28; do
echo == $v; DCR_LIB_PATH=$PWD/build/libdcr_$v.so PROBE_WORLDS=1,8 timeout 300 python profiles/range_scaling_probe.py 2>&1 | tail -2 | cut -c1-100
done
