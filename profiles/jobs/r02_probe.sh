cd $GRAFT_REPO_ROOT
for lib in build/libdcr_*.so; do
echo "== $lib"
DCR_LIB_PATH=$PWD/$lib PROBE_WORLDS=1,4,8 timeout 300 python profiles/shard_scaling_probe.py 2>&1 | tail -3
done
