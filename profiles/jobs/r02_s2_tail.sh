cd $GRAFT_REPO_ROOT
for v in trace onlygroup onlylight; do
DCR_LIB_PATH=$PWD/build/libdcr_$v.so timeout 300 python profiles/range_tail_probe.py 2>&1 | tail -20
done
