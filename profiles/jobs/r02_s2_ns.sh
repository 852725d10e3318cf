cd $GRAFT_REPO_ROOT
for v in nswarp base; do DCR_LIB_PATH=$PWD/build/libdcr_$v.so timeout 120 python profiles/plan_probe.py 2>&1 | tail -1; done
