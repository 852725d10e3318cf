cd $GRAFT_REPO_ROOT
PROBE_WORLDS=2,4,8 timeout 300 python profiles/range_scaling_probe.py 2>&1 | tail -3
