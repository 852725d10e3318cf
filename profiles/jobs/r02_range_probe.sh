cd $GRAFT_REPO_ROOT
timeout 600 python profiles/range_scaling_probe.py 2>&1 | tail -6
