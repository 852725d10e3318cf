cd $GRAFT_REPO_ROOT
timeout 600 python profiles/cost_model_fit.py 2>&1 | tail -8
