cd $GRAFT_REPO_ROOT
timeout 900 python tools/dev/gm_sweep.py 2>&1 | tail -8
