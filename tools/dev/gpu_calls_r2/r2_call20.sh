cd $GRAFT_REPO_ROOT
timeout 300 python tools/dev/debug_write_sem.py 2>&1 | grep -v "^  " | tail -20
