cd $GRAFT_REPO_ROOT
echo "== default"; timeout 300 python tools/dev/debug_determinism.py 2>&1 | tail -8
echo "== BN_BULK=0"; MXB200_BN_BULK=0 timeout 300 python tools/dev/debug_determinism.py 2>&1 | tail -8
