cd $GRAFT_REPO_ROOT
timeout 300 python tools/dev/debug_kernel_determinism.py 2>&1 | tail -14
timeout 300 python tools/dev/debug_model_determinism.py 2>&1 | tail -30
