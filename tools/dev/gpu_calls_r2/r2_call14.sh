cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "write_gradient_semantics or gluon_fused or module_fused or fused_allreduce" 2>&1 | tail -6
