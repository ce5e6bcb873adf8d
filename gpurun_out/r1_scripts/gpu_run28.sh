#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fast_paths or graph_fusion or module_f or gluon_generic" > gpurun_out/test28.log 2>&1; echo "tests exit $?"; tail -15 gpurun_out/test28.log | cut -c1-300
timeout 300 python tools/prof_module.py --steps 8 > gpurun_out/prof_module28.log 2>&1; echo "prof exit $?"; head -3 gpurun_out/prof_module28.log
