cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/ncu_conv_fwd_56x56x64_r2_call28 python tools/prof_kernels.py conv_fwd 512 56 56 64 64 3 1 > gpurun_out/ncu28_a.log 2>&1
MXB200_CONV_SHIFT=2 timeout 300 $NCU -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/ncu_conv_fwd_56x56x64_shift2_r2_call28 python tools/prof_kernels.py conv_fwd 512 56 56 64 64 3 1 > gpurun_out/ncu28_b.log 2>&1
timeout 300 $NCU -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/ncu_conv_dgrad_56x56x64_r2_call28 python tools/prof_kernels.py conv_dgrad 512 56 56 64 64 3 1 > gpurun_out/ncu28_c.log 2>&1
timeout 300 $NCU -k "regex:nvjet|cutlass|sm100|gemm" -s 3 -c 1 -f -o gpurun_out/ncu_cublas_8192_r2_call28 python tools/prof_kernels.py cublas 8192 8192 8192 > gpurun_out/ncu28_d.log 2>&1
timeout 300 $NCU -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/ncu_gemm_2cta_8192_r2_call28 python tools/prof_gemm.py gemm_fwd 8192 8192 8192 > gpurun_out/ncu28_e.log 2>&1
echo "== microbench conv shift modes after loop v2"
MXB200_CONV_SHIFT=2 timeout 300 python tools/microbench_conv.py --out gpurun_out/microbench_conv_shift2_r2_call28.json 2>&1 | tail -1
echo "== unique test + inference"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "unique" 2>&1 | tail -2
timeout 600 python examples/image-classification/benchmark_score.py --networks resnet50_v1b,resnet152_v1b --batch-sizes 32,128 --dtype bfloat16 2>&1 | tail -6
ls gpurun_out/*call28*.ncu-rep | wc -l
