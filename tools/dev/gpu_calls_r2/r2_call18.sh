cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/ncu_gemm_2cta_8192_r2_call18 python tools/prof_gemm.py gemm_fwd 8192 8192 8192 > gpurun_out/ncu18_a.log 2>&1
timeout 300 $NCU -s 3 -c 1 -f -o gpurun_out/ncu_cublas_8192_r2_call18 python tools/prof_kernels.py cublas 8192 8192 8192 > gpurun_out/ncu18_b.log 2>&1
timeout 300 $NCU -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/ncu_conv_fwd_14x14x256_r2_call18 python tools/prof_kernels.py conv_fwd 512 14 14 256 256 3 1 > gpurun_out/ncu18_c.log 2>&1
timeout 300 $NCU -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/ncu_conv_fwd_56x56x64_r2_call18 python tools/prof_kernels.py conv_fwd 512 56 56 64 64 3 1 > gpurun_out/ncu18_d.log 2>&1
timeout 300 $NCU -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/ncu_conv_wgrad_28x28x128_r2_call18 python tools/prof_kernels.py conv_wgrad 512 28 28 128 128 3 1 > gpurun_out/ncu18_e.log 2>&1
timeout 300 $NCU -k regex:fwd_kernel -s 2 -c 1 -f -o gpurun_out/ncu_attn_fwd_r2_call18 python tools/prof_kernels.py attn 64 128 12 > gpurun_out/ncu18_f.log 2>&1
timeout 300 $NCU -k regex:bwd_kernel -s 2 -c 1 -f -o gpurun_out/ncu_attn_bwd_r2_call18 python tools/prof_kernels.py attn 64 128 12 > gpurun_out/ncu18_g.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu18_*.log
