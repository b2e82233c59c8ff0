mkdir -p gpurun_out
SANITIZE_TIMEOUT=45 tools/sanitize.sh racecheck "cast_transpose or test_gemm_bias_relu or softmax_xent or im2col" 2>&1 | tail -6
SANITIZE_TIMEOUT=45 tools/sanitize.sh memcheck "push_matches or pull_copies or maxpool or gemm_pair_kernel_epilogues or gemm_fused" 2>&1 | tail -6
