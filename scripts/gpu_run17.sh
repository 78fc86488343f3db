set -x
MPF_PROBE=gemm MPF_REPS=5 timeout 200 python benchmarks/kernel_probe.py 2>&1 | cut -c1-120
MPF_GEMM_NO_STAGE=1 MPF_PROBE=gemm MPF_REPS=5 timeout 200 python benchmarks/kernel_probe.py 2>&1 | cut -c1-120
MPF_GEMM_NO_STAGE=1 timeout 300 python benchmarks/torch_profile_step.py 2>&1 | grep -E "gemm_tf32x3_kernel|_FFNBackward  |Self CUDA time" | cut -c1-70,180-260
