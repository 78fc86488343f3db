# Round 2, GPU call 21: radix select with warp-aggregated histogram atomics; per-kernel tables of the current step.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_h_criterion.py -m gpu -q -x 2>&1 | tail -3
MPF_B=16 MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/r2u_kernels_step_b16.txt 2>&1; grep -E "topk|lsap|match_cost|total self" gpurun_out/r2u_kernels_step_b16.txt | cut -c1-150
MPF_B=2 MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/r2u_kernels_step_b2.txt 2>&1; grep -E "topk|lsap|match_cost|total self" gpurun_out/r2u_kernels_step_b2.txt | cut -c1-150
