set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_b_gemm.py -x -q 2>&1 | tail -30 > gpurun_out/pytest_gemm.log; tail -15 gpurun_out/pytest_gemm.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
