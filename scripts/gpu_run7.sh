set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_b_gemm.py -q 2>&1 | tail -30 > gpurun_out/pytest_gemm.log; grep -E "FAILED|passed|failed" gpurun_out/pytest_gemm.log | cut -c1-200
