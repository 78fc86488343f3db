set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_b_gemm.py -q -x 2>&1 | tail -30 > gpurun_out/pytest_gemm.log; tail -12 gpurun_out/pytest_gemm.log | cut -c1-250
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
MPF_PROBE=msda timeout 200 python benchmarks/kernel_probe.py 2>&1 | cut -c1-200
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -2 gpurun_out/bench4.err; cat gpurun_out/bench4.json | cut -c1-300
timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/torch_profile_r1e.txt 2>&1
