set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_b2_gemm_bf16x3.py -x -q 2>&1 | tail -25 > gpurun_out/pytest_bf16x3.log; grep -E "^E |FAILED|passed|failed|Error" gpurun_out/pytest_bf16x3.log | head -20 | cut -c1-300
timeout 300 python benchmarks/gemm_debug_probe.py > gpurun_out/gemm_debug_probe2.jsonl 2> gpurun_out/gemm_debug_probe2.err; cat gpurun_out/gemm_debug_probe2.jsonl; tail -3 gpurun_out/gemm_debug_probe2.err
timeout 800 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; grep -E "^E |FAILED|passed|failed" gpurun_out/pytest_gpu.log | head | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; tail -2 gpurun_out/bench_r1j.err | cut -c1-300; cut -c1-700 gpurun_out/bench_r1j.json
