set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed" gpurun_out/pytest_gpu.log | head -20 | cut -c1-300
MPF_PROBE_KERNELS=bf16x3 timeout 200 python benchmarks/gemm_debug_probe.py > gpurun_out/gemm_debug_probe3.jsonl 2> gpurun_out/gemm_debug_probe3.err; cat gpurun_out/gemm_debug_probe3.jsonl; tail -3 gpurun_out/gemm_debug_probe3.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; tail -2 gpurun_out/bench_r1k.err | cut -c1-300; cut -c1-400 gpurun_out/bench_r1k.json
timeout 400 python benchmarks/forward_vs_stock.py > gpurun_out/forward_vs_stock_r1k.json 2> gpurun_out/fvs.err; cat gpurun_out/forward_vs_stock_r1k.json | cut -c1-1200
timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/torch_profile_r1k.txt 2>&1
