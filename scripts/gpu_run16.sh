set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; grep -E "^E |FAILED|passed|failed" gpurun_out/pytest_gpu.log | head | cut -c1-200
MPF_PROBE=gemm MPF_REPS=5 timeout 200 python benchmarks/kernel_probe.py 2>&1 | cut -c1-220
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench10.json 2> gpurun_out/bench10.err; tail -2 gpurun_out/bench10.err | cut -c1-300; cat gpurun_out/bench10.json | cut -c1-900
timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/torch_profile_r1k.txt 2>&1
timeout 400 python benchmarks/forward_vs_stock.py > gpurun_out/forward_vs_stock_r1k.json 2> gpurun_out/fvs.err; cat gpurun_out/forward_vs_stock_r1k.json | cut -c1-900
