set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; grep -E "^E |FAILED|passed|failed" gpurun_out/pytest_gpu.log | head | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err; tail -2 gpurun_out/bench7.err | cut -c1-300; cat gpurun_out/bench7.json | cut -c1-300
timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/torch_profile_r1h.txt 2>&1
timeout 600 python benchmarks/forward_vs_stock.py > gpurun_out/forward_vs_stock_r1h.json 2> gpurun_out/fvs.err; cat gpurun_out/forward_vs_stock_r1h.json | cut -c1-700
