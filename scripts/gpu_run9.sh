set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
MPF_SHAPES=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/torch_profile_shapes_r1f.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -2 gpurun_out/bench5.err; cat gpurun_out/bench5.json | cut -c1-300
timeout 600 python benchmarks/forward_vs_stock.py > gpurun_out/forward_vs_stock_r1f.json 2> gpurun_out/fvs.err; cat gpurun_out/forward_vs_stock_r1f.json
