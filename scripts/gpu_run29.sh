set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_gpu.log | head -40 | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1r.json 2> gpurun_out/bench_r1r.err; tail -3 gpurun_out/bench_r1r.err | cut -c1-400; cut -c1-330 gpurun_out/bench_r1r.json
timeout 600 python benchmarks/forward_vs_stock.py > gpurun_out/forward_vs_stock_r1r.json 2> gpurun_out/fvs.err; tail -3 gpurun_out/fvs.err | cut -c1-300; cat gpurun_out/forward_vs_stock_r1r.json | cut -c1-1200
MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/kernels_r1r.txt 2> gpurun_out/kernels_r1r.err; head -30 gpurun_out/kernels_r1r.txt | cut -c1-200; tail -3 gpurun_out/kernels_r1r.err
