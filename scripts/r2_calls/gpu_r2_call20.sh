# Round 2, GPU call 20: full GPU suite + default bench line (all legs) on the current tree.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2t_pytest_gpu.log; cat gpurun_out/r2t_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.err; tail -2 gpurun_out/r2t_bench_n1.err | cut -c1-300; cut -c1-600 gpurun_out/r2t_bench_n1.json
