set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_gpu.log | head -40 | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1v.json 2> gpurun_out/bench_r1v.err; tail -3 gpurun_out/bench_r1v.err | cut -c1-400; cut -c1-330 gpurun_out/bench_r1v.json
MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/kernels_r1v.txt 2> gpurun_out/kernels_r1v.err; head -24 gpurun_out/kernels_r1v.txt | cut -c1-200; grep -E "transpose_split|tf32x3" gpurun_out/kernels_r1v.txt | cut -c1-160
