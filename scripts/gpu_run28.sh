set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_gpu.log | head -40 | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1q.json 2> gpurun_out/bench_r1q.err; tail -3 gpurun_out/bench_r1q.err | cut -c1-400; cut -c1-330 gpurun_out/bench_r1q.json
MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/kernels_r1q.txt 2> gpurun_out/kernels_r1q.err; head -60 gpurun_out/kernels_r1q.txt | cut -c1-200; tail -3 gpurun_out/kernels_r1q.err
