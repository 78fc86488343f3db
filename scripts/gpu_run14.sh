set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_a_msda.py -q -x 2>&1 | tail -25 > gpurun_out/pytest_a.log; grep -E "^E |FAILED|passed|failed" gpurun_out/pytest_a.log | head | cut -c1-250
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; grep -E "^E |FAILED|passed|failed" gpurun_out/pytest_gpu.log | head | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; tail -2 gpurun_out/bench8.err | cut -c1-300; cat gpurun_out/bench8.json | cut -c1-300
timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/torch_profile_r1i.txt 2>&1
