set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; grep -E "^E |FAILED|passed|failed" gpurun_out/pytest_gpu.log | head | cut -c1-200
MPF_PROBE=msda MPF_REPS=5 timeout 200 python benchmarks/kernel_probe.py 2>&1 | cut -c1-140
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench9.json 2> gpurun_out/bench9.err; tail -2 gpurun_out/bench9.err | cut -c1-300; cat gpurun_out/bench9.json | cut -c1-300
timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/torch_profile_r1j.txt 2>&1
