set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_f_configs.py -m gpu -q -k tensor_core 2>&1 | grep -E "^E  .*Assertion|passed|failed" | cut -c1-900
MPF_NO_CONV3X3_KERNEL=1 timeout 600 python -m pytest tests/test_gpu_f_configs.py -m gpu -q -k tensor_core 2>&1 | grep -E "^E  .*Assertion|passed|failed" | cut -c1-900
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_gpu.log | head -40 | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1y.json 2> gpurun_out/bench_r1y.err; tail -3 gpurun_out/bench_r1y.err | cut -c1-400; cut -c1-330 gpurun_out/bench_r1y.json
MPF_PROBE=layernorm,fpn MPF_REPS=5 timeout 300 python benchmarks/kernel_probe.py 2>&1 | grep -E "colsum|groupnorm" | cut -c1-200
