set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_d_decoder_ops.py tests/test_gpu_f_configs.py -m gpu -q -k "conv3x3 or config or tensor_core" 2>&1 | tail -40 > gpurun_out/pytest_conv.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_conv.log | head -30 | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_gpu.log | head -40 | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1x.json 2> gpurun_out/bench_r1x.err; tail -3 gpurun_out/bench_r1x.err | cut -c1-400; cut -c1-330 gpurun_out/bench_r1x.json
