set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_d_decoder_ops.py -m gpu -q -k "masked_cross_attention" 2>&1 | tail -30 > gpurun_out/pytest_xattn.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_xattn.log | head -30 | cut -c1-400
MPF_PROBE=xattn MPF_REPS=5 timeout 300 python benchmarks/kernel_probe.py 2>&1 | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_gpu.log | head -40 | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -3 gpurun_out/bench_r2d.err | cut -c1-400; cut -c1-330 gpurun_out/bench_r2d.json
