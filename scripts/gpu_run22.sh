set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_b2_gemm_bf16x3.py tests/test_gpu_d_decoder_ops.py -q 2>&1 | tail -40 > gpurun_out/pytest_tn.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_tn.log | head -30 | cut -c1-250
MPF_PROBE=gemm MPF_REPS=5 timeout 200 python benchmarks/kernel_probe.py 2>&1 | cut -c1-330
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1l.json 2> gpurun_out/bench_r1l.err; tail -2 gpurun_out/bench_r1l.err | cut -c1-300; cut -c1-400 gpurun_out/bench_r1l.json
