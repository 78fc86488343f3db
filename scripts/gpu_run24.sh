set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_gpu.log | head -30 | cut -c1-300
MPF_PROBE_KERNELS=bf16x3 timeout 200 python benchmarks/gemm_debug_probe.py > gpurun_out/gemm_debug_probe4.jsonl 2> gpurun_out/gemm_debug_probe4.err; cat gpurun_out/gemm_debug_probe4.jsonl | cut -c1-400; tail -3 gpurun_out/gemm_debug_probe4.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1n.json 2> gpurun_out/bench_r1n.err; tail -3 gpurun_out/bench_r1n.err | cut -c1-400; cut -c1-330 gpurun_out/bench_r1n.json
MPF_SHAPES=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/torch_profile_shapes_r1n.txt 2>&1
