set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_d_decoder_ops.py tests/test_gpu_c_modules.py -m gpu -q -k "masked_cross_attention or decoder or training_step" 2>&1 | tail -30 > gpurun_out/pytest_xattn.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_xattn.log | head -30 | cut -c1-400
MPF_PROBE=xattn_bwd MPF_REPS=5 timeout 300 python benchmarks/kernel_probe.py 2>&1 | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; tail -3 gpurun_out/bench_r2f.err | cut -c1-400; cut -c1-330 gpurun_out/bench_r2f.json
MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/kernels_r2f.txt 2> gpurun_out/kernels_r2f.err; grep -E "xattn|total self" gpurun_out/kernels_r2f.txt | cut -c1-160
