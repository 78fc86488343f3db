# Round 2, GPU call 24: static gradients + one concatenation in the captured step; split-K for the long reductions with few
# output tiles (decoder FFN second product).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_b2_gemm_bf16x3.py tests/test_gpu_b_gemm.py tests/test_gpu_c_modules.py tests/test_gpu_d_decoder_ops.py tests/test_gpu_e_graph.py tests/test_gpu_f_configs.py -m gpu -q 2>&1 | tail -3
MPF_B=2 MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/r2z_kernels_step_b2.txt 2>&1; head -4 gpurun_out/r2z_kernels_step_b2.txt | cut -c1-150
timeout 600 python bench.py --steps 20 --warmup 3 --batch 2 --no-cpu-baseline --no-stock --no-parity > gpurun_out/r2z_bench_b2.json 2> gpurun_out/r2z_bench_b2.err; tail -1 gpurun_out/r2z_bench_b2.err | cut -c1-200; cut -c1-330 gpurun_out/r2z_bench_b2.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock --no-parity > gpurun_out/r2z_bench_b16.json 2> gpurun_out/r2z_bench_b16.err; cut -c1-330 gpurun_out/r2z_bench_b16.json
