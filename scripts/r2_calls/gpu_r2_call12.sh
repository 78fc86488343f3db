# Round 2, GPU call 12: joint evaluation of all prediction heads in the criterion (one pass over the rows of the 20
# loss evaluations), fused BCE + dice kernel, multi-head assignment launch, in-place head-gradient layout.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_h_criterion.py tests/test_gpu_g_matcher.py tests/test_gpu_c_modules.py tests/test_abi.py -m gpu -q -x 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock > gpurun_out/r2l_bench_b16.json 2> gpurun_out/r2l_bench_b16.err; tail -3 gpurun_out/r2l_bench_b16.err; cut -c1-1500 gpurun_out/r2l_bench_b16.json
timeout 600 python bench.py --steps 20 --warmup 3 --batch 2 --no-cpu-baseline --no-stock --no-parity > gpurun_out/r2l_bench_b2.json 2> gpurun_out/r2l_bench_b2.err; tail -3 gpurun_out/r2l_bench_b2.err; cut -c1-700 gpurun_out/r2l_bench_b2.json
MPF_B=2 MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/r2l_kernels_step_b2.txt 2>&1; head -60 gpurun_out/r2l_kernels_step_b2.txt | cut -c1-170
MPF_B=16 MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/r2l_kernels_step_b16.txt 2>&1; head -40 gpurun_out/r2l_kernels_step_b16.txt | cut -c1-170
