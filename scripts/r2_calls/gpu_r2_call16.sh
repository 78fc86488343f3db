# Round 2, GPU call 16: matcher with streamed prediction samples; native GEMM for the two small attention-backward
# products; lsap with the cost matrix in shared memory.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_g_matcher.py tests/test_gpu_h_criterion.py tests/test_gpu_d_decoder_ops.py tests/test_gpu_c_modules.py -m gpu -q -x 2>&1 | tail -8
MPF_SORT_POINTS=1 timeout 300 python benchmarks/matcher_probe.py 2>&1 | tail -2 | cut -c1-500
timeout 300 python benchmarks/matcher_probe.py 2>&1 | tail -2 | cut -c1-500
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock --no-parity > gpurun_out/r2p_bench_b16.json 2> gpurun_out/r2p_bench_b16.err; tail -2 gpurun_out/r2p_bench_b16.err | cut -c1-300; cut -c1-400 gpurun_out/r2p_bench_b16.json
timeout 600 python bench.py --steps 20 --warmup 3 --batch 2 --no-cpu-baseline --no-stock --no-parity > gpurun_out/r2p_bench_b2.json 2> gpurun_out/r2p_bench_b2.err; cut -c1-400 gpurun_out/r2p_bench_b2.json
