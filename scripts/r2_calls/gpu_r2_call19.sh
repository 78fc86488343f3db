# Round 2, GPU call 19: cost kernel whose warps skip the point reduction when their four targets do not exist.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_g_matcher.py tests/test_gpu_h_criterion.py -m gpu -q -x 2>&1 | tail -3
MPF_PROFILE=1 MPF_SORT_POINTS=1 MPF_STREAM=0 timeout 300 python benchmarks/matcher_probe.py 2>&1 | grep -E "^#|^\{" | cut -c1-330 | head -4 | tee gpurun_out/r2s_matcher_probe_sort1.txt
MPF_PROFILE=1 MPF_B=2 MPF_SORT_POINTS=0 timeout 300 python benchmarks/matcher_probe.py 2>&1 | grep -E "^#|^\{" | cut -c1-330 | head -4 | tee gpurun_out/r2s_matcher_probe_b2.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock --no-parity > gpurun_out/r2s_bench_b16.json 2> gpurun_out/r2s_bench_b16.err; cut -c1-330 gpurun_out/r2s_bench_b16.json
timeout 600 python bench.py --steps 20 --warmup 3 --batch 2 --no-cpu-baseline --no-stock --no-parity > gpurun_out/r2s_bench_b2.json 2> gpurun_out/r2s_bench_b2.err; cut -c1-330 gpurun_out/r2s_bench_b2.json
