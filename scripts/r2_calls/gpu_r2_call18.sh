# Round 2, GPU call 18: streaming sampler with 512 threads and two points in flight per thread.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_g_matcher.py -m gpu -q -x 2>&1 | tail -3
MPF_PROFILE=1 MPF_SORT_POINTS=1 MPF_STREAM=1 timeout 300 python benchmarks/matcher_probe.py 2>&1 | grep -E "^#|^\{" | cut -c1-330 | head -6 | tee gpurun_out/r2r_matcher_probe_sort1_stream1.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock --no-parity > gpurun_out/r2r_bench_b16.json 2> gpurun_out/r2r_bench_b16.err; cut -c1-330 gpurun_out/r2r_bench_b16.json
