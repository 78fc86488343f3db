# Round 2, GPU call 17: matcher components (gather / sorted gather / sorted + streamed with bulk-copied bands).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_g_matcher.py -m gpu -q -x 2>&1 | tail -3
for cfg in "0 0" "1 0" "1 1"; do set -- $cfg
MPF_PROFILE=1 MPF_SORT_POINTS=$1 MPF_STREAM=$2 timeout 300 python benchmarks/matcher_probe.py 2>&1 | grep -E "^#|^\{" | cut -c1-330 | head -22 | tee gpurun_out/r2q_matcher_probe_sort$1_stream$2.txt
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock --no-parity > gpurun_out/r2q_bench_b16.json 2> gpurun_out/r2q_bench_b16.err; cut -c1-330 gpurun_out/r2q_bench_b16.json
