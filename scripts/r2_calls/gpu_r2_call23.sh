# Round 2, GPU call 23: model-based split-K count of the short token-reduction GEMMs.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_b2_gemm_bf16x3.py tests/test_gpu_c_modules.py tests/test_gpu_d_decoder_ops.py tests/test_gpu_e_graph.py -m gpu -q -x 2>&1 | tail -3
MPF_B=2 MPF_PROBE=gemm_small timeout 300 python benchmarks/kernel_probe.py 2>&1 | tee gpurun_out/r2w_gemm_small_b2.jsonl | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{') and '_tn' in ln:
        d = json.loads(ln); print('B=2 %-100s %8.2f us' % (d['kernel'][:100], d['ms']*1e3))
"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock --no-parity > gpurun_out/r2w_bench_b16.json 2> gpurun_out/r2w_bench_b16.err; cut -c1-330 gpurun_out/r2w_bench_b16.json
timeout 600 python bench.py --steps 20 --warmup 3 --batch 2 --no-cpu-baseline --no-stock --no-parity > gpurun_out/r2w_bench_b2.json 2> gpurun_out/r2w_bench_b2.err; cut -c1-330 gpurun_out/r2w_bench_b2.json
