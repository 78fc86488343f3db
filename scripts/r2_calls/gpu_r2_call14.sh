# Round 2, GPU call 14: bf16x3 GEMM with the A tile converted in place (48 KiB stages -> four in flight at BN = 256).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_b2_gemm_bf16x3.py tests/test_gpu_b_gemm.py tests/test_gpu_c_modules.py tests/test_gpu_f_configs.py -m gpu -q -x 2>&1 | tail -5
MPF_PROBE=gemm timeout 300 python benchmarks/kernel_probe.py 2>&1 | tee gpurun_out/r2n_gemm_probe_inplace.jsonl | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('%-90s %.4f ms' % (d['kernel'][:90], d['ms']))
"
MPF_GEMM_STAGES=3 MPF_PROBE=gemm timeout 300 python benchmarks/kernel_probe.py 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('3 stages: %-80s %.4f ms' % (d['kernel'][:80], d['ms']))
"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock --no-parity > gpurun_out/r2n_bench_b16.json 2> gpurun_out/r2n_bench_b16.err; tail -2 gpurun_out/r2n_bench_b16.err; cut -c1-400 gpurun_out/r2n_bench_b16.json
