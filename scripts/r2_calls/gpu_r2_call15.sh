# Round 2, GPU call 15: TN GEMM with one ring of in-place converted slots (four 32-token blocks in flight at BN = 256).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_b2_gemm_bf16x3.py tests/test_gpu_b_gemm.py tests/test_gpu_c_modules.py tests/test_gpu_f_configs.py -m gpu -q -x 2>&1 | tail -5
for ring in 0 2 3; do
$( [ "$ring" != 0 ] && echo env MPF_GEMM_TN_RING=$ring ) env MPF_PROBE=gemm timeout 300 python benchmarks/kernel_probe.py 2>&1 | tee gpurun_out/r2o_gemm_probe_tn_ring$ring.jsonl | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln)
        if '_tn' in d['kernel'] or 'wgrad' in d['kernel']: print('ring $ring: %-90s %.4f ms' % (d['kernel'][:90], d['ms']))
"
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock --no-parity > gpurun_out/r2o_bench_b16.json 2> gpurun_out/r2o_bench_b16.err; tail -2 gpurun_out/r2o_bench_b16.err; cut -c1-400 gpurun_out/r2o_bench_b16.json
timeout 600 python bench.py --steps 20 --warmup 3 --batch 2 --no-cpu-baseline --no-stock --no-parity > gpurun_out/r2o_bench_b2.json 2> gpurun_out/r2o_bench_b2.err; cut -c1-400 gpurun_out/r2o_bench_b2.json
