# Round 2, GPU call 27: narrow N tiles for K-major GEMMs with few output tiles, A/B on one box.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_b2_gemm_bf16x3.py tests/test_gpu_c_modules.py tests/test_gpu_d_decoder_ops.py -m gpu -q 2>&1 | tail -3
for sb in 1 0; do
MPF_GEMM_SMALL_BN=$sb timeout 600 python bench.py --steps 20 --warmup 3 --batch 2 --no-cpu-baseline --no-stock --no-parity --no-e2e > gpurun_out/r2C_bench_b2_smallbn$sb.json 2> gpurun_out/r2C_bench_b2_smallbn$sb.err
MPF_GEMM_SMALL_BN=$sb timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock --no-parity --no-e2e > gpurun_out/r2C_bench_b16_smallbn$sb.json 2> gpurun_out/r2C_bench_b16_smallbn$sb.err
python - <<PY
import json
for f in ("gpurun_out/r2C_bench_b2_smallbn$sb.json", "gpurun_out/r2C_bench_b16_smallbn$sb.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print("small_bn=$sb", f, round(d["ms_per_step"], 3), d["clocks"]["sm_mhz"])
PY
done
