# Round 2, GPU call 28: second tier of 32-column tiles (MPF_GEMM_SMALL_BN=2) vs the 64-column tier, one box.
set -x
mkdir -p gpurun_out
MPF_GEMM_SMALL_BN=2 timeout 600 python -m pytest tests/test_gpu_c_modules.py tests/test_gpu_d_decoder_ops.py -m gpu -q 2>&1 | tail -2
for sb in 2 1; do
MPF_GEMM_SMALL_BN=$sb timeout 600 python bench.py --steps 30 --warmup 3 --batch 2 --no-cpu-baseline --no-stock --no-parity --no-e2e > gpurun_out/r2D_bench_b2_smallbn$sb.json 2> gpurun_out/r2D_bench_b2_smallbn$sb.err
python - <<PY
import json
for l in open("gpurun_out/r2D_bench_b2_smallbn$sb.json"):
    if l.startswith("{"):
        d = json.loads(l); print("small_bn=$sb", round(d["ms_per_step"], 3), d["clocks"]["sm_mhz"])
PY
done
