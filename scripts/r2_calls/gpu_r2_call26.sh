# Round 2, GPU call 26: bias gradients from the weight-gradient GEMM (column sums in the TN kernel's converter warps).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_b2_gemm_bf16x3.py tests/test_gpu_b_gemm.py tests/test_gpu_c_modules.py tests/test_gpu_d_decoder_ops.py tests/test_gpu_e_graph.py tests/test_gpu_f_configs.py tests/test_gpu_j_full_config.py -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 --batch 2 --no-cpu-baseline --no-stock --no-parity --no-e2e > gpurun_out/r2B_bench_b2.json 2> gpurun_out/r2B_bench_b2.err; tail -1 gpurun_out/r2B_bench_b2.err | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock --no-parity --no-e2e > gpurun_out/r2B_bench_b16.json 2> gpurun_out/r2B_bench_b16.err
python - <<PY
import json
for f in ("gpurun_out/r2B_bench_b2.json", "gpurun_out/r2B_bench_b16.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print(f, round(d["ms_per_step"], 3), d["gpu_launches"], d["clocks"]["sm_mhz"])
PY
