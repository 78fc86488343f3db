# Round 2, GPU call 25: weight operand cache (one split launch per step) A/B on one box; warp-per-cell cost finish kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e_graph.py tests/test_gpu_g_matcher.py tests/test_gpu_c_modules.py tests/test_gpu_d_decoder_ops.py -m gpu -q 2>&1 | tail -4
for mode in "" "--no-weight-cache"; do
tag=cache; [ -n "$mode" ] && tag=nocache
timeout 600 python bench.py --steps 20 --warmup 3 --batch 2 --no-cpu-baseline --no-stock --no-parity --no-e2e $mode > gpurun_out/r2A_bench_b2_$tag.json 2> gpurun_out/r2A_bench_b2_$tag.err; tail -1 gpurun_out/r2A_bench_b2_$tag.err | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock --no-parity --no-e2e $mode > gpurun_out/r2A_bench_b16_$tag.json 2> gpurun_out/r2A_bench_b16_$tag.err
python - <<PY
import json
for f in ("gpurun_out/r2A_bench_b2_$tag.json", "gpurun_out/r2A_bench_b16_$tag.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print("$tag", f, round(d["ms_per_step"], 3), d["gpu_launches"], d["clocks"]["sm_mhz"], d["impl_notes"].get("weight_operand_cache"))
PY
done
