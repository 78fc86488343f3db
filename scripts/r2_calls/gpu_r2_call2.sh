# Round 2, GPU call 2: LSU instruction-shape micro-benchmark (design input for the MSDeformAttn kernels), the default
# bench again now that the criterion captures into the CUDA graph (B=16 and the per-rank B=2 of the 8-GPU split).
set -x
mkdir -p gpurun_out
./benchmarks/micro/lsu_patterns > gpurun_out/r2b_lsu_patterns.jsonl 2>&1; cat gpurun_out/r2b_lsu_patterns.jsonl | cut -c1-260
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-stock > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; tail -3 gpurun_out/r2b_bench_n1.err | cut -c1-300; cut -c1-300 gpurun_out/r2b_bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['impl_notes']['cuda_graph'])"
timeout 400 python bench.py --steps 20 --warmup 3 --batch 2 --no-cpu-baseline --no-parity --no-stock > gpurun_out/r2b_bench_b2.json 2> gpurun_out/r2b_bench_b2.err; tail -3 gpurun_out/r2b_bench_b2.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_b2.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['impl_notes']['cuda_graph'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'match_cost_partial' -c 1 -o gpurun_out/r2b_ncu_match_cost python benchmarks/matcher_probe.py > gpurun_out/r2b_ncu_match_cost.log 2>&1; tail -2 gpurun_out/r2b_ncu_match_cost.log
