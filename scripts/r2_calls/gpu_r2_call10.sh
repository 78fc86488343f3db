# Round 2, GPU call 10: CTA-pair (cta_group::2) GEMM: parity vs single CTA, then bench with / without.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_b2_gemm_bf16x3.py -m gpu -q -x -k "cta_pairs" 2>&1 | tail -25 | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q -x -k "not gradcheck" 2>&1 | tail -8 | cut -c1-300
for pm in 0 1; do
MPF_GEMM_PAIR=$pm timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-stock > gpurun_out/r2j_bench_pair$pm.json 2> gpurun_out/r2j_bench_pair$pm.err; tail -2 gpurun_out/r2j_bench_pair$pm.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_pair$pm.json')); r=d['roofline_gemm']; print('PAIR=$pm', d['value'], d['ms_per_step'], d['e2e']['value']); print({k: r[k] for k in ('tensor_TFLOPs_algorithmic','tensor_issue_frac','hbm_frac','frac_of_combined_roof','avg_launch_ms','share_of_step','launches_timed')})"
done
MPF_PROBE=gemm,conv timeout 300 python benchmarks/kernel_probe.py 2>&1 | tail -12 | cut -c1-400
MPF_GEMM_PAIR=0 MPF_PROBE=gemm,conv timeout 300 python benchmarks/kernel_probe.py 2>&1 | tail -12 | cut -c1-400
