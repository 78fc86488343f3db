# Round 2, GPU call 7: key-split debug, self-attention kernel, full suite, bench.
set -x
mkdir -p gpurun_out
timeout 300 python benchmarks/debug_xattn_split.py 2>&1 | tail -20 | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q -k "not gradcheck and not key_splits" 2>&1 | tail -30 > gpurun_out/r2g_pytest_gpu.log; tail -12 gpurun_out/r2g_pytest_gpu.log | cut -c1-400
for b in 16 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline --no-parity --no-stock > gpurun_out/r2g_bench_b$b.json 2> gpurun_out/r2g_bench_b$b.err; tail -2 gpurun_out/r2g_bench_b$b.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_b$b.json')); print('B=$b', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['impl_notes']['cuda_graph'][:40]); print(d['impl_notes']['routes'])"
done
