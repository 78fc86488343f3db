# Round 2, GPU call 9: self-attention kernels v2, top-k NaN rule, key-split tolerance: tests + bench B=16 / B=2.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "not gradcheck" 2>&1 | tail -30 > gpurun_out/r2i_pytest_gpu.log; tail -8 gpurun_out/r2i_pytest_gpu.log | cut -c1-400
for b in 16 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline --no-parity --no-stock > gpurun_out/r2i_bench_b$b.json 2> gpurun_out/r2i_bench_b$b.err; tail -2 gpurun_out/r2i_bench_b$b.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_b$b.json')); print('B=$b', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['impl_notes']['cuda_graph'][:40])"
done
MPF_B=2 MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/r2i_kernels_step_b2.txt 2> gpurun_out/r2i_kernels_step_b2.err; grep -E "self_attn|lsap|match_cost|topk_gather|xattn" gpurun_out/r2i_kernels_step_b2.txt | cut -c1-140
