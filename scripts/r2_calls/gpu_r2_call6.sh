# Round 2, GPU call 6: key-split attention, radix-select top-k, reworked match-cost kernel: parity tests + bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "not gradcheck" 2>&1 | tail -30 > gpurun_out/r2f_pytest_gpu.log; tail -12 gpurun_out/r2f_pytest_gpu.log | cut -c1-400
timeout 120 python benchmarks/matcher_probe.py > gpurun_out/r2f_matcher_probe.json 2> gpurun_out/r2f_matcher_probe.err; cat gpurun_out/r2f_matcher_probe.json | cut -c1-500
MPF_SORT_POINTS=1 timeout 120 python benchmarks/matcher_probe.py > gpurun_out/r2f_matcher_probe_sorted.json 2>> gpurun_out/r2f_matcher_probe.err; cat gpurun_out/r2f_matcher_probe_sorted.json | cut -c1-500
timeout 200 python benchmarks/criterion_probe.py > gpurun_out/r2f_criterion_probe.json 2> gpurun_out/r2f_criterion_probe.err; tail -2 gpurun_out/r2f_criterion_probe.err | cut -c1-300; cat gpurun_out/r2f_criterion_probe.json | cut -c1-900
for b in 16 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline --no-parity --no-stock > gpurun_out/r2f_bench_b$b.json 2> gpurun_out/r2f_bench_b$b.err; tail -2 gpurun_out/r2f_bench_b$b.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_b$b.json')); print('B=$b', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['impl_notes']['cuda_graph'][:40]); print(json.dumps(d['roofline']['north_star'])[:1200]); print(d['impl_notes']['routes'])"
done
MPF_B=2 MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/r2f_kernels_step_b2.txt 2> gpurun_out/r2f_kernels_step_b2.err; head -30 gpurun_out/r2f_kernels_step_b2.txt | cut -c1-160
