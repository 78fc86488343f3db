# Round 2, GPU call 8: full GPU suite, default bench line (all blocks), BASELINE configs 1 / 4 / 5 lines, B=2 profile.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2h_pytest_gpu.log; tail -6 gpurun_out/r2h_pytest_gpu.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; tail -2 gpurun_out/r2h_bench_n1.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2h_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches']); print(d['parity']); print(d['vs_stock_cuda']); print(d['cpu_baseline']); print(d['clocks'])"
for c in 4 5; do
timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-stock > gpurun_out/r2h_bench_config$c.json 2> gpurun_out/r2h_bench_config$c.err; tail -2 gpurun_out/r2h_bench_config$c.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2h_bench_config$c.json')); print('config $c', d['value'], d['ms_per_step'], d['e2e'], d['parity'], d['cpu_baseline']); print(d['impl_notes']['routes'])"
done
timeout 300 python bench.py --config 1 --steps 20 --warmup 5 > gpurun_out/r2h_bench_config1.json 2> gpurun_out/r2h_bench_config1.err; tail -2 gpurun_out/r2h_bench_config1.err | cut -c1-300; cut -c1-900 gpurun_out/r2h_bench_config1.json
MPF_B=2 MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/r2h_kernels_step_b2.txt 2> gpurun_out/r2h_kernels_step_b2.err; head -3 gpurun_out/r2h_kernels_step_b2.txt | cut -c1-160
MPF_B=16 MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/r2h_kernels_step_b16.txt 2> gpurun_out/r2h_kernels_step_b16.err; head -3 gpurun_out/r2h_kernels_step_b16.txt | cut -c1-160
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/r2h_bench_reference_arm.json
