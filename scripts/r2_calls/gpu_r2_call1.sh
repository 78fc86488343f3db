# Round 2, GPU call 1: full GPU suite (new full-size parity tests), smoke, the new default bench (criterion in the
# step, parity + stock blocks), the round-1 mode for comparison, the per-rank workload of the 8-GPU split (B=2).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -60 > gpurun_out/r2a_pytest_gpu.log; tail -30 gpurun_out/r2a_pytest_gpu.log | cut -c1-400
timeout 300 python -m pytest tests/test_gpu_j_full_config.py -m gpu -q -s 2>&1 | tail -30 > gpurun_out/r2a_pytest_full_config.log; grep -E "flip|passed|failed|Error|assert" gpurun_out/r2a_pytest_full_config.log | cut -c1-600
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; tail -3 gpurun_out/r2a_bench_n1.err | cut -c1-300; cut -c1-1500 gpurun_out/r2a_bench_n1.json
timeout 400 python bench.py --steps 10 --warmup 3 --loss pseudo --no-cpu-baseline --no-parity --no-stock > gpurun_out/r2a_bench_n1_pseudo.json 2> gpurun_out/r2a_bench_n1_pseudo.err; tail -3 gpurun_out/r2a_bench_n1_pseudo.err | cut -c1-300; cut -c1-700 gpurun_out/r2a_bench_n1_pseudo.json
timeout 400 python bench.py --steps 20 --warmup 3 --batch 2 --no-cpu-baseline --no-parity --no-stock > gpurun_out/r2a_bench_b2.json 2> gpurun_out/r2a_bench_b2.err; tail -3 gpurun_out/r2a_bench_b2.err | cut -c1-300; cut -c1-700 gpurun_out/r2a_bench_b2.json
MPF_B=2 MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/r2a_kernels_step_b2.txt 2> gpurun_out/r2a_kernels_step_b2.err; head -50 gpurun_out/r2a_kernels_step_b2.txt | cut -c1-180
MPF_B=16 MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/r2a_kernels_step_b16.txt 2> gpurun_out/r2a_kernels_step_b16.err; head -40 gpurun_out/r2a_kernels_step_b16.txt | cut -c1-180
timeout 200 python benchmarks/criterion_probe.py > gpurun_out/r2a_criterion_probe.json 2> gpurun_out/r2a_criterion_probe.err; tail -2 gpurun_out/r2a_criterion_probe.err | cut -c1-300; cat gpurun_out/r2a_criterion_probe.json | cut -c1-1500
timeout 120 python benchmarks/matcher_probe.py > gpurun_out/r2a_matcher_probe.json 2> gpurun_out/r2a_matcher_probe.err; cat gpurun_out/r2a_matcher_probe.json | cut -c1-800
MPF_SORT_POINTS=1 timeout 120 python benchmarks/matcher_probe.py > gpurun_out/r2a_matcher_probe_sorted.json 2>> gpurun_out/r2a_matcher_probe.err; cat gpurun_out/r2a_matcher_probe_sorted.json | cut -c1-800
