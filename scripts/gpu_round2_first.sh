# First GPU call of the next round: validates what round 1 could only test on the CPU after its GPU budget ran out
# (bench.py --criterion), times the criterion side at the bench geometry and captures its kernels with ncu.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2_first.sh'
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_gpu.log | head -20 | cut -c1-300
timeout 120 python benchmarks/matcher_probe.py > gpurun_out/matcher_probe.json 2> gpurun_out/matcher_probe.err; cat gpurun_out/matcher_probe.json
MPF_SORT_POINTS=1 timeout 120 python benchmarks/matcher_probe.py > gpurun_out/matcher_probe_sorted.json 2>> gpurun_out/matcher_probe.err; cat gpurun_out/matcher_probe_sorted.json
timeout 200 python benchmarks/lsap_fuzz.py > gpurun_out/lsap_fuzz.json 2> gpurun_out/lsap_fuzz.err; cat gpurun_out/lsap_fuzz.json
timeout 200 python benchmarks/criterion_probe.py > gpurun_out/criterion_probe.json 2> gpurun_out/criterion_probe.err; tail -2 gpurun_out/criterion_probe.err | cut -c1-300; cat gpurun_out/criterion_probe.json
timeout 300 python bench.py --steps 5 --warmup 3 --criterion --no-cpu-baseline > gpurun_out/bench_criterion.json 2> gpurun_out/bench_criterion.err; tail -3 gpurun_out/bench_criterion.err | cut -c1-300; cut -c1-400 gpurun_out/bench_criterion.json
timeout 300 python bench.py --steps 5 --warmup 3 --criterion --no-graph --no-e2e --no-cpu-baseline > gpurun_out/bench_criterion_eager.json 2> gpurun_out/bench_criterion_eager.err; cut -c1-400 gpurun_out/bench_criterion_eager.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'match_cost_partial|lsap_kernel|point_sample_rows' -c 6 -o gpurun_out/ncu_matcher python benchmarks/matcher_probe.py > gpurun_out/ncu_matcher.log 2>&1; tail -3 gpurun_out/ncu_matcher.log
