# Round 2, final single-GPU evidence on the final tree (files r2E_*; the first run of this script produced r2x_*): full GPU suite, ncu launch list (time + DRAM bytes) of one step,
# bench lines of every BASELINE configuration, the reference arm, the per-rank batch of the 8-GPU split.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2E_pytest_gpu.log; cat gpurun_out/r2E_pytest_gpu.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2E_launches_step_b16.csv python benchmarks/bench_step_for_ncu.py > gpurun_out/r2E_launches_step_b16.log 2>&1; tail -1 gpurun_out/r2E_launches_step_b16.log
python scripts/summarize_launches_multi.py gpurun_out/r2E_launches_step_b16.csv > gpurun_out/r2E_launch_summary_step_b16.txt; head -12 gpurun_out/r2E_launch_summary_step_b16.txt
gzip -f gpurun_out/r2E_launches_step_b16.csv
timeout 900 python bench.py > gpurun_out/r2E_bench_n1.json 2> gpurun_out/r2E_bench_n1.err; cut -c1-300 gpurun_out/r2E_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2E_bench_reference_arm.json 2> gpurun_out/r2E_bench_reference_arm.err; cut -c1-400 gpurun_out/r2E_bench_reference_arm.json
for c in 4; do timeout 900 python bench.py --config $c --no-cpu-baseline --no-stock > gpurun_out/r2E_bench_config$c.json 2> gpurun_out/r2E_bench_config$c.err; cut -c1-260 gpurun_out/r2E_bench_config$c.json; done
timeout 600 python bench.py --steps 20 --warmup 3 --batch 2 --no-cpu-baseline --no-stock --no-parity > gpurun_out/r2E_bench_b2.json 2> gpurun_out/r2E_bench_b2.err; cut -c1-260 gpurun_out/r2E_bench_b2.json
