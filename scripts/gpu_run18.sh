set -x
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; grep -E "^E |FAILED|passed|failed" gpurun_out/pytest_gpu.log | head | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; tail -2 gpurun_out/bench_r1h.err | cut -c1-300; cat gpurun_out/bench_r1h.json | cut -c1-1500
timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/torch_profile_r1h.txt 2>&1
timeout 400 python benchmarks/forward_vs_stock.py > gpurun_out/forward_vs_stock_r1h.json 2> gpurun_out/fvs.err; cat gpurun_out/forward_vs_stock_r1h.json | cut -c1-1200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_r1h.csv 3 > gpurun_out/launch_summary_r1h.txt 2>&1; head -30 gpurun_out/launch_summary_r1h.txt
