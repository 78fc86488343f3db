set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
nproc
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 > gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python benchmarks/msda_microbench.py > gpurun_out/msda_microbench.jsonl 2> gpurun_out/msda_microbench.err; tail -3 gpurun_out/msda_microbench.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -3 gpurun_out/bench1.err; cat gpurun_out/bench1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 1 --warmup 1 --batch 16 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
MPF_B=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 12 -c 4 -o gpurun_out/prof_msda_r1a python benchmarks/msda_microbench.py > gpurun_out/ncu_msda.log 2>&1
ls -la gpurun_out
