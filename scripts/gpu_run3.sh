set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_d_decoder_ops.py -q -x 2>&1 | tail -40 > gpurun_out/pytest_d.log; tail -25 gpurun_out/pytest_d.log | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -3 gpurun_out/bench2.err; cat gpurun_out/bench2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 1 --batch 16 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
