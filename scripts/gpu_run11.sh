set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_d_decoder_ops.py -q -x 2>&1 | tail -30 > gpurun_out/pytest_d.log; grep -E "^E |FAILED|passed|failed|Error" gpurun_out/pytest_d.log | head -20 | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; grep -E "FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-200
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; tail -2 gpurun_out/bench6.err | cut -c1-300; cat gpurun_out/bench6.json | cut -c1-300
timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/torch_profile_r1g.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 1 --warmup 1 --batch 16 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
