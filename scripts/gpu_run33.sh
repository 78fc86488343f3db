set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_gpu.log | head -40 | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1t.json 2> gpurun_out/bench_r1t.err; tail -3 gpurun_out/bench_r1t.err | cut -c1-400; cut -c1-330 gpurun_out/bench_r1t.json
MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/kernels_r1t.txt 2> gpurun_out/kernels_r1t.err; head -45 gpurun_out/kernels_r1t.txt | cut -c1-200; tail -3 gpurun_out/kernels_r1t.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
