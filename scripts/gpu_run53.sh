set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_gpu.log | head -40 | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; tail -2 gpurun_out/bench_r2i.err | cut -c1-300; cut -c1-330 gpurun_out/bench_r2i.json
