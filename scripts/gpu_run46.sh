set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_gpu.log | head -40 | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/bench_r2e_default.json 2> gpurun_out/bench_r2e_default.err ) 2>&1 | tail -4; cut -c1-330 gpurun_out/bench_r2e_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2e_reference.json 2> gpurun_out/bench_r2e_reference.err; cut -c1-300 gpurun_out/bench_r2e_reference.json
timeout 600 python benchmarks/forward_vs_stock.py > gpurun_out/forward_vs_stock_r2e.json 2> gpurun_out/fvs.err; tail -2 gpurun_out/fvs.err | cut -c1-300; cut -c1-700 gpurun_out/forward_vs_stock_r2e.json
MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/kernels_r2e.txt 2> gpurun_out/kernels_r2e.err; head -3 gpurun_out/kernels_r2e.txt | cut -c1-200
