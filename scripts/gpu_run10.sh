set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_d_decoder_ops.py -q -x 2>&1 | tail -30 > gpurun_out/pytest_d.log; grep -E "^E |FAILED|passed|failed|Error" gpurun_out/pytest_d.log | head -20 | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; grep -E "FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-200
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -2 gpurun_out/bench5.err | cut -c1-300; cat gpurun_out/bench5.json | cut -c1-300
MPF_SHAPES=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/torch_profile_shapes_r1f.txt 2>&1
timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/torch_profile_r1f.txt 2>&1
timeout 600 python benchmarks/forward_vs_stock.py > gpurun_out/forward_vs_stock_r1f.json 2> gpurun_out/fvs.err; cat gpurun_out/forward_vs_stock_r1f.json
