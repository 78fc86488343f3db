set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_d_decoder_ops.py -m gpu -q -k "conv3x3 or upsample2x_add_channels_last" 2>&1 | tail -40 > gpurun_out/pytest_conv.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_conv.log | head -30 | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_gpu.log | head -40 | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1w.json 2> gpurun_out/bench_r1w.err; tail -3 gpurun_out/bench_r1w.err | cut -c1-400; cut -c1-330 gpurun_out/bench_r1w.json
timeout 600 python benchmarks/forward_vs_stock.py > gpurun_out/forward_vs_stock_r1w.json 2> gpurun_out/fvs.err; tail -3 gpurun_out/fvs.err | cut -c1-300; cat gpurun_out/forward_vs_stock_r1w.json | cut -c1-700
MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/kernels_r1w.txt 2> gpurun_out/kernels_r1w.err; head -24 gpurun_out/kernels_r1w.txt | cut -c1-200; grep -E "_Conv3x3|upsample|cudnn|sgemm" gpurun_out/kernels_r1w.txt | cut -c1-160
