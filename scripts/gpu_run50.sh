set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_d_decoder_ops.py -m gpu -q -k "masked_cross_attention" 2>&1 | tail -30 > gpurun_out/pytest_xattn.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_xattn.log | head -30 | cut -c1-400
MPF_PROBE=xattn MPF_REPS=10 timeout 300 python benchmarks/kernel_probe.py 2>&1 | cut -c1-200
