mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_h_criterion.py -m gpu -q -x 2>&1 | tail -60 > gpurun_out/pytest_criterion.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_criterion.log | head -20 | cut -c1-300
