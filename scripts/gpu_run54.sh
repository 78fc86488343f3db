set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_g_matcher.py -m gpu -q 2>&1 | tail -150 > gpurun_out/pytest_matcher.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_matcher.log | head -40 | cut -c1-300
timeout 60 python benchmarks/matcher_probe.py > gpurun_out/matcher_probe.json 2> gpurun_out/matcher_probe.err; tail -3 gpurun_out/matcher_probe.err | cut -c1-300; cat gpurun_out/matcher_probe.json
