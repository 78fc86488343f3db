set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_gpu.log | head -40 | cut -c1-400
timeout 600 python benchmarks/forward_vs_stock.py > gpurun_out/forward_vs_stock_r1s.json 2> gpurun_out/fvs.err; tail -3 gpurun_out/fvs.err | cut -c1-300; cat gpurun_out/forward_vs_stock_r1s.json | cut -c1-600
timeout 900 python bench.py > gpurun_out/bench_r1s.json 2> gpurun_out/bench_r1s.err; tail -3 gpurun_out/bench_r1s.err | cut -c1-400; cut -c1-330 gpurun_out/bench_r1s.json
MPF_PROBE=fpn,layernorm timeout 300 python benchmarks/kernel_probe.py > gpurun_out/kernel_probe_r1s.jsonl 2> gpurun_out/kernel_probe.err; cut -c1-250 gpurun_out/kernel_probe_r1s.jsonl; tail -3 gpurun_out/kernel_probe.err
