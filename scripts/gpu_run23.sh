set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_e_graph.py tests/test_gpu_c_modules.py -q 2>&1 | tail -40 > gpurun_out/pytest_graph.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_graph.log | head -30 | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; tail -3 gpurun_out/bench_r1m.err | cut -c1-400; cut -c1-3000 gpurun_out/bench_r1m.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph --no-e2e > gpurun_out/bench_r1m_nograph.json 2> gpurun_out/bench_r1m_nograph.err; tail -3 gpurun_out/bench_r1m_nograph.err | cut -c1-300; cut -c1-330 gpurun_out/bench_r1m_nograph.json
MPF_SHAPES=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/torch_profile_shapes_r1m.txt 2>&1
