set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_b_gemm.py -q 2>&1 | tail -12 > gpurun_out/pytest_gemm.log; grep -E "^E |FAILED|passed|failed" gpurun_out/pytest_gemm.log | head -12 | cut -c1-250
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; grep -E "FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-200
for bn in 128 256; do echo "BN=$bn"; MPF_GEMM_BN=$bn MPF_PROBE=gemm,masklogits MPF_REPS=5 timeout 200 python benchmarks/kernel_probe.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        r = json.loads(l); print('   %-60s %.3f ms  tensor %.0f%%' % (r['kernel'][:60], r['ms'], 100*r.get('frac_of_tf32_peak',0)))
"; done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err; tail -2 gpurun_out/bench7.err | cut -c1-300; cat gpurun_out/bench7.json | cut -c1-300
timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/torch_profile_r1h.txt 2>&1
