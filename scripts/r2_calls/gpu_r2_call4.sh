# Round 2, GPU call 4: first run of the TMA-staged MSDeformAttn kernels (parity tests, A/B probe, compute-sanitizer).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_a_msda.py tests/test_gpu_j_full_config.py -m gpu -q -x -k "not gradcheck" 2>&1 | tail -40 > gpurun_out/r2d_pytest_msda.log; tail -25 gpurun_out/r2d_pytest_msda.log | cut -c1-300
timeout 300 python benchmarks/msda_enc_probe.py > gpurun_out/r2d_msda_enc_probe.jsonl 2> gpurun_out/r2d_msda_enc_probe.err; tail -3 gpurun_out/r2d_msda_enc_probe.err | cut -c1-300; cut -c1-1200 gpurun_out/r2d_msda_enc_probe.jsonl
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 10 --log-file gpurun_out/r2d_memcheck_msda_staged.log python -m pytest tests/test_gpu_a_msda.py -m gpu -q -x -k "staged_tile or encoder_fused" -p no:cacheprovider > gpurun_out/r2d_memcheck_msda_staged.out 2>&1; tail -3 gpurun_out/r2d_memcheck_msda_staged.out | cut -c1-200; grep -E "ERROR SUMMARY" gpurun_out/r2d_memcheck_msda_staged.log
timeout 600 python -m pytest tests -m gpu -q -x -k "not gradcheck" 2>&1 | tail -15 > gpurun_out/r2d_pytest_gpu.log; tail -8 gpurun_out/r2d_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-stock > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; tail -3 gpurun_out/r2d_bench_n1.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2d_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['impl_notes']['cuda_graph']); print(json.dumps(d['roofline']['north_star'])[:900])"
