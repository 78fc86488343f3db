set -x
mkdir -p gpurun_out
timeout 300 python benchmarks/gemm_debug_probe.py > gpurun_out/gemm_debug_probe.jsonl 2> gpurun_out/gemm_debug_probe.err; cat gpurun_out/gemm_debug_probe.jsonl; tail -3 gpurun_out/gemm_debug_probe.err
MPF_PROBE=gemm MPF_REPS=2 timeout 500 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 8 -c 2 -o gpurun_out/ncu_gemm_r1i python benchmarks/kernel_probe.py > gpurun_out/ncu_gemm.log 2>&1; tail -3 gpurun_out/ncu_gemm.log
