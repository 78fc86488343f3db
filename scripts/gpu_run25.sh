set -x
mkdir -p gpurun_out
timeout 300 python benchmarks/kernel_probe.py > gpurun_out/kernel_probe_r1n.jsonl 2> gpurun_out/kernel_probe.err; cut -c1-260 gpurun_out/kernel_probe_r1n.jsonl; tail -3 gpurun_out/kernel_probe.err
MPF_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16x3|msda_enc|masked_xattn|add_layernorm" -c 26 -o gpurun_out/ncu_r1n python benchmarks/kernel_probe.py > gpurun_out/ncu_r1n.log 2>&1; tail -3 gpurun_out/ncu_r1n.log
