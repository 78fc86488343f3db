set -x
mkdir -p gpurun_out
# launch list of one bench command (eager launches so that every kernel is a separate launch record)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 10000 --csv --log-file gpurun_out/launches_r1x.csv python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
python scripts/summarize_launches.py gpurun_out/launches_r1x.csv 1 > gpurun_out/launch_summary_r1x.txt; head -30 gpurun_out/launch_summary_r1x.txt
rm -f gpurun_out/launches_r1x.csv
NCU="ncu --set full --clock-control none"
cap() { # name probe regex skip
  MPF_PROBE=$2 MPF_REPS=1 timeout 300 $NCU -k regex:$3 -s $4 -c 1 -o gpurun_out/ncu_r1x_$1 python benchmarks/kernel_probe.py > gpurun_out/ncu_$1.log 2>&1
  python scripts/ncu_summary.py gpurun_out/ncu_r1x_$1.ncu-rep > gpurun_out/ncu_r1x_$1.txt 2>/dev/null
  rm -f gpurun_out/ncu_r1x_$1.ncu-rep
  grep -E "gpu__time_duration|pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|dram__bytes|lsu_wavefronts|dram_throughput" gpurun_out/ncu_r1x_$1.txt | head -8
}
cap conv_fwd conv gemm_bf16x3_kernel 2
cap conv_wgrad conv gemm_bf16x3_tn 2
cap msda_enc_fwd msda msda_enc_fwd 2
cap xattn_bwd_dkv xattn_bwd masked_xattn_bwd_dkv 5
cap xattn_bwd_dq xattn_bwd masked_xattn_bwd_dq 5
cap upsample_add_cl fpn upsample2x_add_cl_fwd 2
timeout 300 python benchmarks/kernel_probe.py > gpurun_out/kernel_probe_r1x.jsonl 2> gpurun_out/kernel_probe.err; cut -c1-230 gpurun_out/kernel_probe_r1x.jsonl; tail -3 gpurun_out/kernel_probe.err
timeout 900 python -m pytest tests/test_gpu_f_configs.py -m gpu -q 2>&1 | tail -5
