set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed|Error" gpurun_out/pytest_gpu.log | head -30 | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1o.json 2> gpurun_out/bench_r1o.err; tail -3 gpurun_out/bench_r1o.err | cut -c1-400; cut -c1-330 gpurun_out/bench_r1o.json
NCU="ncu --set full --clock-control none"
cap() { # name probe regex skip [extra]
  MPF_PROBE=$2 MPF_REPS=1 timeout 300 $NCU $5 -k regex:$3 -s $4 -c 1 -o gpurun_out/ncu_r1o_$1 python benchmarks/kernel_probe.py > gpurun_out/ncu_$1.log 2>&1
  python scripts/ncu_summary.py gpurun_out/ncu_r1o_$1.ncu-rep > gpurun_out/ncu_r1o_$1.txt 2>/dev/null
}
cap gemm_ffn1 gemm gemm_bf16x3_kernel 2 "--import-source on"
cap gemm_tn gemm gemm_bf16x3_tn 2
cap msda_enc_fwd msda msda_enc_fwd 2
cap msda_enc_bwd msda msda_enc_bwd 2
cap xattn_fwd_16384 xattn masked_xattn_fwd 8
cap layernorm_fwd layernorm add_layernorm_fwd 2
du -sm gpurun_out
if [ $(du -sm gpurun_out | cut -f1) -gt 58 ]; then rm -f gpurun_out/ncu_r1o_msda_enc_bwd.ncu-rep gpurun_out/ncu_r1o_layernorm_fwd.ncu-rep gpurun_out/ncu_r1o_gemm_tn.ncu-rep; fi
du -sm gpurun_out
