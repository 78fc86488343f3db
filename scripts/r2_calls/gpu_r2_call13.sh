# Round 2, GPU call 13: lsap with the cost matrix in shared memory; GEMM epilogue with two staging tiles per group
# (A/B over MPF_GEMM_SBUFS / MPF_GEMM_STAGES / MPF_GEMM_BN); ncu of the attention forward at HW = 16384.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_g_matcher.py tests/test_gpu_h_criterion.py tests/test_gpu_b2_gemm_bf16x3.py -m gpu -q -x 2>&1 | tail -5
for cfg in "1 0 0" "1 2 0" "2 2 0" "2 0 128" "1 0 128"; do
  set -- $cfg
  echo "== SBUFS=$1 STAGES=$2 BN=$3"
  env MPF_GEMM_SBUFS=$1 $( [ "$2" != 0 ] && echo MPF_GEMM_STAGES=$2 ) $( [ "$3" != 0 ] && echo MPF_GEMM_BN=$3 ) MPF_PROBE=gemm timeout 300 python benchmarks/kernel_probe.py 2>&1 | tee gpurun_out/r2m_gemm_probe_sbufs$1_stages$2_bn$3.jsonl | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('%-90s %.4f ms' % (d['kernel'][:90], d['ms']))
"
done
MPF_HW=16384 MPF_B=16 MPF_PROBE=xattn timeout 300 ncu --set full --clock-control none --import-source on -k regex:'masked_xattn_fwd_kernel' -c 1 --launch-skip 2 -o gpurun_out/r2m_ncu_xattn_fwd_b16_hw16384 python benchmarks/kernel_probe.py > gpurun_out/r2m_ncu_xattn_fwd.log 2>&1; tail -2 gpurun_out/r2m_ncu_xattn_fwd.log
MPF_HW=16384 MPF_B=2 MPF_PROBE=xattn timeout 300 ncu --set full --clock-control none --import-source on -k regex:'masked_xattn_fwd_kernel' -c 1 --launch-skip 2 -o gpurun_out/r2m_ncu_xattn_fwd_b2_hw16384 python benchmarks/kernel_probe.py > gpurun_out/r2m_ncu_xattn_fwd_b2.log 2>&1; tail -2 gpurun_out/r2m_ncu_xattn_fwd_b2.log
timeout 300 python benchmarks/matcher_probe.py 2>&1 | tail -4 | cut -c1-400
