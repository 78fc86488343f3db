# compute-sanitizer passes over the hand-written kernels (the reference has no race / memory checking).
#   gpurun --timeout 2400 -- 'bash scripts/gpu_sanitizer.sh'
# memcheck over the small-geometry parity tests of every kernel family (incl. the criterion's failed-assignment path
# and point_sample_rows_bwd's atomics), initcheck over the grad_value memset paths / atomically accumulated buffers,
# racecheck (shared-memory hazards) over the matcher / MSDeformAttn / row kernels AND the mbarrier pipelines (GEMMs,
# cross-attention forward + backward, TMA reduce-add epilogue).  Summaries -> gpurun_out/san_*.log (copied to profiles/).
set -x
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
K="not full_size and not bench_geometry and not 16384 and not config1 and not config2 and not gradcheck"
run() {  # tool, tag, time limit, pytest args...
  tool=$1; tag=$2; lim=$3; shift 3
  timeout $lim $SAN --tool $tool --error-exitcode 9 --print-limit 20 --log-file gpurun_out/san_${tool}_${tag}.log \
    python -m pytest "$@" -m gpu -q -x -p no:cacheprovider > gpurun_out/san_${tool}_${tag}.out 2>&1
  echo "== $tool $tag rc=$?"; tail -2 gpurun_out/san_${tool}_${tag}.out | cut -c1-200; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/san_${tool}_${tag}.log | tail -3
}
run memcheck matcher_criterion_inference 600 tests/test_gpu_g_matcher.py tests/test_gpu_h_criterion.py tests/test_gpu_i_inference.py -k "$K"
run memcheck msda 400 tests/test_gpu_a_msda.py -k "golden_cases or vec_path or encoder_fused or error_behaviour"
run memcheck gemm_xattn 600 tests/test_gpu_b_gemm.py tests/test_gpu_b2_gemm_bf16x3.py tests/test_gpu_d_decoder_ops.py -k "$K"
run memcheck modules 600 tests/test_gpu_c_modules.py -k "golden or attn_mask_bits"
run initcheck msda_criterion 500 tests/test_gpu_a_msda.py tests/test_gpu_h_criterion.py -k "golden_cases or vec_path or point_sample or failed_assignment"
run racecheck matcher_msda 500 tests/test_gpu_g_matcher.py tests/test_gpu_a_msda.py -k "golden or lsap or vec_path or encoder_fused"
run racecheck gemm_xattn 700 tests/test_gpu_b2_gemm_bf16x3.py tests/test_gpu_d_decoder_ops.py -k "$K"
