# compute-sanitizer over the kernels added or restructured in the second half of round 2: the in-place operand conversion of
# both bf16x3 GEMMs (named barriers inside the mbarrier pipelines), mask_loss.cu, the streamed matcher sampler (bulk copies
# into shared memory), the assignment kernel with its shared-memory cost matrix, the criterion's joint path.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_sanitizer2.sh'
set -x
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
K="not full_size and not bench_geometry and not 16384 and not config1 and not config2 and not gradcheck and not pairs"
run() {  # tool, tag, time limit, pytest args...
  tool=$1; tag=$2; lim=$3; shift 3
  timeout $lim $SAN --tool $tool --error-exitcode 9 --print-limit 20 --log-file gpurun_out/san2_${tool}_${tag}.log \
    python -m pytest "$@" -m gpu -q -x -p no:cacheprovider > gpurun_out/san2_${tool}_${tag}.out 2>&1
  echo "== $tool $tag rc=$?"; tail -n 2 gpurun_out/san2_${tool}_${tag}.out | cut -c1-200; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/san2_${tool}_${tag}.log | tail -3
}
run memcheck matcher_criterion 300 tests/test_gpu_g_matcher.py tests/test_gpu_h_criterion.py -k "$K"
run racecheck gemm_bf16x3 360 tests/test_gpu_b2_gemm_bf16x3.py -k "$K"
run racecheck matcher_criterion 240 tests/test_gpu_g_matcher.py tests/test_gpu_h_criterion.py -k "$K and (sample_shared or lsap or mask_loss or golden or heads)"
