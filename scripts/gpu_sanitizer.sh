# compute-sanitizer passes over the hand-written kernels (the reference has no race / memory checking).
# Not run in round 1 (GPU budget spent on parity + profiling); second GPU call of the next round:
#   gpurun --timeout 1500 -- 'bash scripts/gpu_sanitizer.sh'
# memcheck over the small-geometry parity tests of every kernel family, racecheck (shared-memory hazards) over the
# kernels that stage through shared memory with hand-placed barriers (matcher cost / LSAP, MSDeformAttn, row ops).
set -x
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
K="not full_size and not bench_geometry and not 16384"
timeout 700 $SAN --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_matcher.log \
  python -m pytest tests/test_gpu_g_matcher.py tests/test_gpu_h_criterion.py tests/test_gpu_i_inference.py -m gpu -q -x -k "$K" > gpurun_out/memcheck_matcher.out 2>&1
echo "memcheck matcher/criterion/inference rc=$?"; tail -3 gpurun_out/memcheck_matcher.out; grep -c "ERROR SUMMARY: 0 errors" gpurun_out/memcheck_matcher.log
timeout 500 $SAN --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_msda.log \
  python -m pytest tests/test_gpu_a_msda.py -m gpu -q -x -k "tiny or small or ragged" > gpurun_out/memcheck_msda.out 2>&1
echo "memcheck msda rc=$?"; tail -3 gpurun_out/memcheck_msda.out
timeout 500 $SAN --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck_matcher.log \
  python -m pytest tests/test_gpu_g_matcher.py -m gpu -q -x -k "golden or lsap" > gpurun_out/racecheck_matcher.out 2>&1
echo "racecheck matcher rc=$?"; tail -3 gpurun_out/racecheck_matcher.out; tail -5 gpurun_out/racecheck_matcher.log
