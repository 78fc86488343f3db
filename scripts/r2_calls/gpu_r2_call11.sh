# Round 2, GPU call 11: tight gradient test, ncu launch list (time + DRAM bytes) of one bench step, ncu --set full of
# the attention forward at B=2 (key split) and B=16.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_c_modules.py -m gpu -q -s -k "teacher_forced_tight or matches_oracle_grads" 2>&1 | grep -E "worst five|passed|failed|assert" | cut -c1-700
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2k_launches_step_b16.csv python benchmarks/bench_step_for_ncu.py > gpurun_out/r2k_launches_step_b16.log 2>&1; tail -2 gpurun_out/r2k_launches_step_b16.log
python scripts/summarize_launches_multi.py gpurun_out/r2k_launches_step_b16.csv > gpurun_out/r2k_launch_summary_step_b16.txt; head -30 gpurun_out/r2k_launch_summary_step_b16.txt
gzip -f gpurun_out/r2k_launches_step_b16.csv
MPF_B=2 MPF_PROBE=xattn timeout 300 ncu --set full --clock-control none --import-source on -k regex:'masked_xattn_fwd_kernel' -c 1 --launch-skip 2 -o gpurun_out/r2k_ncu_xattn_fwd_b2 python benchmarks/kernel_probe.py > gpurun_out/r2k_ncu_xattn_fwd_b2.log 2>&1; tail -2 gpurun_out/r2k_ncu_xattn_fwd_b2.log
MPF_B=16 MPF_PROBE=xattn timeout 300 ncu --set full --clock-control none --import-source on -k regex:'masked_xattn_fwd_kernel' -c 1 --launch-skip 2 -o gpurun_out/r2k_ncu_xattn_fwd_b16 python benchmarks/kernel_probe.py > gpurun_out/r2k_ncu_xattn_fwd_b16.log 2>&1; tail -2 gpurun_out/r2k_ncu_xattn_fwd_b16.log
MPF_B=2 MPF_PROBE=xattn timeout 200 python benchmarks/kernel_probe.py 2>&1 | tail -6 | cut -c1-300
