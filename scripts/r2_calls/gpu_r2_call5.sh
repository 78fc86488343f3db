# Round 2, GPU call 5: staged MSDeformAttn kernels after the code-size fix: parity, A/B probe, ncu --set full of both.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_a_msda.py -m gpu -q -x -k "staged_tile or encoder_fused or golden" 2>&1 | tail -5 | cut -c1-300
timeout 300 python benchmarks/msda_enc_probe.py > gpurun_out/r2e_msda_enc_probe.jsonl 2> gpurun_out/r2e_msda_enc_probe.err; tail -3 gpurun_out/r2e_msda_enc_probe.err | cut -c1-300; python - <<'PY'
import json
for l in open('gpurun_out/r2e_msda_enc_probe.jsonl'):
    d=json.loads(l); print(d['B'], d['shapes'][0], d['offsets'], 'fwd staged/gather %.3f/%.3f  bwd %.3f/%.3f' % (d['staged_fwd_ms'], d['gather_fwd_ms'], d['staged_bwd_ms'], d['gather_bwd_ms']), d['fwd_bit_identical'], d['grad_value_max_abs_diff'])
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'staged_kernel' -c 2 -o gpurun_out/r2e_ncu_msda_staged python benchmarks/msda_enc_probe.py > gpurun_out/r2e_ncu_msda_staged.log 2>&1; tail -2 gpurun_out/r2e_ncu_msda_staged.log
for m in sm__throughput.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed sm__issue_active.avg.pct_of_peak_sustained_elapsed gpu__time_duration.sum l1tex__data_bank_conflicts_pipe_lsu.sum smsp__average_warps_issue_stalled sm__warps_active launch__occupancy lts__throughput dram__bytes l1tex__data_pipe_lsu_wavefronts_mem_shared; do :; done
ncu -i gpurun_out/r2e_ncu_msda_staged.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY'
import csv, sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]
keys=['gpu__time_duration.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__issue_active.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__data_bank_conflicts_pipe_lsu.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','lts__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__inst_executed.sum','sm__inst_executed_pipe_lsu.sum']
ki=hdr.index('Kernel Name')
for r in rows[2:]:
    print('==', r[ki][:60])
    for k in keys:
        if k in hdr: print('   ', k, '=', r[hdr.index(k)])
PY
