# Round 2, GPU call 22: GEMM probes at the per-rank batch of the 8-GPU split (2 images) and decoder-side shapes.
set -x
mkdir -p gpurun_out
for b in 2 16; do
MPF_B=$b MPF_PROBE=gemm,gemm_small timeout 300 python benchmarks/kernel_probe.py 2>&1 | tee gpurun_out/r2v_gemm_probe_b$b.jsonl | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('B=$b %-100s %8.2f us  issue %.2f  hbm %.2f' % (d['kernel'][:100], d['ms']*1e3, d.get('frac_of_tf32_peak',0) if False else d.get('tensor_TFLOPs_issued',0)/1373.2, d.get('frac_of_measured_hbm',0)))
"
done
