set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python benchmarks/kernel_probe.py > gpurun_out/kernel_probe_r1c.jsonl 2> gpurun_out/kernel_probe.err; tail -2 gpurun_out/kernel_probe.err; cat gpurun_out/kernel_probe_r1c.jsonl | cut -c1-400
timeout 400 python benchmarks/msda_microbench.py > gpurun_out/msda_microbench_r1c.jsonl 2>&1
MPF_REPS=2 MPF_PROBE=msda timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 4 -c 2 -o gpurun_out/prof_msda_r1c python benchmarks/kernel_probe.py > gpurun_out/ncu1.log 2>&1
MPF_REPS=2 MPF_PROBE=gemm timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 2 -c 1 -o gpurun_out/prof_gemm_r1c python benchmarks/kernel_probe.py > gpurun_out/ncu2.log 2>&1
MPF_REPS=2 MPF_PROBE=masklogits timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 2 -c 1 -o gpurun_out/prof_masklogits_r1c python benchmarks/kernel_probe.py > gpurun_out/ncu3.log 2>&1
MPF_REPS=2 MPF_PROBE=xattn timeout 300 ncu --set full --clock-control none --import-source on -k regex:masked_xattn -s 10 -c 1 -o gpurun_out/prof_xattn_r1c python benchmarks/kernel_probe.py > gpurun_out/ncu4.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -2 gpurun_out/bench3.err; cat gpurun_out/bench3.json | cut -c1-600
ls -la gpurun_out | tail -12
