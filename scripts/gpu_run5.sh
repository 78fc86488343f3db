set -x
mkdir -p gpurun_out
timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/torch_profile_r1d.txt 2>&1; tail -60 gpurun_out/torch_profile_r1d.txt | cut -c1-230
timeout 600 python benchmarks/forward_vs_stock.py > gpurun_out/forward_vs_stock_r1d.json 2> gpurun_out/fvs.err; tail -3 gpurun_out/fvs.err; cat gpurun_out/forward_vs_stock_r1d.json
MPF_REPS=2 MPF_PROBE=msda timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 2 -c 1 -o gpurun_out/prof_msda_fwd_r1d python benchmarks/kernel_probe.py > gpurun_out/ncu1.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json | cut -c1-500
