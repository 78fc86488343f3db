set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2e_n2.json 2> gpurun_out/bench_r2e_n2.err; tail -2 gpurun_out/bench_r2e_n2.err | cut -c1-300; cut -c1-400 gpurun_out/bench_r2e_n2.json
