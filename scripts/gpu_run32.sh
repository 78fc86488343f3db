set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r1s_n2.json 2> gpurun_out/bench_r1s_n2.err; tail -3 gpurun_out/bench_r1s_n2.err | cut -c1-300; cut -c1-400 gpurun_out/bench_r1s_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r1s_n2_ref.json 2> gpurun_out/bench_r1s_n2_ref.err; tail -2 gpurun_out/bench_r1s_n2_ref.err | cut -c1-300; cut -c1-300 gpurun_out/bench_r1s_n2_ref.json
