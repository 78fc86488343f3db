# Round 2, final tree: multi-GPU bench, weak scaling (16 images per GPU).  usage: gpurun --gpus N -- 'bash scripts/gpu_r2_scale3.sh N'
set -x
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --weak > gpurun_out/r2y_bench_n${N}_weak.json 2> gpurun_out/r2y_bench_n${N}_weak.err; python -c "
import json
for l in open('gpurun_out/r2y_bench_n${N}_weak.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=$N weak', d['value'], d['ms_per_step'], d['e2e'], d['config']['images_per_gpu'], d['config']['global_batch'], d['scaling'])"
