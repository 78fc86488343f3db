# Round 2, final tree: multi-GPU bench, BASELINE split (global batch 16 over the ranks).  usage: gpurun --gpus N -- 'bash scripts/gpu_r2_scale2.sh N'
set -x
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2F_bench_n${N}_strong.json 2> gpurun_out/r2F_bench_n${N}_strong.err; tail -3 gpurun_out/r2F_bench_n${N}_strong.err | cut -c1-300; python -c "
import json
for l in open('gpurun_out/r2F_bench_n${N}_strong.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=$N strong', d['value'], d['ms_per_step'], d['e2e'], d['config']['images_per_gpu'], d['config']['global_batch'], d['scaling'])"
