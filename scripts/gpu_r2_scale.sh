# Round 2: multi-GPU bench.  usage: gpurun --gpus N -- 'bash scripts/gpu_r2_scale.sh N'
set -x
N=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for mode in "" "--weak"; do
tag=strong; [ -n "$mode" ] && tag=weak
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 $mode > gpurun_out/r2s_bench_n${N}_${tag}.json 2> gpurun_out/r2s_bench_n${N}_${tag}.err; tail -3 gpurun_out/r2s_bench_n${N}_${tag}.err | cut -c1-300; python -c "
import json
for l in open('gpurun_out/r2s_bench_n${N}_${tag}.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=$N $tag', d['value'], d['ms_per_step'], d['e2e'], d['config']['images_per_gpu'], d['config']['global_batch'], d['scaling'], d['impl_notes']['cuda_graph'][:30])"
done
