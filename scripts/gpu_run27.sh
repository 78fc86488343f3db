set -x
mkdir -p gpurun_out
# A/B of the GroupNorm kernels (r1n 180 ms/step -> r1o 215 ms/step): complete kernel tables of one eager step
MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/kernels_r1p_gn.txt 2> gpurun_out/kernels_gn.err; head -40 gpurun_out/kernels_r1p_gn.txt | cut -c1-180
MPF_NO_GN_KERNEL=1 MPF_KERNELS=1 timeout 300 python benchmarks/torch_profile_step.py > gpurun_out/kernels_r1p_nogn.txt 2> gpurun_out/kernels_nogn.err; head -40 gpurun_out/kernels_r1p_nogn.txt | cut -c1-180
# the driver's default invocations, timed
( time timeout 900 python bench.py > gpurun_out/bench_r1p_default.json 2> gpurun_out/bench_r1p_default.err ) 2>&1 | tail -4; cut -c1-300 gpurun_out/bench_r1p_default.json
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1p_reference.json 2> gpurun_out/bench_r1p_reference.err ) 2>&1 | tail -4; cat gpurun_out/bench_r1p_reference.json
MPF_NO_GN_KERNEL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1p_nogn.json 2> gpurun_out/bench_r1p_nogn.err; cut -c1-300 gpurun_out/bench_r1p_nogn.json
