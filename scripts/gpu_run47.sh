set -x
mkdir -p gpurun_out
MPF_PROBE=msda MPF_REPS=10 timeout 300 python benchmarks/kernel_probe.py 2>&1 | grep enc_bwd | cut -c1-200
MPF_MSDA_BWD_OCC4=1 MPF_PROBE=msda MPF_REPS=10 timeout 300 python benchmarks/kernel_probe.py 2>&1 | grep enc_bwd | cut -c1-200
