#!/bin/bash
# ncu --set full of the last blocks' kernels of the blocked sampler (current default route): pull GEMMs, in-block kernel, push
set -u
TAG=${1:-r02bi}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2_kernel|ar_block_mma" -s 100 -c 9 -f -o gpurun_out/${TAG}_sampler python scripts/sampler_target.py --reps 1 > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log
