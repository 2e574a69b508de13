#!/bin/bash
# ncu --set full of the blocked sampler's kernels in its third pass: the last pull GEMMs (K = 448) and the last in-block
# launch.  usage: gpu_ncu_sampler.sh <tag>
set -u
TAG=${1:-r02ab}
mkdir -p gpurun_out
timeout 120 python scripts/sampler_target.py --reps 1 > gpurun_out/sampler_plain_$TAG.json 2>&1; echo "plain rc=$?"; cat gpurun_out/sampler_plain_$TAG.json
# pass 3 = launches 2*37.. : skip the first two passes (74 library launches of these kernels ~ 2 x 37), take the next 37
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2_kernel|ar_block_warp" -s 100 -c 10 -f -o gpurun_out/${TAG}_sampler python scripts/sampler_target.py --reps 1 > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log
