#!/bin/bash
# launch list (ncu gpu__time_duration) of one pass of the blocked sampler, current default route
set -u
TAG=${1:-r02bd}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tc2_kernel|ar_block|ar_finish" -s 75 -c 45 --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/sampler_target.py --reps 1 > gpurun_out/ncu_list_$TAG.log 2>&1; echo "list rc=$?"
