#!/bin/bash
# pass ar: ncu of the mma.sync in-block kernel (blocked sampler, variant 2): launch list of one pass + full capture of two launches
set -u
TAG=${1:-r02ar}
mkdir -p gpurun_out
export NFB200_OPTIONS=3:2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tc2_kernel|ar_block|ar_finish" -s 100 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/sampler_target.py --reps 1 > gpurun_out/ncu_list_$TAG.log 2>&1; echo "list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ar_block_mma" -s 20 -c 2 -f -o gpurun_out/${TAG}_mma python scripts/sampler_target.py --reps 1 > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log
