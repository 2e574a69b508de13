#!/bin/bash
# pass as: mma.sync in-block kernel after the cheaper split: sampler timing for both variants, per-launch time list, one full ncu capture
set -u
TAG=${1:-r02as}
mkdir -p gpurun_out
for v in 2 3; do
  timeout 120 python scripts/sampler_target.py --variant $v --reps 10 > gpurun_out/sampler_${TAG}_v$v.json 2> gpurun_out/sampler_${TAG}_v$v.err; echo "sampler v$v rc=$?"; cat gpurun_out/sampler_${TAG}_v$v.json; tail -3 gpurun_out/sampler_${TAG}_v$v.err
done
export NFB200_OPTIONS=3:3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ar_block_mma" -s 20 -c 1 -f -o gpurun_out/${TAG}_mma python scripts/sampler_target.py --reps 1 > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log
