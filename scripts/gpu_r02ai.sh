#!/bin/bash
# launch lists of the 2-D training steps at 2^20 rows (C1 / C2 shapes through the layered route).  usage: <tag>
set -u
TAG=${1:-r02ai}
mkdir -p gpurun_out
for m in realnvp2 spline2; do
  timeout 300 python scripts/train_step_bench.py --model $m --batch 1048576 --steps 3 > gpurun_out/train_${m}_$TAG.json 2>&1; echo "$m rc=$?"; cut -c1-200 gpurun_out/train_${m}_$TAG.json
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_${m}_launches.csv python scripts/train_step_bench.py --model $m --batch 1048576 --steps 1 --warmup 1 > gpurun_out/ncu_${m}_$TAG.log 2>&1; echo "ncu $m rc=$?"
done
