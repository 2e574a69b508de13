#!/bin/bash
set -u
mkdir -p gpurun_out
for m in realnvp2 spline2 maf64; do
  for g in "" "--graph"; do
    timeout 300 python scripts/train_step_bench.py --model $m --batch 5000 --steps 50 --warmup 5 $g > gpurun_out/c1_${m}${g}.json 2> gpurun_out/c1_${m}${g}.err; echo "$m $g rc=$?"; tail -1 gpurun_out/c1_${m}${g}.json; tail -3 gpurun_out/c1_${m}${g}.err
  done
done
