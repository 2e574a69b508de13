#!/bin/bash
# pass t: where the training steps of C4 / C5 spend their time -- step timings in the three precision modes and an ncu
# launch list (gpu__time_duration per launch) of one step each.  usage: gpu_r02t.sh <tag>
set -u
TAG=${1:-r02t}
mkdir -p gpurun_out
: > gpurun_out/train_step_$TAG.jsonl
for m in "realnvp256 65536" "spline784 4096" "maf256 65536" "maf64 262144" "spline2 1048576" "realnvp2 1048576"; do
  set -- $m
  for p in fp32 bf16; do
    timeout 300 python scripts/train_step_bench.py --model $1 --batch $2 --steps 5 --precision $p >> gpurun_out/train_step_$TAG.jsonl 2>> gpurun_out/train_step_$TAG.err; echo "$1 $p rc=$?"
  done
done
cat gpurun_out/train_step_$TAG.jsonl | cut -c1-400
for m in "realnvp256 65536" "spline784 4096"; do
  set -- $m
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 0 -c 4000 --csv --log-file gpurun_out/${TAG}_$1_launches.csv python scripts/train_step_bench.py --model $1 --batch $2 --steps 1 --warmup 1 > gpurun_out/ncu_$1_$TAG.log 2>&1; echo "ncu $1 rc=$?"
done
