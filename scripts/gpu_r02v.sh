#!/bin/bash
# pass v: the four published configs at n = 2^20 and n = 4000, both directions (timings + launch list).  usage: gpu_r02v.sh <tag>
set -u
TAG=${1:-r02v}
mkdir -p gpurun_out
timeout 300 python scripts/published_target.py --n 1048576 > gpurun_out/published_$TAG.jsonl 2> gpurun_out/published_$TAG.err; echo "rc=$?"
timeout 300 python scripts/published_target.py --n 4000 --reps 20 >> gpurun_out/published_$TAG.jsonl 2>> gpurun_out/published_$TAG.err; echo "rc=$?"
cat gpurun_out/published_$TAG.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_published_launches.csv python scripts/published_target.py --n 1048576 --reps 1 > gpurun_out/ncu_published_$TAG.log 2>&1; echo "ncu rc=$?"
