#!/bin/bash
# one ncu --set full capture of the C2 headline kernel (two launches: log_prob + sample).  usage: gpu_ncu_c2.sh <tag>
set -u
TAG=${1:-r02u}
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-also > gpurun_out/bench_plain_$TAG.json 2>&1; echo "plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spline_stack_tc -s 4 -c 2 -o gpurun_out/prof_${TAG}_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-also > gpurun_out/ncu_full_${TAG}_c2.log 2>&1; echo "c2 full rc=$?"
