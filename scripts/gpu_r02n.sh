#!/bin/bash
# Round 2, pass n: ONE ncu capture of the persistent in-block kernel of the blocked sampler.  usage: gpu_r02n.sh <tag>
set -u
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout 300 python scripts/sampler_target.py > gpurun_out/sampler_$TAG.json 2> gpurun_out/sampler_$TAG.err; echo "sampler rc=$?"; cat gpurun_out/sampler_$TAG.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ar_block_warp -s 10 -c 1 -o gpurun_out/prof_${TAG}_arblock -f python scripts/sampler_target.py --reps 1 > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu rc=$?"
