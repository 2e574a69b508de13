#!/bin/bash
# ncu --set full of the stand-alone HBM kernels (one launch each) driven by scripts/microbench.py
# usage: scripts/gpu_profile_micro.sh <only-list> <kernel regex> <tag>
set -u
ONLY=${1:-rqs}; KRE=${2:-rqs_unit_fwd}; TAG=${3:-r01_micro}
CMD="python scripts/microbench.py --only $ONLY"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
