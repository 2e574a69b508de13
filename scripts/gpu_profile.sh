#!/bin/bash
# Run on the B200 box (gpurun): plain bench run, then the ncu launch list and one full capture of the top kernel.
# usage: scripts/gpu_profile.sh <workload c2|c3> <kernel regex> <tag> [skip launches] [count]
set -u
WL=${1:-c2}; KRE=${2:-spline_stack}; TAG=${3:-r01_$WL}; SKIP=${4:-4}; CNT=${5:-2}
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $CNT -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
