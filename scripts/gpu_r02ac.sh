#!/bin/bash
# pass ac: the model shapes of the reference's notebooks (hidden 128 / 256, 2-D) at 2^20 and 20000 rows.  usage: <tag>
set -u
TAG=${1:-r02ac}
mkdir -p gpurun_out
timeout 600 python scripts/published_target.py --n 1048576 --reps 2 --only nb_realnvp256,nb_maf128,nb_iaf128,nb_spline128 > gpurun_out/notebook_$TAG.jsonl 2> gpurun_out/notebook_$TAG.err; echo "rc=$?"
timeout 300 python scripts/published_target.py --n 20000 --reps 10 --only nb_realnvp256,nb_maf128,nb_iaf128,nb_spline128 >> gpurun_out/notebook_$TAG.jsonl 2>> gpurun_out/notebook_$TAG.err; echo "rc=$?"
cat gpurun_out/notebook_$TAG.jsonl; tail -3 gpurun_out/notebook_$TAG.err
