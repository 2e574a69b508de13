#!/bin/bash
# launch list of RealNVP(2, 8, 256) (notebook 1 / plots/fig_gif.py shape) at 2^20 rows.  usage: <tag>
set -u
TAG=${1:-r02ag}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_nb_realnvp256_launches.csv python scripts/published_target.py --n 1048576 --reps 1 --only nb_realnvp256 > gpurun_out/ncu_nb_$TAG.log 2>&1; echo "ncu rc=$?"
