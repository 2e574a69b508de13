#!/bin/bash
# ncu --set full capture of the compact spline transform kernels (forward, inverse, backward on three shapes) + rqs_unit.
# usage: gpu_ncu_spline.sh <tag>
set -u
TAG=${1:-r02q}
mkdir -p gpurun_out
timeout 120 python scripts/microbench.py --only spline_tf_ncu > gpurun_out/ncu_plain_$TAG.log 2>&1; echo "plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spline_stream|spline_transform_compact|rqs_unit" -c 14 -f -o gpurun_out/${TAG}_spline_tf python scripts/microbench.py --only spline_tf_ncu > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$TAG.log
