#!/bin/bash
# quick pass: stream-kernel tests, spline / rqs microbench rows, ncu capture of the spline kernels.  usage: gpu_r02r.sh <tag>
set -u
TAG=${1:-r02r}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spline_stream.py -q -x -p no:cacheprovider --timeout=120 > gpurun_out/pytest_stream_$TAG.log 2>&1; echo "stream tests rc=$?"; tail -6 gpurun_out/pytest_stream_$TAG.log
timeout 300 python scripts/microbench.py --only rqs,spline_tf > gpurun_out/microbench_$TAG.log 2>&1; echo "microbench rc=$?"; grep -v "full " gpurun_out/microbench_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spline_stream" -c 9 -f -o gpurun_out/${TAG}_spline_tf python scripts/microbench.py --only spline_tf_ncu > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log
