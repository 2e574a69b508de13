#!/bin/bash
# Round 2, pass p: state check after the container re-creation (full GPU suite on HEAD), spline / rqs microbench rows,
# and one ncu --set full capture of the compact spline transform kernels (forward + backward) to guide their rework.
set -u
TAG=${1:-r02p}
mkdir -p gpurun_out
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
timeout 300 python scripts/microbench.py --only rqs,spline_tf > gpurun_out/microbench_$TAG.log 2>&1; echo "microbench rc=$?"; cat gpurun_out/microbench_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spline_transform_compact|rqs_unit" -c 12 -f -o gpurun_out/${TAG}_spline_tf python scripts/microbench.py --only spline_tf_ncu > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$TAG.log
