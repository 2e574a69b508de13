#!/bin/bash
# Round 2, pass q: TMA-staged compact spline transform kernels (spline_stream.cu) + the single-evaluation reverse sweep.
# New tests first under a short timeout (a wrong mbarrier phase would hang), then the suite, then the microbench rows.
set -u
TAG=${1:-r02q}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spline_stream.py -q -x -p no:cacheprovider --timeout=120 > gpurun_out/pytest_stream_$TAG.log 2>&1; echo "stream tests rc=$?"; tail -15 gpurun_out/pytest_stream_$TAG.log
timeout 300 python scripts/microbench.py --only rqs,spline_tf > gpurun_out/microbench_$TAG.log 2>&1; echo "microbench rc=$?"; cat gpurun_out/microbench_$TAG.log
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
