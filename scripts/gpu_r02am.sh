#!/bin/bash
# pass am: MADE stack tensor-core kernel: published tests, goldens / parity, timings, full suite.  usage: <tag>
set -u
TAG=${1:-r02am}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_published.py -q -x -p no:cacheprovider --timeout=120 > gpurun_out/pytest_pub_$TAG.log 2>&1; echo "published tests rc=$?"; tail -12 gpurun_out/pytest_pub_$TAG.log
timeout 300 python scripts/published_target.py --n 1048576 --only maf,iaf > gpurun_out/published_$TAG.jsonl 2> gpurun_out/published_$TAG.err; echo "rc=$?"
timeout 300 python scripts/published_target.py --n 4000 --reps 20 --only maf,iaf >> gpurun_out/published_$TAG.jsonl 2>> gpurun_out/published_$TAG.err; echo "rc=$?"
cat gpurun_out/published_$TAG.jsonl; tail -3 gpurun_out/published_$TAG.err
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_$TAG.log
