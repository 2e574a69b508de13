#!/bin/bash
# pass w: large-batch routes of the published configs (hidden 128 coupling / spline stacks, 2-D MADE chains):
# new tests, full suite, published-config timings, bench line.  usage: gpu_r02w.sh <tag>
set -u
TAG=${1:-r02w}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_published.py -q -p no:cacheprovider --timeout=120 -s > gpurun_out/pytest_pub_$TAG.log 2>&1; echo "published tests rc=$?"; grep "published\]\|passed\|failed\|Error" gpurun_out/pytest_pub_$TAG.log | tail -30
timeout 300 python scripts/published_target.py --n 1048576 > gpurun_out/published_$TAG.jsonl 2> gpurun_out/published_$TAG.err; echo "rc=$?"
timeout 300 python scripts/published_target.py --n 4000 --reps 20 >> gpurun_out/published_$TAG.jsonl 2>> gpurun_out/published_$TAG.err; echo "rc=$?"
cat gpurun_out/published_$TAG.jsonl
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
