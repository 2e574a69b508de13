#!/bin/bash
# pass ad: 2-D sequential direction through the chain: tests, timings (published + notebook shapes), full suite.  usage: <tag>
set -u
TAG=${1:-r02ad}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_published.py -q -p no:cacheprovider --timeout=120 > gpurun_out/pytest_pub_$TAG.log 2>&1; echo "published tests rc=$?"; tail -12 gpurun_out/pytest_pub_$TAG.log
timeout 300 python scripts/published_target.py --n 1048576 > gpurun_out/published_$TAG.jsonl 2> gpurun_out/published_$TAG.err; echo "rc=$?"
timeout 300 python scripts/published_target.py --n 4000 --reps 20 >> gpurun_out/published_$TAG.jsonl 2>> gpurun_out/published_$TAG.err; echo "rc=$?"
timeout 600 python scripts/published_target.py --n 1048576 --reps 2 --only nb_realnvp256,nb_maf128,nb_iaf128,nb_spline128 >> gpurun_out/published_$TAG.jsonl 2>> gpurun_out/published_$TAG.err; echo "rc=$?"
cat gpurun_out/published_$TAG.jsonl
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
