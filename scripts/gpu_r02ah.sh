#!/bin/bash
# pass ah: 128-bit skinny kernels: tests, microbench rows, published / notebook timings, full suite.  usage: <tag>
set -u
TAG=${1:-r02ah}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensorcore.py -q -x -k skinny -p no:cacheprovider --timeout=120 > gpurun_out/pytest_skinny_$TAG.log 2>&1; echo "skinny tests rc=$?"; tail -4 gpurun_out/pytest_skinny_$TAG.log
timeout 300 python scripts/microbench.py --only skinny > gpurun_out/microbench_skinny_$TAG.log 2>&1; echo "rc=$?"; cat gpurun_out/microbench_skinny_$TAG.log
timeout 600 python scripts/published_target.py --n 1048576 --reps 2 --only realnvp,maf,iaf,nb_realnvp256,nb_maf128,nb_iaf128,nb_spline128 > gpurun_out/published_$TAG.jsonl 2> gpurun_out/published_$TAG.err; echo "rc=$?"; cat gpurun_out/published_$TAG.jsonl
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
