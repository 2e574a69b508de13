#!/bin/bash
# pass aj: wide skinny_k variant + 4-row skinny_reduce: tests, microbench, train steps.  usage: <tag>
set -u
TAG=${1:-r02aj}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_gradients.py -q -x -p no:cacheprovider --timeout=120 > gpurun_out/pytest_skinny_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/pytest_skinny_$TAG.log
timeout 300 python scripts/microbench.py --only skinny > gpurun_out/microbench_skinny_$TAG.log 2>&1; echo "rc=$?"; cat gpurun_out/microbench_skinny_$TAG.log
for m in realnvp2 spline2; do timeout 300 python scripts/train_step_bench.py --model $m --batch 1048576 --steps 3 2>/dev/null | cut -c1-200; done
