#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -x -q -k "wgrad" > gpurun_out/pytest_wgrad.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_wgrad.log
timeout 300 python scripts/wgrad_bench.py > gpurun_out/wgrad_bench.log 2>&1; echo "bench rc=$?"; cat gpurun_out/wgrad_bench.log
