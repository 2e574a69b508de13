#!/bin/bash
# GEMM microbench, tensor-core tests, then ncu --set full of the short-chain dense-layer kernel in both precision modes
set -u
TAG=${1:-r01w}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -x -q -m gpu > gpurun_out/pytest_tc_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_tc_$TAG.log
timeout 300 python scripts/microbench.py --only gemm > gpurun_out/microbench_gemm_$TAG.log 2>&1; echo "microbench rc=$?"; grep "linear_tc" gpurun_out/microbench_gemm_$TAG.log | cut -c1-200
for p in fp32 tf32; do
  python scripts/gemm_one.py 262144 512 512 $p > gpurun_out/plain_${TAG}_gemm2_$p.log 2>&1 || { echo plain $p failed; exit 1; }
  ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 3 -c 1 -o gpurun_out/prof_${TAG}_gemm2_$p -f python scripts/gemm_one.py 262144 512 512 $p > gpurun_out/ncu_full_${TAG}_gemm2_$p.log 2>&1; echo "gemm2 $p capture rc=$?"
done
