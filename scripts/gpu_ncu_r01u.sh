#!/bin/bash
# ncu --set full of the short-chain dense-layer kernel in both precision modes (after a plain run of the same command)
set -u
mkdir -p gpurun_out
for p in fp32 tf32; do
  python scripts/gemm_one.py 262144 512 512 $p > gpurun_out/plain_r01u_gemm2_$p.log 2>&1 || { echo plain $p failed; exit 1; }
  ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 3 -c 1 -o gpurun_out/prof_r01u_gemm2_$p -f python scripts/gemm_one.py 262144 512 512 $p > gpurun_out/ncu_full_r01u_gemm2_$p.log 2>&1; echo "gemm2 $p capture rc=$?"
done
