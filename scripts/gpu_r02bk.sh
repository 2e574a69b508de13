#!/bin/bash
# cold-cache DRAM traffic of the C3 log_prob chain (for bench.py's also.c3 roofline.traffic): ncu --set full --cache-control all of the
# third log_prob pass of scripts/c3_logprob_target.py
set -u
TAG=${1:-r02bk}
mkdir -p gpurun_out
timeout 120 python scripts/c3_logprob_target.py > gpurun_out/c3_target_$TAG.log 2>&1; echo "plain rc=$?"; cat gpurun_out/c3_target_$TAG.log
timeout 600 ncu --set full --cache-control all --clock-control none -k regex:"gemm_tc2_kernel|gemm_tc_kernel|affine_ar|std_normal" -s 10 -c 5 -f -o gpurun_out/${TAG}_c3_chain python scripts/c3_logprob_target.py > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log
