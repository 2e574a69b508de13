#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/wgrad_one.py 262144 512 512 > gpurun_out/plain_r01g_wgrad.log 2>&1 || { echo plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 3 -c 1 -o gpurun_out/prof_r01g_wgrad -f python scripts/wgrad_one.py 262144 512 512 > gpurun_out/ncu_full_r01g_wgrad.log 2>&1; echo "wgrad capture rc=$?"
python scripts/microbench.py --only spline_tf > gpurun_out/plain_r01g_stf.log 2>&1 || { echo plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:spline_transform_compact_bwd -s 3 -c 1 -o gpurun_out/prof_r01g_stbwd -f python scripts/microbench.py --only spline_tf > gpurun_out/ncu_full_r01g_stbwd.log 2>&1; echo "spline bwd capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spline_transform_compact_fwd -s 3 -c 1 -o gpurun_out/prof_r01g_stfwd -f python scripts/microbench.py --only spline_tf > gpurun_out/ncu_full_r01g_stfwd.log 2>&1; echo "spline fwd capture rc=$?"
