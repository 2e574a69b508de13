#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/gemm_one.py 262144 512 512 > gpurun_out/plain_r01r_gemm2.log 2>&1 || { echo plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 3 -c 1 -o gpurun_out/prof_r01r_gemm2 -f python scripts/gemm_one.py 262144 512 512 > gpurun_out/ncu_full_r01r_gemm2.log 2>&1; echo "gemm2 capture rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_r01r_c2.log 2>&1 || { echo plain c2 failed; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01r_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch_r01r_c2.log 2>&1; echo "c2 launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spline_stack_tc -s 4 -c 2 -o gpurun_out/prof_r01r_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_r01r_c2.log 2>&1; echo "c2 full rc=$?"
python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_r01r_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01r_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch_r01r_c3.log 2>&1; echo "c3 launch list rc=$?"
