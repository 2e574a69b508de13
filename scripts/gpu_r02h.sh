#!/bin/bash
# Round 2, pass h: C2 stack kernel with packed fp32 split / bias adds (bench + full-size parity), whole suite, synchronised
# BatchNorm equivalence with 2 ranks on one GPU over gloo, default bench, ONE ncu capture (C2 kernel).  usage: gpu_r02h.sh <tag>
set -u
TAG=${1:-r02h}
mkdir -p gpurun_out
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_$TAG.log
NF_DIST_BACKEND=gloo timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 scripts/syncbn_check.py > gpurun_out/syncbn_check_$TAG.json 2> gpurun_out/syncbn_check_$TAG.err; echo "syncbn (2 ranks, gloo, one GPU) rc=$?"; tail -1 gpurun_out/syncbn_check_$TAG.json
timeout 900 python bench.py > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo "bench c2 rc=$?"; tail -1 gpurun_out/bench_${TAG}_c2.json | cut -c1-300; tail -3 gpurun_out/bench_${TAG}_c2.err
python bench.py --steps 2 --warmup 3 --no-cpu --no-also > gpurun_out/plain_${TAG}_c2.log 2>&1 &&
timeout 600 ncu --set full --cache-control all --clock-control none --import-source on -k regex:spline_stack_tc -s 4 -c 2 -o gpurun_out/prof_${TAG}_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-also > gpurun_out/ncu_full_${TAG}_c2.log 2>&1; echo "c2 full rc=$?"
