#!/bin/bash
# Round 2, pass j: packed-fp32 layer 1 / spline knots everywhere (whole suite = parity of every spline kernel), default bench,
# HBM microbench, chain bench, ONE ncu capture (C2 kernel).  usage: gpu_r02j.sh <tag>
set -u
TAG=${1:-r02j}
mkdir -p gpurun_out
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo "bench c2 rc=$?"; tail -1 gpurun_out/bench_${TAG}_c2.json | cut -c1-300; tail -3 gpurun_out/bench_${TAG}_c2.err
timeout 600 python scripts/microbench.py > gpurun_out/microbench_$TAG.log 2>&1; echo "microbench rc=$?"; grep -E "spline_transform compact|rqs_unit" gpurun_out/microbench_$TAG.log | head -12
timeout 200 python scripts/chain_bench.py > gpurun_out/chain_bench_$TAG.jsonl 2>&1; echo "chain bench rc=$?"; grep bf16 gpurun_out/chain_bench_$TAG.jsonl
python bench.py --steps 2 --warmup 3 --no-cpu --no-also > gpurun_out/plain_${TAG}_c2.log 2>&1 &&
timeout 600 ncu --set full --cache-control all --clock-control none --import-source on -k regex:spline_stack_tc -s 4 -c 2 -o gpurun_out/prof_${TAG}_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-also > gpurun_out/ncu_full_${TAG}_c2.log 2>&1; echo "c2 full rc=$?"
