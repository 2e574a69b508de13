#!/bin/bash
# chain kernel iteration: phase profile, bf16 tests, chain bench, microbench lines of the touched HBM kernels, whole suite,
# ONE ncu capture (chain kernel).  usage: gpu_r02g.sh <tag>
set -u
TAG=${1:-r02g}
mkdir -p gpurun_out
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 200 python scripts/chain_phase_profile.py > gpurun_out/chain_phase_$TAG.log 2>&1; echo "phase profile rc=$?"; cat gpurun_out/chain_phase_$TAG.log
timeout 300 python -m pytest tests/test_gpu_bf16.py -q -s -p no:cacheprovider > gpurun_out/pytest_bf16_$TAG.log 2>&1; echo "bf16 pytest rc=$?"; tail -3 gpurun_out/pytest_bf16_$TAG.log
timeout 200 python scripts/chain_bench.py > gpurun_out/chain_bench_$TAG.jsonl 2>&1; echo "chain bench rc=$?"; grep bf16 gpurun_out/chain_bench_$TAG.jsonl
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 --deselect tests/test_gpu_bf16.py > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python scripts/microbench.py > gpurun_out/microbench_$TAG.log 2>&1; echo "microbench rc=$?"; grep -E "std_normal|col_stats" gpurun_out/microbench_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:made_chain_bf16 -s 3 -c 1 -o gpurun_out/prof_${TAG}_chain -f python scripts/chain_bench.py 262144 4 > gpurun_out/ncu_full_${TAG}_chain.log 2>&1; echo "chain full rc=$?"
