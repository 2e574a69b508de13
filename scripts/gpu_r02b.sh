#!/bin/bash
# Round 2, second GPU pass: bf16 chain tests (short timeout), chain bench, whole GPU suite, then ONE ncu capture of the chain
# kernel (launch list of chain_bench + full set of made_chain_bf16_kernel).  usage: gpu_r02b.sh <tag>
set -u
TAG=${1:-r02b}
mkdir -p gpurun_out
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 300 python -m pytest tests/test_gpu_bf16.py -q -s -p no:cacheprovider > gpurun_out/pytest_bf16_$TAG.log 2>&1; BF=$?; echo "bf16 pytest rc=$BF"; grep "^\[bf16\]" gpurun_out/pytest_bf16_$TAG.log | head -40; tail -12 gpurun_out/pytest_bf16_$TAG.log
timeout 200 python scripts/chain_bench.py > gpurun_out/chain_bench_$TAG.jsonl 2>&1; CB=$?; echo "chain bench rc=$CB"; cat gpurun_out/chain_bench_$TAG.jsonl
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 --deselect tests/test_gpu_bf16.py > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_$TAG.log
if [ $CB -eq 0 ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:made_chain_bf16 -s 3 -c 1 -o gpurun_out/prof_${TAG}_chain -f python scripts/chain_bench.py 262144 4 > gpurun_out/ncu_full_${TAG}_chain.log 2>&1; echo "chain full rc=$?"
fi
