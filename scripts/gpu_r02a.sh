#!/bin/bash
# Round 2, first GPU pass: new bf16 chain kernel first (short timeout: a deadlock must not eat the call), whole GPU test suite
# (no -x), smoke, default bench (c2 + c3 block + eager/cpu legs), reference arm, reference-harness report, microbench, then
# ONE ncu capture (cold-cache full set of the C2 kernel).  usage: gpu_r02a.sh <tag>
set -u
TAG=${1:-r02a}
mkdir -p gpurun_out
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 300 python -m pytest tests/test_gpu_bf16.py -x -q -s -p no:cacheprovider > gpurun_out/pytest_bf16_$TAG.log 2>&1; BF=$?; echo "bf16 pytest rc=$BF"; tail -25 gpurun_out/pytest_bf16_$TAG.log
if [ $BF -ne 0 ]; then export NF_SKIP_BF16=1; fi
if [ $BF -eq 0 ]; then timeout 200 python scripts/chain_bench.py > gpurun_out/chain_bench_$TAG.jsonl 2>&1; echo "chain bench rc=$?"; cat gpurun_out/chain_bench_$TAG.jsonl; fi
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 --deselect tests/test_gpu_bf16.py > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo "bench c2 rc=$?"; tail -1 gpurun_out/bench_${TAG}_c2.json | cut -c1-600; tail -5 gpurun_out/bench_${TAG}_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "bench ref rc=$?"; tail -1 gpurun_out/bench_${TAG}_ref.json | cut -c1-300
timeout 600 python scripts/reference_harness_report.py > gpurun_out/${TAG}_reference_harness.jsonl 2> gpurun_out/${TAG}_reference_harness.err; echo "harness rc=$?"; cat gpurun_out/${TAG}_reference_harness.jsonl | cut -c1-400; tail -3 gpurun_out/${TAG}_reference_harness.err
timeout 600 python scripts/microbench.py > gpurun_out/microbench_$TAG.log 2>&1; echo "microbench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-also > gpurun_out/plain_${TAG}_c2.log 2>&1 &&
timeout 600 ncu --set full --cache-control all --clock-control none --import-source on -k regex:spline_stack_tc -s 4 -c 2 -o gpurun_out/prof_${TAG}_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-also > gpurun_out/ncu_full_${TAG}_c2.log 2>&1; echo "c2 full rc=$?"
