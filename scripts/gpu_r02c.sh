#!/bin/bash
# Round 2, third GPU pass: chain kernel after the epilogue fix (tests + bench), HBM microbench after the vectorised kernels,
# full suite, then ONE compute-sanitizer tool (argument 2: memcheck | racecheck | synccheck).  usage: gpu_r02c.sh <tag> <tool>
set -u
TAG=${1:-r02c}
TOOL=${2:-memcheck}
mkdir -p gpurun_out
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 300 python -m pytest tests/test_gpu_bf16.py -q -s -p no:cacheprovider > gpurun_out/pytest_bf16_$TAG.log 2>&1; BF=$?; echo "bf16 pytest rc=$BF"; tail -4 gpurun_out/pytest_bf16_$TAG.log
timeout 200 python scripts/chain_bench.py > gpurun_out/chain_bench_$TAG.jsonl 2>&1; echo "chain bench rc=$?"; cat gpurun_out/chain_bench_$TAG.jsonl
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 --deselect tests/test_gpu_bf16.py > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_$TAG.log
timeout 600 python scripts/microbench.py > gpurun_out/microbench_$TAG.log 2>&1; echo "microbench rc=$?"; grep -E "std_normal|feature_affine|col_stats" gpurun_out/microbench_$TAG.log
timeout 120 python scripts/sanitizer_target.py > gpurun_out/sanitizer_plain_$TAG.log 2>&1; SP=$?; echo "sanitizer target plain rc=$SP"; tail -2 gpurun_out/sanitizer_plain_$TAG.log
if [ $SP -eq 0 ]; then
timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python scripts/sanitizer_target.py > gpurun_out/sanitizer_${TOOL}_$TAG.log 2>&1; echo "$TOOL rc=$?"; tail -8 gpurun_out/sanitizer_${TOOL}_$TAG.log
fi
