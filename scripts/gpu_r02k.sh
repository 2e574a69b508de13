#!/bin/bash
# Round 2, pass k: packed-fp32 FMAs in the FP32-pipe stack kernels, the sampler's in-block kernel and the FP32-pipe GEMM:
# whole suite, default bench (C2, C3 incl. the sampler), reference-harness report, microbench.  usage: gpu_r02k.sh <tag>
set -u
TAG=${1:-r02k}
mkdir -p gpurun_out
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo "bench c2 rc=$?"; tail -1 gpurun_out/bench_${TAG}_c2.json | cut -c1-300; tail -3 gpurun_out/bench_${TAG}_c2.err
timeout 600 python scripts/reference_harness_report.py > gpurun_out/${TAG}_reference_harness.jsonl 2> gpurun_out/${TAG}_reference_harness.err; echo "harness rc=$?"; cat gpurun_out/${TAG}_reference_harness.jsonl | cut -c1-330
timeout 600 python scripts/microbench.py > gpurun_out/microbench_$TAG.log 2>&1; echo "microbench rc=$?"; grep -E "nf_gemm fp32" gpurun_out/microbench_$TAG.log | head -6
timeout 300 python scripts/train_step_bench.py --model realnvp2 --batch 5000 --steps 50 --warmup 5 --graph 2>/dev/null | tail -1 | cut -c1-250
