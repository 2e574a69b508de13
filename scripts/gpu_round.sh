#!/bin/bash
# Run on the B200 box: GPU tests, microbench, training-step benchmarks.  usage: scripts/gpu_round.sh <tag>
set -u
TAG=${1:-r01f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
timeout 900 python scripts/microbench.py --json gpurun_out/microbench_$TAG.json > gpurun_out/microbench_$TAG.log 2>&1; echo "microbench rc=$?"
grep -i "backward\|bwd\|col_sum\|wgrad\|compact" gpurun_out/microbench_$TAG.log
for m in realnvp256:65536 maf256:65536 spline784:4096 maf64:262144 realnvp2:1048576 spline2:1048576; do
  M=${m%%:*}; B=${m##*:}
  timeout 300 python scripts/train_step_bench.py --model $M --batch $B --steps 5 > gpurun_out/train_${TAG}_${M}.json 2> gpurun_out/train_${TAG}_${M}.err; echo "train $M rc=$?"; tail -1 gpurun_out/train_${TAG}_${M}.json
done
