#!/bin/bash
# Run on the B200 box: reduced-precision (one TF32 pass) mode -- tests, GEMM microbench, training steps in both modes.  usage: gpu_tf32.sh <tag>
set -u
TAG=${1:-r01t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_$TAG.log
timeout 300 python scripts/tf32_mode_accuracy.py > gpurun_out/tf32_mode_accuracy_$TAG.log 2>&1; cut -c1-300 gpurun_out/tf32_mode_accuracy_$TAG.log
timeout 300 python scripts/microbench.py --only gemm > gpurun_out/microbench_gemm_$TAG.log 2>&1; echo "microbench rc=$?"; cat gpurun_out/microbench_gemm_$TAG.log | cut -c1-200
for m in spline784:4096 realnvp256:65536 maf256:65536 maf64:262144; do
  M=${m%%:*}; B=${m##*:}
  for p in fp32 tf32; do
    timeout 300 python scripts/train_step_bench.py --model $M --batch $B --steps 5 --precision $p 2>gpurun_out/train_${TAG}_${M}_$p.err | tail -1 > gpurun_out/train_${TAG}_${M}_$p.json; echo "train $M $p rc=$?"; cut -c1-260 gpurun_out/train_${TAG}_${M}_$p.json
  done
done
