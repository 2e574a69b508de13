#!/bin/bash
# Run on the B200 box: full GPU test suite, training-step benchmarks, ncu launch lists of two training steps.
set -u
TAG=${1:-r01e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
for m in realnvp256:65536 maf256:65536 spline784:4096 maf64:262144 realnvp2:1048576 spline2:1048576; do
  M=${m%%:*}; B=${m##*:}
  timeout 300 python scripts/train_step_bench.py --model $M --batch $B --steps 5 > gpurun_out/train_${TAG}_${M}.json 2> gpurun_out/train_${TAG}_${M}.err; echo "train $M rc=$?"; tail -1 gpurun_out/train_${TAG}_${M}.json
done
for m in realnvp256:65536 maf256:65536 spline784:4096; do
  M=${m%%:*}; B=${m##*:}
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_train_${TAG}_${M}.csv python scripts/train_step_bench.py --model $M --batch $B --steps 1 --warmup 1 > gpurun_out/ncu_train_${TAG}_${M}.log 2>&1; echo "ncu $M rc=$?"
done
