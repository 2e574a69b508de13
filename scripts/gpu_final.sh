#!/bin/bash
# Run on the B200 box: full GPU test suite, smoke, both bench workloads, reference arm, training steps.  usage: gpu_final.sh <tag>
set -u
TAG=${1:-r01m}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo "bench c2 rc=$?"; tail -1 gpurun_out/bench_${TAG}_c2.json | cut -c1-400
timeout 900 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_c3.json 2> gpurun_out/bench_${TAG}_c3.err; echo "bench c3 rc=$?"; tail -1 gpurun_out/bench_${TAG}_c3.json | cut -c1-300
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "bench ref rc=$?"; tail -1 gpurun_out/bench_${TAG}_ref.json | cut -c1-300
for m in realnvp256:65536 maf256:65536 spline784:4096 maf64:262144 realnvp2:1048576 spline2:1048576; do
  M=${m%%:*}; B=${m##*:}
  timeout 300 python scripts/train_step_bench.py --model $M --batch $B --steps 5 2>/dev/null | tail -1 > gpurun_out/train_${TAG}_${M}.json; echo "train $M rc=$?"; cat gpurun_out/train_${TAG}_${M}.json | cut -c1-200
done
for m in spline784:4096 realnvp256:65536 maf256:65536 maf64:262144; do
  M=${m%%:*}; B=${m##*:}
  timeout 300 python scripts/train_step_bench.py --model $M --batch $B --steps 5 --precision tf32 2>/dev/null | tail -1 > gpurun_out/train_${TAG}_${M}_tf32.json; echo "train $M tf32 rc=$?"; cat gpurun_out/train_${TAG}_${M}_tf32.json | cut -c1-200
done
timeout 300 python scripts/microbench.py > gpurun_out/microbench_$TAG.log 2>&1; echo "microbench rc=$?"
for m in realnvp2 spline2 maf64; do
  for g in "" "--graph"; do
    timeout 300 python scripts/train_step_bench.py --model $m --batch 5000 --steps 50 --warmup 5 $g 2>/dev/null | tail -1 > gpurun_out/c1_${TAG}_${m}${g}.json; cat gpurun_out/c1_${TAG}_${m}${g}.json | cut -c1-200
  done
done
