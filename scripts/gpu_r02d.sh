#!/bin/bash
# Round 2, fourth GPU pass: where the chain kernel waits (per-role counters, debug library), bf16 tests incl. the SS-form dense
# layers, train steps in the three precision modes, then ONE ncu capture (chain kernel, full set + source).  usage: gpu_r02d.sh <tag>
set -u
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 200 python scripts/chain_phase_profile.py > gpurun_out/chain_phase_$TAG.log 2>&1; echo "phase profile rc=$?"; cat gpurun_out/chain_phase_$TAG.log
timeout 300 python -m pytest tests/test_gpu_bf16.py -q -s -p no:cacheprovider > gpurun_out/pytest_bf16_$TAG.log 2>&1; echo "bf16 pytest rc=$?"; grep "^\[bf16\] \(realnvp\|spline\)" gpurun_out/pytest_bf16_$TAG.log; tail -3 gpurun_out/pytest_bf16_$TAG.log
timeout 200 python scripts/chain_bench.py > gpurun_out/chain_bench_$TAG.jsonl 2>&1; echo "chain bench rc=$?"; grep bf16 gpurun_out/chain_bench_$TAG.jsonl
for m in spline784:4096 realnvp256:65536 maf256:65536 maf64:262144; do
  M=${m%%:*}; B=${m##*:}
  for prec in fp32 tf32 bf16; do
    timeout 300 python scripts/train_step_bench.py --model $M --batch $B --steps 5 --precision $prec 2>/dev/null | tail -1 >> gpurun_out/train_$TAG.jsonl
  done
done
cat gpurun_out/train_$TAG.jsonl | cut -c1-260
timeout 600 ncu --set full --clock-control none --import-source on -k regex:made_chain_bf16 -s 3 -c 1 -o gpurun_out/prof_${TAG}_chain -f python scripts/chain_bench.py 262144 4 > gpurun_out/ncu_full_${TAG}_chain.log 2>&1; echo "chain full rc=$?"
