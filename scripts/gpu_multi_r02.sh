#!/bin/bash
# Round 2, multi-GPU pass (gpurun --gpus N, N = 2, 4 or 8): data-parallel TRAINING steps of BASELINE config 5 at
# 1, 2, 4, ... N GPUs (weak scaling: 262144-row micro-batches, 8 per GPU and optimizer step = 16 M global rows at N = 8, under no_sync(), one
# bucketed NCCL gradient all-reduce overlapped with the last backward), the all-reduce overlap A/B, the synchronised-BatchNorm
# equivalence check, and the inference bench (no collective) at N.   usage: gpu_multi_r02.sh <tag> <N>
set -u
TAG=${1:-r02m}
NMAX=${2:-2}
mkdir -p gpurun_out
OUT=gpurun_out/train_scaling_$TAG.jsonl
: > $OUT
TR() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
P=29600
for n in 1 2 4 8; do
  if [ $n -gt $NMAX ]; then break; fi
  for m in realnvp256:262144 maf256:262144; do
    M=${m%%:*}; B=${m##*:}
    if [ $M = maf256 ] && [ $n -ne 1 ] && [ $n -ne $NMAX ]; then continue; fi
    P=$((P+1))
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P \
        scripts/train_step_bench.py --model $M --batch $B --micro 8 --steps 3 --warmup 1 2> gpurun_out/train_${TAG}_${M}_n$n.err | tail -1 >> $OUT
    echo "train $M n=$n rc=$?"
  done
done
# overlap A/B and per-shard-BatchNorm A/B at the largest N
for extra in "--no-overlap" "--no-sync-bn"; do
  P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port $P \
      scripts/train_step_bench.py --model realnvp256 --batch 65536 --micro 1 --steps 4 --warmup 2 $extra 2>> gpurun_out/train_${TAG}_ab.err | tail -1 >> $OUT
done
P=$((P+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port $P \
    scripts/train_step_bench.py --model realnvp256 --batch 65536 --micro 1 --steps 4 --warmup 2 2>> gpurun_out/train_${TAG}_ab.err | tail -1 >> $OUT
cat $OUT
P=$((P+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P scripts/syncbn_check.py > gpurun_out/syncbn_check_$TAG.json 2> gpurun_out/syncbn_check_$TAG.err; echo "syncbn n=2 rc=$?"; cat gpurun_out/syncbn_check_$TAG.json
if [ $NMAX -gt 2 ]; then
  P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port $P scripts/syncbn_check.py > gpurun_out/syncbn_check_${TAG}_n$NMAX.json 2>> gpurun_out/syncbn_check_$TAG.err; echo "syncbn n=$NMAX rc=$?"; cat gpurun_out/syncbn_check_${TAG}_n$NMAX.json
fi
# (the inference bench at 1, 2, 4, 8 GPUs is the driver's own scaling run at round end)
