#!/bin/bash
# Run on a 2-GPU B200 box (gpurun --gpus 2): weak-scaling bench at N=1,2 and the data-parallel training step.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "n1 rc=$?"
$TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err; echo "n2 rc=$?"
$TR --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/scale_ref_n2.json 2> gpurun_out/scale_ref_n2.err; echo "ref n2 rc=$?"
for m in realnvp256:65536 maf256:65536 spline784:4096 maf64:262144 realnvp2:1048576 spline2:1048576; do
  M=${m%%:*}; B=${m##*:}
  timeout 300 python scripts/train_step_bench.py --model $M --batch $B --steps 5 > gpurun_out/train_${M}_n1.json 2> gpurun_out/train_${M}_n1.err; echo "train $M n1 rc=$?"
done
for m in realnvp256:65536 maf256:65536 spline784:4096; do
  M=${m%%:*}; B=${m##*:}
  timeout 300 $TR --master-port 29513 scripts/train_step_bench.py --model $M --batch $B --steps 5 > gpurun_out/train_${M}_n2.json 2> gpurun_out/train_${M}_n2.err; echo "train $M n2 rc=$?"
done
tail -n 2 gpurun_out/scale_n*.json gpurun_out/train_*.json
