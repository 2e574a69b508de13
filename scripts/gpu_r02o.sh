#!/bin/bash
# Round 2, pass o: in-block kernel without the layer-3 tile (8 warps per SM), per-degree padded layout, 128-bit epilogue
# stores: tests that touch the autoregressive flows + full-size parity, sampler timing, launch list.  usage: gpu_r02o.sh <tag>
set -u
TAG=${1:-r02o}
mkdir -p gpurun_out
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_$TAG.log
timeout 300 python scripts/sampler_target.py > gpurun_out/sampler_$TAG.json 2> gpurun_out/sampler_$TAG.err; echo "sampler rc=$?"; cat gpurun_out/sampler_$TAG.json; tail -3 gpurun_out/sampler_$TAG.err
timeout 300 python scripts/sampler_target.py --precision bf16 >> gpurun_out/sampler_$TAG.json 2>> gpurun_out/sampler_$TAG.err; echo "sampler bf16 rc=$?"; tail -1 gpurun_out/sampler_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_c3_sampler_launches.csv python scripts/sampler_target.py --reps 1 > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
