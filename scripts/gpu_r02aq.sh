#!/bin/bash
# pass aq: in-block steps of the blocked sequential direction on the tensor cores (mma.sync variant, nf_set_option(3, 2))
# against the FP32-pipe variant: parity tests of the route under both, C3 sampler timing.  usage: <tag>
set -u
TAG=${1:-r02aq}
mkdir -p gpurun_out
for v in 1 2; do
  timeout 120 python scripts/sampler_target.py --variant $v --reps 10 > gpurun_out/sampler_${TAG}_v$v.json 2> gpurun_out/sampler_${TAG}_v$v.err; echo "sampler v$v rc=$?"; cat gpurun_out/sampler_${TAG}_v$v.json; tail -3 gpurun_out/sampler_${TAG}_v$v.err
done
NFB200_OPTIONS=3:2 timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 -k "blocked or c3 or sequential" > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_${TAG}.log
