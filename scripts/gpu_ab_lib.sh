#!/bin/bash
# A/B of a variant library (scripts/build_variant_lib.py) against the shipped one on the bench line, alternating runs.
# usage: gpu_ab_lib.sh <tag> <variant>
set -u
TAG=${1:-r02z}; VAR=${2:-hint}
mkdir -p gpurun_out
for rep in 1 2; do
  for which in base $VAR; do
    if [ $which = base ]; then unset NFB200_LIB; else export NFB200_LIB=$PWD/normalizing-flows-study_b200/lib/libnfb200_$VAR.so; fi
    timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/ab_${TAG}_${which}_$rep.json 2> gpurun_out/ab_${TAG}_${which}_$rep.err; echo "$which $rep rc=$?"
    python - <<PY
import json
d=json.loads(open('gpurun_out/ab_${TAG}_${which}_$rep.json').read().strip().splitlines()[-1])
a=d.get('also',{})
print('$which', $rep, 'c2 ms', d['roofline']['ms_per_launch'], 'e2e', round(d['e2e']['value']/1e9,3), {k:v.get('ms_per_launch') for k,v in a.items()})
PY
  done
done
