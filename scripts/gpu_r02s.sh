#!/bin/bash
# pass s: shared scalar-math changes (packed exponentials, fast reciprocals, FMNMX.NaN) touch every float32 kernel:
# full GPU suite, spline / rqs microbench rows, the default bench line, ncu of the spline kernels.  usage: gpu_r02s.sh <tag>
set -u
TAG=${1:-r02s}
mkdir -p gpurun_out
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_$TAG.log
timeout 300 python scripts/microbench.py --only rqs,spline_tf > gpurun_out/microbench_$TAG.log 2>&1; echo "microbench rc=$?"; grep -v "full " gpurun_out/microbench_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'], {k:(v.get('value'), v.get('ms_per_pass', v.get('ms'))) for k,v in d.get('also',{}).items()})"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spline_stream" -c 9 -f -o gpurun_out/${TAG}_spline_tf python scripts/microbench.py --only spline_tf_ncu > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log
