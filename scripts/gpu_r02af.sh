#!/bin/bash
# pass af: full suite, full microbench, default bench line (both arms), for the round-end numbers.  usage: <tag>
set -u
TAG=${1:-r02af}
mkdir -p gpurun_out
rm -f gpurun_out/parity_fallbacks.jsonl gpurun_out/parity_fullsize.jsonl
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=300 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 900 python scripts/microbench.py > gpurun_out/microbench_$TAG.log 2>&1; echo "microbench rc=$?"; grep -c . gpurun_out/microbench_$TAG.log; grep "col_sum\|col_stats\|std_normal" gpurun_out/microbench_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step_runs'], d['roofline']['frac'], d['cpu_baseline'], d.get('eager_cuda',{}).get('value'), {k:(v.get('value'), v.get('e2e',{}).get('value')) for k,v in d.get('also',{}).items()})"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_$TAG.json
