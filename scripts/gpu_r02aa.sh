#!/bin/bash
# pass aa: stream-kernel grids sized by occupancy (microbench), smoke(), reference arm sanity, harness report.  usage: <tag>
set -u
TAG=${1:-r02aa}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spline_stream.py -q -x -p no:cacheprovider --timeout=120 > gpurun_out/pytest_stream_$TAG.log 2>&1; echo "stream tests rc=$?"; tail -2 gpurun_out/pytest_stream_$TAG.log
timeout 300 python scripts/microbench.py --only rqs,spline_tf > gpurun_out/microbench_$TAG.log 2>&1; echo "microbench rc=$?"; grep -v "full " gpurun_out/microbench_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref arm rc=$?"; cut -c1-400 gpurun_out/bench_ref_$TAG.json
timeout 600 python scripts/reference_harness_report.py > gpurun_out/reference_harness_$TAG.jsonl 2> gpurun_out/reference_harness_$TAG.err; echo "harness rc=$?"; cat gpurun_out/reference_harness_$TAG.jsonl | cut -c1-700
