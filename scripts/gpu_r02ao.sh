#!/bin/bash
# pass ao: bench.py under torchrun at N = 2 exactly as the driver's scaling run launches it (the `also` blocks included),
# to check the multi-rank path of the default line and the end-to-end efficiency.  usage: <tag> <N>
set -u
TAG=${1:-r02ao}
N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err; echo "bench n=$N rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/scale_${TAG}_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step_runs'], d.get('host_affinity'))
print({k: (v.get('value'), v.get('e2e', {}).get('value'), v.get('unavailable')) for k, v in d.get('also', {}).items()})
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 \
    bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/scale_ref_${TAG}_n$N.json 2> gpurun_out/scale_ref_${TAG}_n$N.err; echo "ref n=$N rc=$?"; cut -c1-400 gpurun_out/scale_ref_${TAG}_n$N.json
