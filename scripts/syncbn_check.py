"""N-GPU == 1-GPU for a train-mode RealNVP step with synchronised BatchNorm statistics.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 scripts/syncbn_check.py

Every rank builds the same model and the same global batch, trains ONE step on its row shard through
DataParallelFlow (sync_batchnorm on / off), and rank 0 compares the averaged gradients, the loss and the BatchNorm
running statistics with a single-process step on the whole batch.  Prints one JSON line; exit code 1 on mismatch."""
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import nfb200 as N  # noqa: E402
from nfb200 import parallel as P  # noqa: E402


def nll(model, x):
    z, ld = model.inverse(x)
    return -N.ops.std_normal_log_prob(z, ld).sum()        # sum (not mean): shard losses add up to the global loss


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0")) % max(1, torch.cuda.device_count())
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # NF_DIST_BACKEND=gloo: several ranks on ONE GPU (gloo copies CUDA tensors through the host), for 1-GPU boxes
    backend = os.environ.get("NF_DIST_BACKEND", "nccl")
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group(backend)
    res = {}
    for D, H, B in ((2, 64, 10001), (16, 128, 4096)):
        torch.manual_seed(0)
        base = N.RealNVP(D, 4, H)
        with torch.no_grad():
            for p in base.parameters():
                p.add_(0.02 * torch.randn_like(p))
        x = torch.randn(B, D, generator=torch.Generator().manual_seed(1)).to(dev)
        # single-process step on the whole batch
        ref = copy.deepcopy(base).to(dev).train()
        P.enable_sync_batchnorm(False)
        nll(ref, x).backward()
        for sync in (True, False):
            m = copy.deepcopy(base).to(dev).train()
            dp = P.DataParallelFlow(m, sync_batchnorm=sync, overlap=False)
            lo, hi = P.shard_bounds(B, rank, world)
            loss = nll(dp, x[lo:hi].contiguous())
            loss.backward()
            dp.sync_gradients()                              # averages: multiply back by world to compare sums
            # relative to the largest gradient entry of the whole model: the biases of the Linears that feed a BatchNorm
            # have an exactly-zero true gradient, so a per-parameter relative error would compare rounding noise
            gmax = max(float(q.grad.abs().max()) for q in ref.parameters())
            gerr = max(float((p.grad * world - q.grad).abs().max()) for p, q in zip(m.parameters(), ref.parameters())) / gmax
            serr = max(float((a - b).abs().max()) for (n, a), (_, b) in zip(m.named_buffers(), ref.named_buffers())
                       if "running" in n)
            res[f"D{D}_H{H}_sync{int(sync)}"] = {"grad_rel_err": gerr, "running_stat_err": serr}
        P.enable_sync_batchnorm(False)
    ok = all(v["grad_rel_err"] < 1e-4 and v["running_stat_err"] < 1e-5 for k, v in res.items() if k.endswith("sync1"))
    differs = all(v["grad_rel_err"] > 1e-4 for k, v in res.items() if k.endswith("sync0"))
    if rank == 0:
        print(json.dumps({"world": world, "synchronised_equals_single_gpu": ok, "unsynchronised_differs": differs, "cases": res}))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
