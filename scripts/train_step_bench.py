#!/usr/bin/env python
"""Training-step throughput of the layered (autograd) route, optionally data-parallel over NCCL.

    python scripts/train_step_bench.py --model realnvp256 --batch 65536 --steps 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        scripts/train_step_bench.py --model realnvp256 --batch 65536 --steps 5

One step = inverse -> NLL (standard-normal head kernel) -> backward -> gradient allreduce (N > 1) -> Adam.
--batch is per GPU (weak scaling).  Prints one JSON line on rank 0; with N > 1 it also checks that every rank holds
identical parameters after the run (the allreduce really synchronised them).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import nfb200 as N  # noqa: E402

MODELS = {
    "realnvp2": lambda: (N.RealNVP(2, 8, 64), 2),                                   # C1
    "realnvp256": lambda: (N.RealNVP(256, 8, 512), 256),                            # C5a
    "maf256": lambda: (N.NormalizingFlowModel([N.MaskedAutoregressiveFlow(256, 1024) for _ in range(4)]), 256),   # C5b
    "maf64": lambda: (N.MaskedAutoregressiveFlow(64, 512), 64),                     # C3 (training direction)
    "spline784": lambda: (N.RealNVPSpline(784, 16, 1024), 784),                     # C4
    "spline2": lambda: (N.NormalizingFlowModel([N.SplineCouplingLayer(2, 64, m, num_bins=8) for m in
                                                [torch.tensor([1., 0.]) if i % 2 == 0 else torch.tensor([0., 1.]) for i in range(8)]]), 2),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="realnvp256", choices=list(MODELS))
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--graph", action="store_true", help="capture the whole step (fwd, bwd, Adam) in one CUDA graph (1 GPU)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32", "bf16"],
                    help="dense-layer precision: fp32 = 3xTF32 (reference tolerances), tf32 = one TF32 pass (reduced-precision mode)")
    ap.add_argument("--micro", type=int, default=1,
                    help="micro-batches of --batch rows per optimizer step; all but the last run under no_sync() (gradient accumulation)")
    ap.add_argument("--no-overlap", action="store_true", help="all-reduce after backward instead of during it (A/B for the overlap)")
    ap.add_argument("--no-sync-bn", action="store_true", help="per-shard BatchNorm statistics (what DistributedDataParallel would do)")
    a = ap.parse_args()
    N.set_gemm_precision(a.precision)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234 + rank)            # different init per rank: the wrapper must broadcast rank 0's weights
    model, D = MODELS[a.model]()
    model.to(dev).train()
    dp = N.parallel.DataParallelFlow(model, overlap=not a.no_overlap, sync_batchnorm=not a.no_sync_bn)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True, capturable=a.graph)
    gen = torch.Generator(device=dev).manual_seed(rank)
    xs = [torch.randn(a.batch, D, device=dev, generator=gen) for _ in range(min(a.micro, 4))]
    x = xs[0]

    def micro_step(xb):
        z, ld = dp.inverse(xb)
        loss = -N.ops.std_normal_log_prob(z, ld).mean() / a.micro
        loss.backward()
        return loss

    def step():
        opt.zero_grad(set_to_none=True)
        if a.micro > 1:
            with dp.no_sync():
                for i in range(a.micro - 1):
                    micro_step(xs[i % len(xs)])
        loss = micro_step(xs[(a.micro - 1) % len(xs)])
        dp.sync_gradients()
        opt.step()
        return loss

    if a.graph:
        # launch-bound small batches (C1: 5000 rows, ~500 launches per step): replay the step as one CUDA graph.
        # Every kernel of the step goes through the C ABI on torch's current stream, so stream capture records them.
        assert world == 1

        def loss_fn(mdl, xin):
            z, ld = mdl.inverse(xin)
            return -N.ops.std_normal_log_prob(z, ld).mean()
        graphed = N.graphs.GraphedTrainStep(model, opt, loss_fn, x)

        def step():        # noqa: F811
            return graphed()
    for _ in range(a.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = N._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    in_sync = True
    if world > 1:
        for p in model.parameters():
            ref = p.detach().clone()
            dist.broadcast(ref, src=0)
            in_sync &= bool(torch.equal(ref, p.detach()))
        flag = torch.tensor([int(in_sync)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        in_sync = bool(flag.item())
    if rank == 0:
        nparam = sum(p.numel() for p in model.parameters())
        print(json.dumps({"model": a.model, "n_gpus": world, "batch_per_gpu": a.batch, "micro_batches": a.micro,
                          "global_rows_per_step": a.batch * a.micro * world, "ms_per_step": ms.item(),
                          "samples_per_s": a.batch * a.micro * world / (ms.item() * 1e-3), "loss": float(loss),
                          "overlap": not a.no_overlap, "sync_batchnorm": not a.no_sync_bn,
                          "params": nparam, "allreduce_bytes_per_step": nparam * 4 if world > 1 else 0,
                          "cuda_graph": bool(a.graph), "gemm_precision": a.precision, "replicas_in_sync": in_sync, "launches_per_step": (N._lib.launch_count() - l0) / a.steps,
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
