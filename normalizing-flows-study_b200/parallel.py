"""Batch-sharded data parallelism for the flow hot path: one process per GPU, weights replicated.

  * inference / sampling: rows are independent, so every rank takes `shard_rows(x)` and nothing is communicated;
  * training: one gradient allreduce (mean) per optimizer step over `torch.distributed` (NCCL over NVLink on the
    GPU box, gloo in the CPU tests), in a few large flat buckets walked in reverse parameter order (the order in
    which backward produces gradients).

The reference has no distributed code; this is new functionality named by north_star.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn


def shard_bounds(n_rows: int, rank: int, world: int):
    """Contiguous balanced row range [lo, hi) of `rank`: the first n_rows % world ranks get one extra row."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(x: torch.Tensor, rank: int = None, world: int = None):
    """This rank's rows of a [B, ...] batch (a view; ragged and empty shards allowed)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


class DataParallelFlow(nn.Module):
    """Wraps a flow module for data-parallel training.  forward / inverse / log_prob delegate to the wrapped
    module on the local shard; call `sync_gradients()` between `backward()` and `optimizer.step()`."""

    def __init__(self, module: nn.Module, process_group=None, bucket_bytes: int = 64 << 20, broadcast: bool = True):
        super().__init__()
        self.module = module
        self.process_group = process_group
        self.bucket_bytes = int(bucket_bytes)
        if broadcast and dist.is_initialized():
            self.broadcast_parameters()

    # -- delegation ---------------------------------------------------------------------------------------
    def forward(self, z):
        return self.module.forward(z)

    def inverse(self, x):
        return self.module.inverse(x)

    def log_prob(self, x, base_dist):
        return self.module.log_prob(x, base_dist)

    # -- collectives --------------------------------------------------------------------------------------
    def _world(self):
        return dist.get_world_size(self.process_group) if dist.is_initialized() else 1

    @torch.no_grad()
    def broadcast_parameters(self, src: int = 0):
        """Rank `src`'s parameters and buffers overwrite everyone's (start of training / after loading)."""
        for t in list(self.module.parameters()) + list(self.module.buffers()):
            dist.broadcast(t.data, src=src, group=self.process_group)

    def _buckets(self):
        params = [p for p in reversed(list(self.module.parameters())) if p.requires_grad]
        bucket, size, key = [], 0, None
        for p in params:
            k = (p.dtype, p.device)
            nbytes = p.numel() * p.element_size()
            if bucket and (k != key or size + nbytes > self.bucket_bytes):
                yield bucket
                bucket, size = [], 0
            bucket.append(p)
            size += nbytes
            key = k
        if bucket:
            yield bucket

    @torch.no_grad()
    def sync_gradients(self):
        """Average gradients over ranks.  Parameters without a gradient on this rank contribute zeros, so the
        collective sequence is identical on every rank (ragged / empty shards included)."""
        world = self._world()
        if world == 1:
            return
        for bucket in self._buckets():
            flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.process_group)
            flat.div_(world)
            off = 0
            for p in bucket:
                n = p.numel()
                g = flat[off:off + n].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n

    @torch.no_grad()
    def sync_running_stats(self):
        """Average floating-point buffers (BatchNorm running statistics) over ranks, e.g. before evaluation."""
        world = self._world()
        if world == 1:
            return
        for b in self.module.buffers():
            if b.is_floating_point() and ("running_mean" in _name_of(self.module, b) or "running_var" in _name_of(self.module, b)):
                dist.all_reduce(b.data, op=dist.ReduceOp.SUM, group=self.process_group)
                b.data.div_(world)


def _name_of(module, buf):
    for n, b in module.named_buffers():
        if b is buf:
            return n
    return ""
