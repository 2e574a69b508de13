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

from . import packing


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


def enable_sync_batchnorm(enabled: bool = True, process_group=None) -> None:
    """Process-wide switch: every train-mode BatchNorm1d of the flow layers all-reduces its batch sums over
    `process_group` (ops._SyncBatchNormFn).  No effect without an initialised process group or at world size 1."""
    from . import ops
    ops.set_sync_batchnorm(enabled, process_group)


class DataParallelFlow(nn.Module):
    """Wraps a flow module for data-parallel training.  forward / inverse / log_prob delegate to the wrapped
    module on the local shard; call `sync_gradients()` between `backward()` and `optimizer.step()`.

    With `overlap=True` (default) the gradient buckets are all-reduced *during* backward: a post-accumulate hook per
    parameter counts its bucket down, and a complete bucket is flattened and handed to an asynchronous all-reduce on
    the communication stream while autograd keeps producing the earlier layers' gradients.  Buckets are always
    launched in the same (reverse-parameter) order on every rank; `sync_gradients()` launches what backward did not
    reach (parameters without a gradient contribute zeros), waits, averages and scatters the results back.
    Gradient accumulation over micro-batches: wrap all but the last backward in `no_sync()`."""

    def __init__(self, module: nn.Module, process_group=None, bucket_bytes: int = 64 << 20, broadcast: bool = True,
                 overlap: bool = True, sync_batchnorm: bool = True):
        super().__init__()
        self.module = module
        self.process_group = process_group
        # train-mode BatchNorm of the coupling conditioners: all-reduce the (sum x, sum x^2, n) triples so that N ranks
        # on shards of a batch compute what one rank computes on the whole batch
        enable_sync_batchnorm(sync_batchnorm and any(isinstance(m, nn.BatchNorm1d) for m in module.modules()),
                              process_group)
        self.bucket_bytes = int(bucket_bytes)
        self.overlap = bool(overlap)
        self._bucket_list = None
        self._pending = []            # (bucket index, flat tensor, work handle) in launch order
        self._ready = None            # per bucket: number of parameters whose gradient has arrived this round
        self._next = 0                # next bucket to launch
        self._dirty = False           # a bucket fired twice without sync_gradients(): fall back to the plain path
        self._hooks_on = True
        self._hook_handles = []
        if broadcast and dist.is_initialized():
            self.broadcast_parameters()
        if self.overlap and dist.is_initialized() and self._world() > 1:
            self._install_hooks()

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
            dist.broadcast(t, src=src, group=self.process_group)      # on the tensor itself: bumps its version
        packing.invalidate_caches()

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

    # -- overlap with backward ------------------------------------------------------------------------------
    def _install_hooks(self):
        """First round: the hooks only record the order in which gradients arrive (it depends on the direction the model
        is trained in: `inverse` walks the layers backwards, so its gradients arrive in *forward* parameter order).
        sync_gradients() then adopts rank 0's order on every rank and cuts the buckets along it."""
        self._params = [p for p in self.module.parameters() if p.requires_grad]
        self._arrival = []
        self._learning = True
        self._bucket_of = {}
        for i, p in enumerate(self._params):
            self._hook_handles.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))

    def _adopt_arrival_order(self):
        n = len(self._params)
        seen = set()
        order = [i for i in self._arrival if not (i in seen or seen.add(i))]
        order += [i for i in reversed(range(n)) if i not in seen]           # never fired: reverse parameter order
        dev = self._params[0].device if n else torch.device("cpu")
        t = torch.tensor(order, dtype=torch.int64, device=dev)
        dist.broadcast(t, src=0, group=self.process_group)                  # every rank cuts the same buckets
        order = t.tolist()
        buckets, bucket, size, key = [], [], 0, None
        for i in order:
            p = self._params[i]
            k = (p.dtype, p.device)
            nbytes = p.numel() * p.element_size()
            if bucket and (k != key or size + nbytes > self.bucket_bytes):
                buckets.append(bucket)
                bucket, size = [], 0
            bucket.append(p)
            size += nbytes
            key = k
        if bucket:
            buckets.append(bucket)
        self._bucket_list = buckets
        self._bucket_of = {id(p): bi for bi, b in enumerate(buckets) for p in b}
        self._ready = [0] * len(buckets)
        self._learning = False

    def _make_hook(self, pi):
        def hook(param):
            if not self._hooks_on:
                return
            if self._learning:
                self._arrival.append(pi)
                return
            bi = self._bucket_of[id(param)]
            if bi < self._next or self._ready[bi] >= len(self._bucket_list[bi]):
                self._dirty = True                      # second backward without sync_gradients() / no_sync()
                return
            self._ready[bi] += 1
            while self._next < len(self._bucket_list) and self._ready[self._next] == len(self._bucket_list[self._next]):
                self._launch(self._next)
                self._next += 1
        return hook

    @torch.no_grad()
    def _launch(self, bi):
        bucket = self._bucket_list[bi]
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.process_group, async_op=True)
        self._pending.append((bi, flat, work))

    class _NoSync:
        def __init__(self, owner):
            self.owner = owner

        def __enter__(self):
            self.prev = self.owner._hooks_on
            self.owner._hooks_on = False

        def __exit__(self, *exc):
            self.owner._hooks_on = self.prev

    def no_sync(self):
        """Context manager: backward passes inside do not start gradient all-reduces (micro-batch accumulation)."""
        return DataParallelFlow._NoSync(self)

    @torch.no_grad()
    def _scatter_back(self, bucket, flat, world):
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
    def sync_gradients(self):
        """Average gradients over ranks.  Parameters without a gradient on this rank contribute zeros, so the
        collective sequence is identical on every rank (ragged / empty shards included)."""
        world = self._world()
        if world == 1:
            return
        if getattr(self, "_learning", False):
            for bucket in self._buckets():                      # first round: plain path, then fix the bucket order
                flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.process_group)
                self._scatter_back(bucket, flat, world)
            self._adopt_arrival_order()
            return
        if self._bucket_list is not None and not self._dirty:
            while self._next < len(self._bucket_list):          # buckets backward did not complete on this rank
                self._launch(self._next)
                self._next += 1
            for bi, flat, work in self._pending:
                work.wait()
                self._scatter_back(self._bucket_list[bi], flat, world)
        else:
            for _, _, work in self._pending:                    # drain what was started, then the plain path
                work.wait()
            for bucket in (self._bucket_list if self._bucket_list is not None else self._buckets()):
                flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.process_group)
                self._scatter_back(bucket, flat, world)
        self._pending = []
        self._next = 0
        self._dirty = False
        if self._ready is not None:
            self._ready = [0] * len(self._bucket_list)

    @torch.no_grad()
    def sync_running_stats(self):
        """Average floating-point buffers (BatchNorm running statistics) over ranks, e.g. before evaluation."""
        world = self._world()
        if world == 1:
            return
        for b in self.module.buffers():
            if b.is_floating_point() and ("running_mean" in _name_of(self.module, b) or "running_var" in _name_of(self.module, b)):
                dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.process_group)
                b.div_(world)
        packing.invalidate_caches()


def _name_of(module, buf):
    for n, b in module.named_buffers():
        if b is buf:
            return n
    return ""
