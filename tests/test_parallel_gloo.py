"""CPU, world_size 2 over gloo: the host-side logic of the data-parallel path (row sharding, parameter broadcast,
bucketed gradient allreduce).  Kernels are not launched here (no GPU); gradients are planted by hand."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import nfb200 as N
from nfb200 import parallel as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, bucket_bytes, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                        # different initial weights per rank
        model = N.RealNVP(4, 2, 8)
        dp = P.DataParallelFlow(model, bucket_bytes=bucket_bytes)      # broadcasts rank 0's weights
        ref = [torch.zeros_like(p) for p in model.parameters()]
        for r, p in zip(ref, model.parameters()):
            r.copy_(p.detach())
            dist.broadcast(r, src=0)
        same_weights = all(torch.equal(r, p.detach()) for r, p in zip(ref, model.parameters()))
        # plant rank-dependent gradients; leave one parameter without a gradient on rank 1
        params = list(model.parameters())
        for i, p in enumerate(params):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        if rank == 1:
            params[3].grad = None
        dp.sync_gradients()
        ok = True
        for i, p in enumerate(params):
            expect = (1 + 2) * (i + 1) / 2.0 if i != 3 else 1 * (i + 1) / 2.0
            ok &= bool(torch.allclose(p.grad, torch.full_like(p, expect)))
        # running stats
        for n, b in model.named_buffers():
            if "running_mean" in n:
                b.fill_(float(rank))
        dp.sync_running_stats()
        stats_ok = all(bool(torch.allclose(b, torch.full_like(b, 0.5))) for n, b in model.named_buffers()
                       if "running_mean" in n)
        # row sharding: ragged split, concatenation restores the batch
        x = torch.arange(11 * 3, dtype=torch.float32).view(11, 3)
        mine = P.shard_rows(x)
        gathered = [torch.zeros(6, 3) for _ in range(world)]
        pad = torch.zeros(6, 3)
        pad[: mine.shape[0]] = mine
        dist.all_gather(gathered, pad)
        sizes = [P.shard_bounds(11, r, world) for r in range(world)]
        cat = torch.cat([g[: hi - lo] for g, (lo, hi) in zip(gathered, sizes)])
        out[rank] = (same_weights, ok, stats_ok, bool(torch.equal(cat, x)), mine.shape[0])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes", [64 << 20, 256])
def test_data_parallel_host_logic_world2(bucket_bytes):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), bucket_bytes, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        same_weights, grads_ok, stats_ok, rows_ok, n = out[rank]
        assert same_weights and grads_ok and stats_ok and rows_ok
    assert out[0][4] == 6 and out[1][4] == 5


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            edges = [P.shard_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
