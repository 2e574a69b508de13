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


def _overlap_worker(rank, world, port, bucket_bytes, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(7)                                 # same weights everywhere
        net = torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                                  torch.nn.Linear(16, 3))
        twin = torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                                   torch.nn.Linear(16, 3))
        twin.load_state_dict(net.state_dict())
        dp = P.DataParallelFlow(net, bucket_bytes=bucket_bytes)
        assert dp._hook_handles                              # hooks installed (world 2)
        g = torch.Generator().manual_seed(50 + rank)
        x1, x2 = torch.randn(9, 5, generator=g), torch.randn(4, 5, generator=g)
        # round 0 learns the gradient arrival order (plain all-reduce), later rounds overlap
        net.zero_grad(set_to_none=True)
        net(x2).pow(2).sum().backward()
        dp.sync_gradients()
        assert dp._bucket_list is not None and not dp._learning
        first_bucket_param = dp._bucket_list[0][0]
        arrival_ok = any(first_bucket_param is q for q in net[-1].parameters())   # the last layer's gradients arrive first

        def expected(batches):
            twin.zero_grad(set_to_none=True)
            for xb in batches:
                twin(xb).pow(2).sum().backward()
            outs = []
            for p in twin.parameters():
                t = p.grad.clone()
                dist.all_reduce(t)
                outs.append(t / world)
            return outs

        results = []
        # (1) one backward: buckets are reduced from the hooks, sync_gradients only waits and scatters
        net.zero_grad(set_to_none=True)
        net(x1).pow(2).sum().backward()
        launched_in_backward = len(dp._pending)
        dp.sync_gradients()
        results.append(all(torch.allclose(p.grad, e, atol=1e-6) for p, e in zip(net.parameters(), expected([x1]))))
        # (2) accumulation over two micro-batches with no_sync on the first
        net.zero_grad(set_to_none=True)
        with dp.no_sync():
            net(x1).pow(2).sum().backward()
        net(x2).pow(2).sum().backward()
        dp.sync_gradients()
        results.append(all(torch.allclose(p.grad, e, atol=1e-6) for p, e in zip(net.parameters(), expected([x1, x2]))))
        # (3) two backwards without no_sync: detected, plain path gives the right average of the accumulated gradients
        net.zero_grad(set_to_none=True)
        net(x1).pow(2).sum().backward()
        net(x2).pow(2).sum().backward()
        dp.sync_gradients()
        results.append(all(torch.allclose(p.grad, e, atol=1e-6) for p, e in zip(net.parameters(), expected([x1, x2]))))
        # (4) a rank whose backward skips the last layers' parameters (unused branch): zeros contributed, same order
        net.zero_grad(set_to_none=True)
        if rank == 0:
            net(x1).pow(2).sum().backward()
        else:
            net[:3](x1).pow(2).sum().backward()              # only the first two Linears get gradients on rank 1
        dp.sync_gradients()
        exp4 = []
        twin.zero_grad(set_to_none=True)
        (twin(x1) if rank == 0 else twin[:3](x1)).pow(2).sum().backward()
        for p in twin.parameters():
            t = p.grad.clone() if p.grad is not None else torch.zeros_like(p)
            dist.all_reduce(t)
            exp4.append(t / world)
        results.append(all(torch.allclose(p.grad, e, atol=1e-6) for p, e in zip(net.parameters(), exp4)))
        results.append(arrival_ok)
        out[rank] = (results, launched_in_backward, len(dp._bucket_list))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes", [64 << 20, 300])
def test_overlapped_gradient_allreduce_world2(bucket_bytes):
    """Gradient buckets all-reduced from post-accumulate hooks during backward (parallel.DataParallelFlow, overlap=True):
    same averages as the plain path, fixed bucket order on every rank, no_sync accumulation, double-backward fallback."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_overlap_worker, args=(world, _free_port(), bucket_bytes, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        results, launched, nb = out[rank]
        assert all(results), (rank, results)
        assert launched == nb                                # every bucket was started before sync_gradients()
    assert out[0][2] == (1 if bucket_bytes > 1000 else out[0][2]) and out[0][2] >= 1


def _bn_sums_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nfb200 import ops
        H = 6
        torch.manual_seed(0)
        x = torch.randn(11, H, dtype=torch.float64)                 # the global batch, identical on every rank
        lo, hi = P.shard_bounds(11, rank, world)
        mine = x[lo:hi]
        ws = torch.cat([mine.sum(0), (mine * mine).sum(0)])
        ws, count = ops.allreduce_bn_sums(ws, mine.shape[0])
        mean = ws[:H] / count
        var = ws[H:] / count - mean * mean
        ok = count == 11 and torch.allclose(mean, x.mean(0)) and torch.allclose(var, x.var(0, unbiased=False))
        # the switch the layers consult: on only with an initialised group and world > 1
        P.enable_sync_batchnorm(True)
        on = ops._sync_bn_world()
        P.enable_sync_batchnorm(False)
        off = ops._sync_bn_world()
        dp = P.DataParallelFlow(N.RealNVP(4, 2, 8))                 # has BatchNorm1d layers: the wrapper switches it on
        wrapped = ops._sync_bn_world()
        P.enable_sync_batchnorm(False)
        out[rank] = (bool(ok), on, off, wrapped)
    finally:
        dist.destroy_process_group()


def test_sync_batchnorm_host_logic_world2():
    """(sum x, sum x^2, n) all-reduce of the synchronised BatchNorm: global mean / biased variance from two uneven shards."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bn_sums_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for rank in range(world):
        ok, on, off, wrapped = out[rank]
        assert ok and on == 2 and off == 1 and wrapped == 2
