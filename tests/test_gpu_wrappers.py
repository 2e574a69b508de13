"""GPU: the reference's optimisation wrappers keep working around the CUDA modules: autocast
regions (MixedPrecisionFlow, optimization/mixed_precision.py:89-105), activation checkpointing (CheckpointedFlow,
optimization/gradient_checkpointing.py:36-64), Flow.log_prob / Flow.sample, clip_grad_norm_, deepcopy / state_dict
round trips, and the samples_per_sec harness recipe (plots/_common.py:264-274)."""
import copy
import time

import pytest
import torch
from torch.utils.checkpoint import checkpoint

import nfb200 as N

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _models():
    torch.manual_seed(0)
    alt = torch.tensor([1., 0., 1., 0.])
    ms = {
        "realnvp": N.RealNVP(4, 4, 16),
        "spline": N.RealNVPSpline(4, 2, 16),
        "maf": N.SequentialFlow([N.MaskedAutoregressiveFlow(4, 16), N.MaskedAutoregressiveFlow(4, 16)]),
        "iaf": N.InverseAutoregressiveFlow(4, 16),
        "coupling": N.CouplingLayer(4, 16, alt),
    }
    gen = torch.Generator().manual_seed(1)
    for m in ms.values():
        with torch.no_grad():
            for p in m.parameters():
                p.add_(0.1 * torch.randn(p.shape, generator=gen))
        m.to(DEV)
    return ms


@pytest.mark.parametrize("amp_dtype", [torch.bfloat16, torch.float16])
def test_autocast_regions(amp_dtype):
    for name, m in _models().items():
        m.eval()
        x = torch.randn(64, 4, device=DEV)
        with torch.no_grad():
            ref = m.inverse(x)
            with torch.autocast("cuda", dtype=amp_dtype):
                got = m.inverse(x)
                got_half = m.inverse(x.to(amp_dtype))          # half inputs are widened, not rejected
        assert torch.equal(ref[0], got[0]) and torch.equal(ref[1], got[1]), name
        assert got_half[0].dtype == torch.float32 and torch.isfinite(got_half[0]).all()
        # training under autocast with a GradScaler-free bf16 recipe
        m.train()
        xt = torch.randn(64, 4, device=DEV)
        with torch.autocast("cuda", dtype=amp_dtype):
            z, ld = m.inverse(xt)
            loss = -(N.ops.std_normal_log_prob(z, ld)).mean()
        loss.backward()
        assert all(p.grad is None or torch.isfinite(p.grad).all() for p in m.parameters()), name


def test_activation_checkpointing_matches_plain_backward():
    for name, m in _models().items():
        m.eval()
        x = torch.randn(32, 4, device=DEV, requires_grad=True)
        z, ld = m.inverse(x)
        (z.sum() + ld.sum()).backward()
        g_plain = [p.grad.clone() for p in m.parameters() if p.grad is not None]
        gx_plain = x.grad.clone()
        m.zero_grad()
        x.grad = None
        z2, ld2 = checkpoint(m.inverse, x, use_reentrant=False)
        (z2.sum() + ld2.sum()).backward()
        g_ckpt = [p.grad for p in m.parameters() if p.grad is not None]
        assert torch.allclose(gx_plain, x.grad, rtol=1e-6, atol=1e-7), name
        assert len(g_plain) == len(g_ckpt) and all(torch.allclose(a, b, rtol=1e-6, atol=1e-7) for a, b in zip(g_plain, g_ckpt)), name


def test_flow_log_prob_and_sample_api():
    base = torch.distributions.Normal(torch.zeros(4, device=DEV), torch.ones(4, device=DEV))
    mv = torch.distributions.MultivariateNormal(torch.zeros(4, device=DEV), torch.eye(4, device=DEV))
    for name in ("maf", "iaf", "coupling"):
        m = _models()[name].eval()
        x = torch.randn(16, 4, device=DEV)
        with torch.no_grad():
            lp = m.log_prob(x, base)
            lp2 = m.log_prob(x, mv)
            z, ld = m.inverse(x)
            fused = N.ops.std_normal_log_prob(z, ld)
            s = m.sample(8, base, device=DEV)
        assert lp.shape == (16,) and torch.allclose(lp, lp2, atol=1e-5) and torch.allclose(lp, fused, atol=1e-5), name
        assert s.shape == (8, 4) and torch.isfinite(s).all()


def test_clip_grad_deepcopy_state_dict_roundtrip():
    m = _models()["realnvp"]
    m.train()
    x = torch.randn(128, 4, device=DEV)
    z, ld = m.inverse(x)
    (-(N.ops.std_normal_log_prob(z, ld)).mean()).backward()
    total = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    assert torch.isfinite(total)
    m.eval()
    m2 = copy.deepcopy(m)
    m3 = N.RealNVP(4, 4, 16).to(DEV).eval()
    m3.load_state_dict(m.state_dict())
    with torch.no_grad():
        a, b, c = m.forward(x), m2.forward(x), m3.forward(x)
    assert torch.equal(a[0], b[0]) and torch.equal(a[0], c[0]) and torch.equal(a[1], c[1])
    # packed-weight caches follow in-place parameter updates
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(1.01)
        d = m.forward(x)
    assert not torch.equal(a[0], d[0])


def test_reference_samples_per_sec_recipe():
    """plots/_common.py:264-274: model.forward(z), n=4000, dim=2, eval, 1 warm-up + 3 reps."""
    torch.manual_seed(0)
    m = N.RealNVPSpline(2, 8, 64).to(DEV).eval()
    z = torch.randn(4000, 2, device=DEV)
    with torch.no_grad():
        m.forward(z)
        torch.cuda.synchronize()
        t = time.time()
        for _ in range(3):
            x, _ = m.forward(z)
        torch.cuda.synchronize()
        dt = (time.time() - t) / 3
    assert x.shape == (4000, 2) and 4000 / dt > 1e5


def test_graphed_train_step_matches_eager_steps():
    """CUDA-graph replay of forward + backward + optimizer (graphs.GraphedTrainStep) follows the eager trajectory."""
    import copy
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    base = N.RealNVP(2, 4, 32).to(dev).train()
    x = torch.randn(3, 2000, 2, device=dev)

    def loss_fn(m, xin):
        z, ld = m.inverse(xin)
        return -N.ops.std_normal_log_prob(z, ld).mean()

    eager_m = copy.deepcopy(base)
    opt_e = torch.optim.SGD(eager_m.parameters(), lr=1e-2, momentum=0.9)       # SGD: no sign(g)-style sensitivity
    graph_m = copy.deepcopy(base)
    opt_g = torch.optim.SGD(graph_m.parameters(), lr=1e-2, momentum=0.9)
    # the constructor's warm-up steps are real optimizer steps on example_input: give the eager twin the same ones
    step = N.graphs.GraphedTrainStep(graph_m, opt_g, loss_fn, x[0], warmup=2)
    for _ in range(2):                      # the 2 warm-up steps (capture records the step, it does not run it)
        opt_e.zero_grad(set_to_none=True)
        loss_fn(eager_m, x[0]).backward()
        opt_e.step()
    losses_g, losses_e = [], []
    for i in range(3):
        losses_g.append(float(step(x[i])))
        opt_e.zero_grad(set_to_none=True)
        le = loss_fn(eager_m, x[i])
        le.backward()
        opt_e.step()
        losses_e.append(float(le))
    assert all(abs(a - b) <= 1e-4 * (1 + abs(b)) for a, b in zip(losses_g, losses_e)), (losses_g, losses_e)
    for (k, p), (_, q) in zip(graph_m.named_parameters(), eager_m.named_parameters()):
        assert torch.allclose(p, q, atol=1e-5, rtol=1e-4), k


@pytest.mark.parametrize("make,shape", [(lambda: N.RealNVP(2, 8, 64), (2,)), (lambda: N.RealNVPSpline(2, 8, 64), (2,)),
                                        (lambda: N.MaskedAutoregressiveFlow(8, 64), (8,)),
                                        (lambda: N.InverseAutoregressiveFlow(8, 64), (8,))])
def test_reference_flow_profiler_recipe(make, shape):
    """FlowProfiler.profile_flow / _profile_single_batch_size (src/flows/utils/profiling.py:63-200) restated: .to(device),
    .eval(), warm-up forward + inverse under no_grad, CUDA-event timing per call at batch sizes 1, 8, 32, 64, peak-memory
    and parameter-count queries."""
    flow = make().to(DEV)
    flow.eval()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for batch_size in (1, 8, 32, 64):
        full = (batch_size,) + shape
        for _ in range(2):
            x = torch.randn(full, device=DEV)
            with torch.no_grad():
                flow.forward(x)
                flow.inverse(x)
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        for fn in (flow.forward, flow.inverse):
            x = torch.randn(full, device=DEV)
            start.record()
            with torch.no_grad():
                out, log_det = fn(x)
            end.record()
            torch.cuda.synchronize()
            assert out.shape == full and log_det.shape == (batch_size,) and start.elapsed_time(end) > 0
            assert bool(torch.isfinite(out).all()) and bool(torch.isfinite(log_det).all())
        assert torch.cuda.max_memory_allocated() > 0
    assert sum(p.numel() for p in flow.parameters()) > 0


# ------------------------------------------------------------------------------------------------
# round 2: cache invalidation, inference mode, mask plans, fused log-prob head (VERDICT f1)
# ------------------------------------------------------------------------------------------------
def _nll(model, x):
    z, ld = model.inverse(x)
    return -(N.ops.std_normal_log_prob(z, ld)).mean()


@pytest.mark.parametrize("kind", ["spline", "realnvp", "maf"])
def test_eval_after_graph_replay_sees_the_new_weights(kind):
    """eval (packs built) -> CUDA-graph replays that contain the optimizer step -> eval: the replay bumps no version
    counter, so GraphedTrainStep must invalidate the derived weight layouts itself."""
    torch.manual_seed(0)
    m = {"spline": lambda: N.RealNVPSpline(2, 2, 16), "realnvp": lambda: N.RealNVP(2, 2, 16),
         "maf": lambda: N.MaskedAutoregressiveFlow(8, 32)}[kind]().to(DEV)
    D = 8 if kind == "maf" else 2
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.1 * torch.randn_like(p))
    x = torch.randn(512, D, device=DEV)
    xe = torch.randn(300, D, device=DEV)
    m.eval()
    with torch.no_grad():
        z0, _ = m.inverse(xe)                                     # builds the fused packs from the initial weights
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=5e-2, capturable=True)
    step = N.graphs.GraphedTrainStep(m, opt, _nll, x)
    for _ in range(5):
        step(x)
    m.eval()
    with torch.no_grad():
        z1, ld1 = m.inverse(xe)                                   # fused route, must use the replayed weights
    xr = xe.clone().requires_grad_()
    z2, ld2 = m.inverse(xr)                                       # layered route reads the parameters directly
    assert not torch.allclose(z0, z1), "training did not move the model: the test would be vacuous"
    assert torch.allclose(z1, z2.detach(), rtol=1e-5, atol=1e-5), (z1 - z2).abs().max().item()
    assert torch.allclose(ld1, ld2.detach(), rtol=1e-5, atol=1e-4)


def test_data_writes_are_covered_by_invalidate_caches():
    """`p.data.add_()` (the reference's EMA pattern, consistency_flow.py:28) bumps no version counter either."""
    torch.manual_seed(0)
    m = N.RealNVPSpline(2, 2, 16).to(DEV).eval()
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.1 * torch.randn_like(p))
    x = torch.randn(256, 2, device=DEV)
    with torch.no_grad():
        z0, _ = m.inverse(x)
        for p in m.parameters():
            p.data.add_(0.05 * torch.randn_like(p))
        N.invalidate_caches()
        z1, _ = m.inverse(x)
    z2, _ = m.inverse(x.clone().requires_grad_())
    assert not torch.allclose(z0, z1)
    assert torch.allclose(z1, z2.detach(), rtol=1e-5, atol=1e-5)


def test_inference_mode_on_every_route():
    """torch.inference_mode tensors carry no version counter: the fused, wide (tcgen05 GEMM) and float64 layered routes
    must all run there, as the reference does."""
    torch.manual_seed(0)
    wide = N.SplineCouplingLayer(16, 160, torch.tensor([1., 0.] * 8), num_bins=8).to(DEV).eval()      # D > 8: not fusable
    small = N.RealNVPSpline(2, 2, 16).to(DEV).eval()
    maf = N.MaskedAutoregressiveFlow(8, 32).to(DEV).eval()
    with torch.no_grad():
        for m in (wide, small, maf):
            for p in m.parameters():
                p.add_(0.05 * torch.randn_like(p))
    x16, x2, x8 = torch.randn(512, 16, device=DEV), torch.randn(512, 2, device=DEV), torch.randn(512, 8, device=DEV)
    with torch.no_grad():
        ref = [wide.inverse(x16), small.inverse(x2), maf.inverse(x8), maf.forward(x8)]
    with torch.inference_mode():
        got = [wide.inverse(x16.clone()), small.inverse(x2.clone()), maf.inverse(x8.clone()), maf.forward(x8.clone())]
        w64 = copy.deepcopy(wide).double()
        z64, _ = w64.inverse(x16.double())
    for (a, b), (c, d) in zip(ref, got):
        assert torch.equal(a, c) and torch.equal(b, d)
    assert torch.allclose(z64.float(), ref[0][0], rtol=1e-4, atol=1e-4)


def test_mask_plan_is_pinned_to_its_mask():
    """A freed MADE mask's address can be handed to a new mask of the same shape with other degrees: the cached zero
    structure must not be reused for it."""
    from nfb200 import ops
    a = N.MADE(64, 512).net[2].mask.to(DEV)                     # [512, 512], degrees of data_dim 64
    pa = ops.mask_plan(a)
    assert pa is not None
    ext_a = pa.k_extent.clone()
    ptr_a = a.data_ptr()
    del a, pa
    b = None
    for _ in range(8):                                          # the caching allocator hands the block out again
        b = N.MADE(16, 512).net[2].mask.to(DEV)
        if b.data_ptr() == ptr_a:
            break
    pb = ops.mask_plan(b)
    nz = b != 0
    last = torch.where(nz, torch.arange(1, 513, device=DEV)[None, :], 0).amax(dim=1).view(-1, 64).amax(dim=1).to(torch.int32)
    assert torch.equal(pb.k_extent, last)
    assert not torch.equal(pb.k_extent, ext_a)


@pytest.mark.parametrize("kind", ["spline_tc", "spline_simt", "coupling_tc", "coupling_simt", "spline_bn"])
def test_fused_log_prob_head_matches_separate_head(kind, monkeypatch):
    """Flow.log_prob with a standard-normal base: one launch, head in the last layer's epilogue, z not written."""
    torch.manual_seed(0)
    monkeypatch.setattr(N.flows, "USE_TENSOR_CORES", kind.endswith("_tc") or kind == "spline_bn")
    if kind.startswith("spline"):
        m = N.RealNVPSpline(2, 4, 64, batch_norm_between_layers=(kind == "spline_bn"))
    else:
        m = N.RealNVP(2, 4, 64)
    m = m.to(DEV).eval()
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.05 * torch.randn_like(p))
        for name, b in m.named_buffers():                         # non-trivial between-layer BatchNorm statistics
            if name.endswith("running_mean"):
                b.add_(0.3 * torch.randn_like(b))
            elif name.endswith("running_var"):
                b.mul_(1.0 + 0.5 * torch.rand_like(b))
    x = torch.randn(3000, 2, device=DEV)
    base = torch.distributions.Normal(torch.zeros(2, device=DEV), torch.ones(2, device=DEV))
    with torch.no_grad():
        z, ld = m.inverse(x)
        want = base.log_prob(z).sum(dim=1) + ld                   # the reference's formula in eager torch
        before = N._lib.launch_count()
        got = m.log_prob(x, base)
        assert N._lib.launch_count() - before == 1, "log_prob must be ONE launch (head fused into the stack kernel)"
        got_default = m.log_prob(x)
        got_flow = m.flow.log_prob(x, torch.distributions.MultivariateNormal(torch.zeros(2, device=DEV), torch.eye(2, device=DEV)))
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5), (got - want).abs().max().item()
    assert torch.equal(got, got_default) and torch.equal(got, got_flow)


def test_log_prob_with_other_bases_and_gradients():
    """Non-standard bases go through the torch distribution; a standard base with autograd uses the head kernel + backward."""
    torch.manual_seed(0)
    m = N.RealNVP(4, 2, 16).to(DEV)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.1 * torch.randn_like(p))
    m.eval()
    x = torch.randn(128, 4, device=DEV)
    wide = torch.distributions.Normal(torch.zeros(4, device=DEV), 2 * torch.ones(4, device=DEV))
    assert not N.flows.is_std_normal(wide, 4)
    with torch.no_grad():
        z, ld = m.inverse(x)
        assert torch.allclose(m.log_prob(x, wide), wide.log_prob(z).sum(1) + ld, rtol=1e-6, atol=1e-6)
    xr = x.clone().requires_grad_()
    lp = m.log_prob(xr)
    lp.sum().backward()
    xr2 = x.clone().requires_grad_()
    z2, ld2 = m.inverse(xr2)
    ((-0.5 * z2 * z2).sum(1) - 2 * 1.8378770664093453 + ld2).sum().backward()
    assert torch.allclose(xr.grad, xr2.grad, rtol=1e-5, atol=1e-6)
    s = m.sample(77)
    assert s.shape == (77, 4) and s.is_cuda and torch.isfinite(s).all()


# ------------------------------------------------------------------------------------------------
# round 2: synchronised BatchNorm statistics -- two "ranks" emulated on one GPU
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,relu", [(64, True), (48, True), (30, False)])
def test_staged_batchnorm_on_two_shards_equals_one_batch(H, relu):
    """The staged entry points (stage 1: shard sums; all-reduce = a plain add here; stage 2: statistics over the global
    count, apply / input gradient) on two uneven shards reproduce the unsharded train-mode BatchNorm: outputs, running
    statistics, gx, and ggamma / gbeta as the SUM of the shards' local sums."""
    from nfb200 import _lib as L
    torch.manual_seed(0)
    B, cut = 3000, 1234
    x = torch.randn(B, H, device=DEV) * 1.7 + 0.3
    gy = torch.randn(B, H, device=DEV)
    bn = torch.nn.BatchNorm1d(H).to(DEV)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
    # unsharded reference through the ordinary op
    bn_ref = copy.deepcopy(bn).train()
    xr = x.clone().requires_grad_()
    yr = N.ops.batchnorm_relu(xr, bn_ref, relu=relu)
    yr.backward(gy)
    # two shards through the staged C ABI
    call, ptr, stream = L.call, L.ptr, L.stream
    shards = [x[:cut].contiguous(), x[cut:].contiguous()]
    gys = [gy[:cut].contiguous(), gy[cut:].contiguous()]
    ws = [torch.empty(2 * H, dtype=torch.float64, device=DEV) for _ in shards]
    for xs, w in zip(shards, ws):
        call("nf_batchnorm_forward_staged", ptr(xs), None, None, None, None, None, None, None, ptr(w), xs.shape[0], H, 0.0, bn.eps,
             int(relu), 1, 0, 0, stream())
    glob = ws[0] + ws[1]                                             # the all-reduce
    ys, sms, srs, rms, rvs = [], [], [], [], []
    for si, xs in enumerate(shards):
        y = torch.empty_like(xs)
        sm, sr = torch.empty(H, device=DEV), torch.empty(H, device=DEV)
        rm, rv = bn.running_mean.clone(), bn.running_var.clone()
        # shard 0: global count from the host; shard 1: count = -1, read from workspace[2H] on the device (what
        # ops._SyncBatchNormFn does: the count is all-reduced with the sums, no host read)
        wsg = torch.cat([glob, glob.new_tensor([float(B)])])
        call("nf_batchnorm_forward_staged", ptr(xs), ptr(bn.weight), ptr(bn.bias), ptr(rm), ptr(rv), ptr(y), ptr(sm), ptr(sr),
             ptr(wsg), xs.shape[0], H, 0.1, bn.eps, int(relu), 2, B if si == 0 else -1, 0, stream())
        ys.append(y); sms.append(sm); srs.append(sr); rms.append(rm); rvs.append(rv)
    assert torch.allclose(torch.cat(ys), yr.detach(), rtol=1e-6, atol=1e-6)
    assert torch.equal(rms[0], rms[1]) and torch.allclose(rms[0], bn_ref.running_mean, rtol=1e-6, atol=1e-7)
    assert torch.allclose(rvs[0], bn_ref.running_var, rtol=1e-6, atol=1e-7)
    wsb = [torch.empty(2 * H, dtype=torch.float64, device=DEV) for _ in shards]
    for xs, y, g, w in zip(shards, ys, gys, wsb):
        # first shard: ReLU mask re-derived from x (gamma / beta given); second shard: mask read from y
        first = xs is shards[0]
        call("nf_batchnorm_backward_staged", ptr(xs), ptr(y), ptr(bn.weight) if first else None, ptr(sms[0]), ptr(srs[0]), ptr(g),
             None, None, None, ptr(w), xs.shape[0], H, int(relu), 1, 0, 0, ptr(bn.bias) if first else None, stream())
    globb = wsb[0] + wsb[1]
    gxs = []
    for si, (xs, y, g) in enumerate(zip(shards, ys, gys)):
        gx = torch.empty_like(xs)
        gg, gb = torch.empty(H, device=DEV), torch.empty(H, device=DEV)
        wsg = torch.cat([globb, globb.new_tensor([float(B)])])
        call("nf_batchnorm_backward_staged", ptr(xs), ptr(y), ptr(bn.weight), ptr(sms[0]), ptr(srs[0]), ptr(g), ptr(gx), ptr(gg),
             ptr(gb), ptr(wsg), xs.shape[0], H, int(relu), 2, B if si == 0 else -1, 0, ptr(bn.bias) if si == 0 else None, stream())
        gxs.append(gx)
    assert torch.allclose(torch.cat(gxs), xr.grad, rtol=1e-5, atol=1e-6)
    assert torch.allclose(globb[:H].float(), bn_ref.weight.grad, rtol=1e-5, atol=1e-5)
    assert torch.allclose(globb[H:].float(), bn_ref.bias.grad, rtol=1e-5, atol=1e-5)
