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
