"""GPU: the reference's four published configs (plots/_common.py:158-170) at a batch size that takes the large-batch
routes -- RealNVP(2, 10, 128) (hidden_dim 128: folded tensor-core GEMMs + streaming first / last layers instead of the
FP32-pipe stack kernel), RealNVPSpline(2, 8, 64), 6 x MAF(2, 64), 6 x IAF(2, 64) (mixed tensor-core / streaming MADE
chain) -- against the CPU oracle in both directions, at the plain north_star bound."""
import pytest
import torch

import nfb200 as N
from oracle import flows_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROWS = 20000           # >= flows.WIDE_OVER_STACK_MIN_ROWS, not a multiple of any tile size


def _compare(what, got, ref, atol, rtol, max_frac=1e-3, hard=20.0):
    got, ref = got.detach().cpu().double(), ref.double()
    assert torch.equal(torch.isnan(got), torch.isnan(ref)), f"{what}: NaN pattern"
    fin = torch.isfinite(ref)
    assert torch.equal(got[~fin & ~torch.isnan(ref)], ref[~fin & ~torch.isnan(ref)]), f"{what}: Inf pattern"
    got, ref = got[fin], ref[fin]
    err = (got - ref).abs() - (atol + rtol * ref.abs())
    frac = (err > 0).double().mean().item()
    print(f"[published] {what}: {frac:.2e} of the elements outside the plain bound, worst excess {err.max().item():.2e}")
    assert frac <= max_frac, f"{what}: {frac:.2e} outside"
    assert ((got - ref).abs() <= hard * (atol + rtol * ref.abs())).all(), what


def _perturbed(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=g))
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.add_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.add_(0.2 * torch.rand(m.running_var.shape, generator=g))
    return model.eval()


@pytest.mark.parametrize("route", ["tc_stack", "gemm"])
@pytest.mark.parametrize("bn_between", [False, True])
@pytest.mark.parametrize("D,hidden,rows", [(2, 128, ROWS), (2, 128, 700), (3, 100, ROWS), (6, 96, 5000)])
def test_realnvp_hidden_up_to_128(bn_between, route, D, hidden, rows, monkeypatch):
    """route tc_stack: the 128-wide variant of the tcgen05 coupling stack kernel (one launch for the whole stack);
    route gemm: what runs without it at large batch -- folded tensor-core GEMMs + streaming first / last layers."""
    from nfb200 import flows as F
    if route == "gemm":
        if rows < F.WIDE_OVER_STACK_MIN_ROWS:
            pytest.skip("the GEMM route starts at WIDE_OVER_STACK_MIN_ROWS")
        monkeypatch.setattr(F, "USE_TENSOR_CORES", False)
    torch.manual_seed(0)
    L = 10 if D == 2 else 4
    model = _perturbed(N.RealNVP(D, L, hidden, batch_norm_between_layers=bn_between), 1)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    specs = [dict(kind="coupling")] * L
    model.to(DEV)
    x = torch.randn(rows, D) * 1.5
    x[3, 0] = float("nan")
    x[5, D - 1] = float("inf")
    assert model.flow.flows[0].fusable(x.to(DEV)) == (route == "tc_stack")
    launches = N._lib.launch_count()
    with torch.no_grad():
        for inverse in (False, True):
            ry, rld = O.flow_model(sd, "flow.", specs, x, inverse, bn_between=bn_between)
            y, ld = model.inverse(x.to(DEV)) if inverse else model.forward(x.to(DEV))
            _compare(f"RealNVP({D},{L},{hidden}) {route} bn={bn_between} inv={inverse} z", y, ry, 1e-5, 1e-5)
            _compare(f"RealNVP({D},{L},{hidden}) {route} bn={bn_between} inv={inverse} log_det", ld, rld, 1e-4, 1e-5)
    if route == "tc_stack":
        assert N._lib.launch_count() - launches == 2, "one launch per direction expected"


def test_spline_hidden_128_large_batch_route():
    torch.manual_seed(0)
    masks = O.realnvp_masks(2, 4)
    model = _perturbed(N.NormalizingFlowModel([N.SplineCouplingLayer(2, 128, m.clone(), num_bins=8) for m in masks]), 2)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    specs = [dict(kind="spline", num_bins=8)] * 4
    model.to(DEV)
    x = torch.randn(ROWS, 2) * 2.0
    with torch.no_grad():
        for inverse in (False, True):
            ry, rld = O.flow_model(sd, "", specs, x, inverse)
            y, ld = model.inverse(x.to(DEV)) if inverse else model.forward(x.to(DEV))
            _compare(f"spline hidden 128 inv={inverse} z", y, ry, 1e-5, 1e-5)
            _compare(f"spline hidden 128 inv={inverse} log_det", ld, rld, 1e-4, 1e-5)


@pytest.mark.parametrize("hidden", [64, 128])
@pytest.mark.parametrize("kind", ["maf", "iaf"])
def test_six_layer_autoregressive_2d_large_batch(kind, hidden):
    """Parallel direction: mixed tensor-core / streaming chain.  Sequential direction (data_dim 2, >= 16384 rows): the
    reference's two-evaluation loop with the second evaluation on that chain (ops.ar_sequential_two_dim)."""
    torch.manual_seed(0)
    cls = N.MaskedAutoregressiveFlow if kind == "maf" else N.InverseAutoregressiveFlow
    model = _perturbed(N.NormalizingFlowModel([cls(2, hidden) for _ in range(6)]), 3)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    specs = [dict(kind=kind)] * 6
    model.to(DEV)
    x = torch.randn(ROWS, 2) * 1.5
    x[3, 0] = float("nan")          # poisons dim 1 of its row in the sequential direction
    x[5, 1] = float("inf")
    x[7, 0] = -float("inf")
    with torch.no_grad():
        for inverse in (False, True):
            ry, rld = O.flow_model(sd, "", specs, x, inverse)
            y, ld = model.inverse(x.to(DEV)) if inverse else model.forward(x.to(DEV))
            _compare(f"6 x {kind}(2,{hidden}) inv={inverse} z", y, ry, 1e-5, 1e-5)
            _compare(f"6 x {kind}(2,{hidden}) inv={inverse} log_det", ld, rld, 1e-4, 1e-5)


@pytest.mark.parametrize("kind,D,H,L,bn_between,use_bn,rows", [
    ("maf", 2, 64, 6, False, False, 4000), ("iaf", 2, 64, 6, True, False, 4000),      # published configs at the published n
    ("maf", 3, 40, 3, False, False, 1500), ("iaf", 5, 64, 4, True, False, 20000),     # data_dim > 2: parallel direction only
    ("maf", 8, 48, 2, False, True, 777), ("iaf", 2, 32, 5, False, True, 130),         # conditioner BatchNorm (eval), small H
])
def test_made_stack_tensor_core_kernel(kind, D, H, L, bn_between, use_bn, rows):
    """Homogeneous MAF / IAF stacks through the containers: ONE made_stack_tc_kernel launch per direction (both directions
    for data_dim == 2, the parallel one otherwise) against the CPU oracle, incl. NaN / Inf rows, between-layer BatchNorm,
    eval-mode conditioner BatchNorm and the fused log-prob head."""
    torch.manual_seed(0)
    cls = N.MaskedAutoregressiveFlow if kind == "maf" else N.InverseAutoregressiveFlow
    model = _perturbed(N.NormalizingFlowModel([cls(D, H, use_batch_norm=use_bn) for _ in range(L)],
                                              batch_norm_between_layers=bn_between), 7)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    specs = [dict(kind=kind)] * L
    model.to(DEV)
    x = torch.randn(rows, D) * 1.5
    x[3, 0] = float("nan")
    x[5, D - 1] = float("inf")
    x[7, 0] = -float("inf")
    x[9, :] = 1e8
    parallel_inverse = (kind == "maf")
    with torch.no_grad():
        for inverse in (False, True):
            ry, rld = O.flow_model(sd, "", specs, x, inverse, bn_between=bn_between)
            xd = x.to(DEV)
            model.inverse(xd) if inverse else model.forward(xd)          # builds the pack (weight folding launches)
            before = N._lib.launch_count()
            y, ld = model.inverse(xd) if inverse else model.forward(xd)
            launches = N._lib.launch_count() - before
            if D == 2 or inverse == parallel_inverse:
                assert launches == 1, f"{kind} D={D} inverse={inverse}: {launches} launches"
            _compare(f"{L} x {kind}({D},{H}) bn={bn_between}/{use_bn} inv={inverse} z", y, ry, 1e-5, 1e-5)
            _compare(f"{L} x {kind}({D},{H}) bn={bn_between}/{use_bn} inv={inverse} log_det", ld, rld, 1e-4, 1e-5)
        # fused head: log N(z; 0, I) + log_det in the same launch, z not stored
        rz, rld = O.flow_model(sd, "", specs, x, True, bn_between=bn_between)
        lp = model.log_prob(x.to(DEV))
        _compare(f"{L} x {kind}({D},{H}) log_prob", lp, O.std_normal_log_prob(rz) + rld, 2e-4, 1e-5)
