"""GPU parity: the CUDA path (through the nn.Module surface -> ctypes -> C ABI -> sm_100a kernels) against
  * the committed golden vectors produced by the unmodified reference (tests/golden), and
  * the CPU oracle on seeded inputs.
Tolerances (north_star): z within 1e-5*max(1,|z|), row log-det within 1e-4 absolute; written at each assert."""
import pytest
import torch

import nfb200 as N
from oracle import flows_oracle as O
from tests import golden_util as G
from tests.build_util import MODULE_KINDS, assert_close, build

pytestmark = pytest.mark.gpu

Z_ATOL, Z_RTOL = 1e-5, 1e-5      # fp32 z: |dz| <= 1e-5 * (1 + |z|)
LD_ATOL, LD_RTOL = 1e-4, 1e-5    # row log-det: 1e-4 absolute (+1e-5 relative for |ld| >> 1)

MODULE_CASES = [n for n in G.golden_names() if G.load(n)["kind"] in MODULE_KINDS]


def _dev():
    return torch.device("cuda:0")


def _run(m, x, inverse):
    return m.inverse(x) if inverse else m.forward(x)


def _loose(g):
    """stress rows drive the conditioners to 1e10-scale activations where fp32 GEMM summation order decides
    between saturated values; compare those rows only for NaN pattern / finiteness + clamped log-det"""
    x = g["x"]
    return (x.abs() > 50).any(dim=1) | ~torch.isfinite(x).all(dim=1)


def _compare(g, y, ld, key, tag):
    ref_y, ref_ld = g[key], torch.as_tensor(g[key + "_ld"])
    y, ld = y.cpu(), ld.cpu()
    if ref_ld.dim() == 0:
        ref_ld = ref_ld.expand(ld.shape)
    wild = _loose(g)
    calm = ~wild
    assert_close(y[calm], ref_y[calm], Z_ATOL, Z_RTOL, f"{tag} {key} z")
    assert_close(ld[calm], ref_ld[calm], LD_ATOL, LD_RTOL, f"{tag} {key} log_det")
    if wild.any():
        assert torch.equal(torch.isfinite(y[wild]), torch.isfinite(ref_y[wild])), f"{tag} {key} finiteness (stress rows)"
        assert torch.equal(torch.isfinite(ld[wild]), torch.isfinite(ref_ld[wild]))


@pytest.mark.parametrize("name", MODULE_CASES)
def test_fused_route_matches_reference_golden(name):
    """no_grad + eval: single-launch kernels (stack / MADE chain / incremental sequential)."""
    g = G.load(name)
    m = build(g).to(_dev())
    x = g["x"].to(_dev())
    before = N._lib.launch_count()
    with torch.no_grad():
        for inverse, key in ((False, "fwd"), (True, "inv")):
            y, ld = _run(m, x, inverse)
            _compare(g, y, ld, key, name + " [fused]")
    assert N._lib.launch_count() > before


@pytest.mark.parametrize("name", MODULE_CASES)
def test_layered_route_matches_reference_golden(name):
    """grad enabled: conditioner GEMMs + BatchNorm + transform kernels, one autograd Function each."""
    g = G.load(name)
    m = build(g).to(_dev())
    x = g["x"].to(_dev()).requires_grad_()
    for inverse, key in ((False, "fwd"), (True, "inv")):
        y, ld = _run(m, x, inverse)
        assert y.requires_grad and ld.requires_grad
        _compare(g, y, ld, key, name + " [layered]")


@pytest.mark.parametrize("name", MODULE_CASES)
def test_float64_route_matches_float64_oracle(name):
    g = G.load(name)
    finite = torch.isfinite(g["x"]).all(dim=1) & (g["x"].abs() < 50).all(dim=1)
    x = g["x"][finite].double()
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in g["sd"].items()}
    g64 = dict(g, sd=sd64, x=x)
    m = build(g).double().to(_dev())
    with torch.no_grad():
        for inverse in (False, True):
            ry, rld = G.oracle_eval(g64, inverse)
            y, ld = _run(m, x.to(_dev()), inverse)
            assert y.dtype == torch.float64
            assert_close(y, ry, 1e-9, 1e-9, f"{name} f64 z inv={inverse}")
            assert_close(ld, torch.as_tensor(rld).expand(ld.shape), 1e-8, 1e-9, f"{name} f64 ld inv={inverse}")


@pytest.mark.parametrize("name", G.golden_names("coupling_train"))
def test_train_mode_coupling_batch_statistics(name):
    """BatchNorm batch statistics + running-stat side effects (coupling_layer.py:20,23)."""
    g = G.load(name)
    m = build(g).to(_dev())
    m.train()
    with torch.no_grad():
        y, ld = m.forward(g["x"].to(_dev()))
    assert_close(y, g["fwd"], Z_ATOL, Z_RTOL, name + " z")
    assert_close(ld, g["fwd_ld"], LD_ATOL, LD_RTOL, name + " ld")
    after = m.state_dict()
    for k, v in g["sd_after"].items():
        if v.is_floating_point():
            assert_close(after[k], v, 1e-6, 1e-5, f"{name} {k}")
        else:
            assert torch.equal(after[k].cpu(), v), k


def test_train_mode_between_layer_batchnorm():
    g = G.load("realnvp_4_4_16_bn_train")
    m = build(g).to(_dev())
    m.train()
    with torch.no_grad():
        y, ld = m.forward(g["x"].to(_dev()))
    assert_close(y, g["fwd"], Z_ATOL, Z_RTOL, "bn-train z")
    assert_close(ld, g["fwd_ld"], LD_ATOL, LD_RTOL, "bn-train ld")
    after = m.state_dict()
    for k, v in g["sd_after"].items():
        if v.is_floating_point():
            assert_close(after[k], v, 1e-6, 1e-5, k)
        else:
            assert torch.equal(after[k].cpu(), v), k


@pytest.mark.parametrize("name", G.golden_names("rqs_unit"))
def test_public_spline_function(name):
    g = G.load(name)
    d = _dev()
    for inverse, key in ((False, "fwd"), (True, "inv")):
        y, ld = N.rational_quadratic_spline(g["x"].to(d), g["w"].to(d), g["h"].to(d), g["d"].to(d), inverse=inverse)
        y64, l64 = O.rqs_unit(g["x"].double(), g["w"].double(), g["h"].double(), g["d"].double(), inverse)
        # judged against the float64 oracle with the reference's own fp32 error as slack
        e_y = (g[key].double() - y64).abs()
        e_l = (g[key + "_ld"].double() - l64).abs()
        assert bool(((y.cpu().double() - y64).abs() <= 2e-6 + 2e-6 * y64.abs() + 2 * e_y + 0.5 * e_y.max()).all())
        assert bool(((ld.cpu().double() - l64).abs() <= 1e-5 + 1e-5 * l64.abs() + 2 * e_l + 0.5 * e_l.max()).all())


@pytest.mark.parametrize("name", [n for n in G.golden_names("spline_") if n.endswith("_rqs")])
def test_bounded_spline_transform_kernel(name):
    """a6 through the C ABI: every element its own row (D=1), so the row log-det is the element log-det."""
    g = G.load(name)
    K = g["K"]
    d = _dev()
    x = g["x"].reshape(-1, 1).to(d)
    params = torch.cat([g["uw"], g["uh"], g["ud"]], dim=-1).reshape(-1, 3 * K - 1).contiguous().to(d)
    mask = torch.zeros(1, device=d)
    tidx = torch.zeros(1, dtype=torch.int32, device=d)
    for inverse, key in ((False, "fwd"), (True, "inv")):
        y, ld = N.ops.spline_transform(x, params, mask, tidx, K, inverse, g["bound"], (1e-3, 1e-3, 1e-3))
        y64, l64 = O.rqs_bounded(g["x"].double(), g["uw"].double(), g["uh"].double(), g["ud"].double(), inverse)
        e_y = (g[key].double() - y64).abs().reshape(-1)
        e_l = (g[key + "_ld"].double() - l64).abs().reshape(-1)
        y64, l64 = y64.reshape(-1), l64.reshape(-1)
        assert bool(((y.cpu().double().reshape(-1) - y64).abs() <= 5e-6 + 2e-6 * y64.abs() + 2 * e_y + 0.5 * e_y.max()).all())
        assert bool(((ld.cpu().double() - l64).abs() <= 1e-5 + 1e-5 * l64.abs() + 2 * e_l + 0.5 * e_l.max()).all())
