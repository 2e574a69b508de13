"""GPU: the bf16 tensor-core mode (`set_gemm_precision("bf16")`): MAF.inverse / IAF.forward as ONE fused launch
(csrc/made_chain_bf16.cu).  north_star allows "documented looser bounds for any bf16 GEMM path"; they are asserted here:
  * against a torch emulation of the same arithmetic (operands and activations rounded to bf16, fp32 accumulation,
    fp32 transform): |dz| <= 2e-3 (1 + |z|), |d log_det| <= 5e-3 -- what is left is accumulation order;
  * against the fp32-parity route on the same weights / inputs: |dz| <= 2e-2 (1 + |z|), |d log_det| <= 1e-1 per row at
    D = 64, H = 512 with the bench's weight scale (SURVEY A.3 measured 2.4e-4 / 4.6e-4 for bf16 operands);
the measured values are printed."""
import pytest
import torch

import nfb200 as N

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _restore_precision():
    yield
    N.set_gemm_precision("fp32")


def _model(cls, D, H, sigma, seed):
    torch.manual_seed(seed)
    m = cls(D, H)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(sigma * torch.randn(p.shape, generator=g))
    return m.to(DEV).eval()


def _emulate(m, x, iaf):
    """Same arithmetic in torch: bf16-rounded operands / activations, fp32 accumulation, fp32 transform."""
    bf = lambda t: t.to(torch.bfloat16).float()
    f = m.conditioner.folded()
    h = bf(x)
    for i in range(3):
        h = bf(torch.relu(h @ bf(f.w[i]).T + f.b[i]))
    out = h @ bf(f.w[3]).T + f.b[3]
    D = x.shape[1]
    mu, al = out[:, :D], out[:, D:]
    if iaf:
        al = al.clamp(-2, 2)
        mu = mu.clamp(-10, 10)
        y = x * torch.exp(al.clamp(-3, 3)) + mu
        ld = al.sum(1).clamp(-50, 50)
    else:
        al = al.clamp(-3, 3)
        y = (x - mu) * torch.exp((-al).clamp(-5, 5))
        ld = (-al.sum(1)).clamp(-100, 100)
    return y, ld


CASES = [(64, 512, 0.02), (32, 256, 0.03), (16, 128, 0.05), (8, 128, 0.05), (64, 128, 0.03), (60, 384, 0.02)]


@pytest.mark.parametrize("iaf", [False, True])
@pytest.mark.parametrize("D,H,sigma", CASES)
def test_bf16_chain_matches_emulation_and_fp32_route(D, H, sigma, iaf):
    cls = N.InverseAutoregressiveFlow if iaf else N.MaskedAutoregressiveFlow
    m = _model(cls, D, H, sigma, 3)
    run = (lambda v: m.forward(v)) if iaf else (lambda v: m.inverse(v))
    for B in (128, 1000, 20011):
        x = torch.randn(B, D, device=DEV) * 1.5
        with torch.no_grad():
            y32, ld32 = run(x)
            N.set_gemm_precision("bf16")
            before = N._lib.launch_count()
            y, ld = run(x)
            assert N._lib.launch_count() - before == 1, "the bf16 mode must run the whole direction in one launch"
            N.set_gemm_precision("fp32")
            ye, lde = _emulate(m, x, iaf)
        e_emu = ((y - ye).abs() / (1 + ye.abs())).max().item()
        l_emu = (ld - lde).abs().max().item()
        e_32 = ((y - y32).abs() / (1 + y32.abs())).max().item()
        l_32 = (ld - ld32).abs().max().item()
        print(f"[bf16] {'IAF' if iaf else 'MAF'}({D},{H}) B={B}: vs emulation z {e_emu:.2e} ld {l_emu:.2e}; vs fp32 route z {e_32:.2e} "
              f"(mean {((y - y32).abs() / (1 + y32.abs())).mean().item():.2e}) ld {l_32:.2e} (mean {(ld - ld32).abs().mean().item():.2e})")
        assert bool(torch.isfinite(y).all()) and bool(torch.isfinite(ld).all())
        assert e_emu <= 2e-3 and l_emu <= 5e-3, (e_emu, l_emu)
        assert e_32 <= 2e-2 and l_32 <= 1e-1, (e_32, l_32)


def test_bf16_chain_log_prob_head_and_row_independence():
    m = _model(N.MaskedAutoregressiveFlow, 64, 512, 0.02, 5)
    x = torch.randn(5000, 64, device=DEV)
    N.set_gemm_precision("bf16")
    with torch.no_grad():
        z, ld = m.inverse(x)
        before = N._lib.launch_count()
        lp = m.log_prob(x)
        assert N._lib.launch_count() - before == 1
        want = (-0.5 * z * z).sum(1) - 32 * 1.8378770664093453 + ld
        assert torch.allclose(lp, want, rtol=1e-5, atol=2e-4), (lp - want).abs().max().item()
        # rows are independent and launches deterministic: two uneven shards give bit-identical rows
        z1, ld1 = m.inverse(x[:1777].contiguous())
        z2, ld2 = m.inverse(x[1777:].contiguous())
        assert torch.equal(torch.cat([z1, z2]), z) and torch.equal(torch.cat([ld1, ld2]), ld)
        z3, ld3 = m.inverse(x)
        assert torch.equal(z3, z) and torch.equal(ld3, ld)


def test_bf16_chain_non_finite_rows_are_scrubbed_and_do_not_leak():
    m = _model(N.MaskedAutoregressiveFlow, 64, 512, 0.02, 7)
    x = torch.randn(600, 64, device=DEV)
    bad = x.clone()
    bad[5, 40] = float("nan")
    bad[130, 3] = float("inf")
    bad[599, 63] = float("-inf")
    N.set_gemm_precision("bf16")
    with torch.no_grad():
        z, ld = m.inverse(x)
        zb, ldb = m.inverse(bad)
    rows = torch.tensor([5, 130, 599], device=DEV)
    keep = torch.ones(600, dtype=torch.bool, device=DEV)
    keep[rows] = False
    assert torch.equal(z[keep], zb[keep]) and torch.equal(ld[keep], ldb[keep])
    assert bool(torch.isfinite(zb).all()) and bool(torch.isfinite(ldb).all())          # NaN/Inf -> 0 (:41-42)


def test_bf16_mode_falls_back_outside_the_envelope():
    """data_dim not a multiple of 4 / hidden_dim not a multiple of 128: the mode silently uses the GEMM chain."""
    m = _model(N.MaskedAutoregressiveFlow, 10, 24, 0.1, 9)
    x = torch.randn(300, 10, device=DEV)
    with torch.no_grad():
        z32, ld32 = m.inverse(x)
        N.set_gemm_precision("bf16")
        z, ld = m.inverse(x)
    assert torch.allclose(z, z32, rtol=1e-2, atol=1e-2) and torch.allclose(ld, ld32, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("kind", ["realnvp256", "spline64"])
def test_bf16_mode_wide_coupling_layers_within_documented_bounds(kind):
    """Outside the MADE flows the bf16 mode runs one TF32 pass with the x operand read straight from shared memory (the
    tensor core truncates it to TF32: 10-bit mantissa, error <= 2^-10 |x| -- inside what bf16 operands would give).
    Documented bounds against the fp32-parity route: |dz| <= 2e-2 (1 + |z|), |d log_det| <= 2e-1 per row."""
    torch.manual_seed(0)
    if kind == "realnvp256":
        m, D = N.RealNVP(256, 2, 512), 256
    else:
        m, D = N.RealNVPSpline(64, 2, 256), 64
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.02 * torch.randn_like(p))
    m = m.to(DEV).eval()
    x = torch.randn(4096, D, device=DEV)
    with torch.no_grad():
        z32, ld32 = m.inverse(x)
        N.set_gemm_precision("bf16")
        z, ld = m.inverse(x)
    ez = ((z - z32).abs() / (1 + z32.abs())).max().item()
    el = (ld - ld32).abs().max().item()
    print(f"[bf16] {kind}: vs fp32 route z {ez:.2e} ld {el:.2e} (mean {(ld - ld32).abs().mean().item():.2e})")
    assert ez > 0 and ez <= 2e-2 and el <= 2e-1, (ez, el)
