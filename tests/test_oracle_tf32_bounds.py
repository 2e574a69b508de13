"""CPU: the documented bounds of the reduced-precision dense-layer mode (set_gemm_precision("tf32"), DESIGN.md
"Reduced-precision mode") checked independently of the GPU -- the oracle is evaluated with every F.linear operand
rounded to the nearest TF32 (what the converter warps and nf_split_tf32 do: add 0x1000, clear the 13 low mantissa bits)
and compared with the plain fp32 oracle on the same weights and inputs."""
import types

import pytest
import torch
import torch.nn.functional as TF

from oracle import flows_oracle as O


def _round_tf32(t: torch.Tensor) -> torch.Tensor:
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def _tf32_linear(x, w, b=None):
    return TF.linear(_round_tf32(x), _round_tf32(w), b)


@pytest.fixture
def tf32_oracle(monkeypatch):
    """flows_oracle with F.linear replaced by the one-pass TF32 emulation (fp32 accumulation)."""
    shim = types.SimpleNamespace(**{k: getattr(TF, k) for k in dir(TF) if not k.startswith("__")})
    shim.linear = _tf32_linear
    monkeypatch.setattr(O, "F", shim)
    return O


def test_tf32_rounding_matches_device_split():
    """Same rounding as tc::split_tf32 / packing.split_tf32: nearest TF32, ties away from zero in magnitude bits."""
    import nfb200 as N
    x = torch.randn(4096) * 3
    hi, _ = N.packing.split_tf32(x.numpy())
    assert torch.equal(_round_tf32(x), torch.from_numpy(hi))
    rel = ((_round_tf32(x) - x).abs() / x.abs()).max().item()
    assert 0 < rel <= 2.0 ** -11


@pytest.mark.parametrize("case", ["maf64", "spline784"])
def test_documented_bounds_hold_for_emulated_tf32(case, tf32_oracle):
    torch.manual_seed(0)
    if case == "maf64":
        sd = O.init_made_sd(64, 512, seed=1, sigma=0.02, prefix="conditioner.")
        x = torch.randn(512, 64)
        run = lambda: O.maf_inverse(sd, "", x)
        z_tol, ld_tol = 5e-3, 5e-2
    else:
        sd = O.init_spline_stack_sd(784, 2, 1024, 10, seed=1, sigma=0.01)
        x = torch.randn(48, 784)
        specs = [{"kind": "spline", "num_bins": 10}] * 2
        run = lambda: O.flow_model(sd, "flow.", specs, x, inverse=True)
        z_tol, ld_tol = 2e-2, 2e-1
    z_t, ld_t = run()                                    # emulated one-pass TF32
    with pytest.MonkeyPatch.context() as mp:
        mp.setattr(O, "F", TF)
        z_r, ld_r = run()                                # plain fp32 oracle
    dz = ((z_t - z_r).abs() / z_r.abs().clamp_min(1)).max().item()
    dl = (ld_t - ld_r).abs().max().item()
    assert torch.isfinite(z_t).all() and torch.isfinite(ld_t).all()
    assert 0 < dz < z_tol, (case, dz)
    assert dl < ld_tol, (case, dl)
