"""Build the product's nn.Modules (nfb200) for a golden case and load the reference weights into them."""
import torch

import nfb200 as N


def _mask_keys(sd, prefix):
    i, out = 0, []
    while f"{prefix}flows.{i}.mask" in sd:
        out.append(sd[f"{prefix}flows.{i}.mask"].clone())
        i += 1
    return out


def _layer_from_spec(spec, D, H, sd, p):
    k = spec["kind"]
    if k == "coupling":
        return N.CouplingLayer(D, H, sd[p + "mask"].clone())
    if k == "spline":
        return N.SplineCouplingLayer(D, H, sd[p + "mask"].clone(), num_bins=spec.get("num_bins", 10))
    if k == "maf":
        return N.MaskedAutoregressiveFlow(D, H)
    if k == "iaf":
        return N.InverseAutoregressiveFlow(D, H)
    raise ValueError(k)


def build(g, sd=None):
    """Returns the module for golden case `g` (CPU, eval mode) with `sd` (default g['sd']) loaded strictly."""
    sd = g["sd"] if sd is None else sd
    k = g["kind"]
    if k in ("coupling", "coupling_train"):
        m = N.CouplingLayer(g["D"], g["H"], sd["mask"].clone())
    elif k == "spline":
        m = N.SplineCouplingLayer(g["D"], g["H"], sd["mask"].clone(), num_bins=g["K"], **g["extra"])
    elif k in ("maf", "maf_train"):
        m = N.MaskedAutoregressiveFlow(g["D"], g["H"], use_batch_norm=g.get("use_batch_norm", False))
    elif k in ("iaf", "iaf_train"):
        m = N.InverseAutoregressiveFlow(g["D"], g["H"], use_batch_norm=g.get("use_batch_norm", False))
    elif k == "arqs":
        m = N.ARQS(g["D"], hidden_dim=g["H"], num_bins=g["K"], **g["extra"])
    elif k in ("realnvp", "realnvp_train"):
        m = N.RealNVP(g["D"], g["L"], g["H"], batch_norm_between_layers=g["bn"])
    elif k == "realnvpspline":
        m = N.RealNVPSpline(g["D"], g["L"], g["H"], batch_norm_between_layers=g["bn"])
    elif k == "splinestack":
        m = N.NormalizingFlowModel([N.SplineCouplingLayer(g["D"], g["H"], mk, num_bins=g["K"])
                                    for mk in _mask_keys(sd, "")])
    elif k == "mixed":
        layers = [_layer_from_spec(s, g["D"], g["H"], sd, f"flows.{i}.") for i, s in enumerate(g["specs"])]
        m = N.NormalizingFlowModel(layers, batch_norm_between_layers=g["bn"])
    elif k == "sequential":
        layers = [_layer_from_spec(s, g["D"], g["H"], sd, f"flows.{i}.") for i, s in enumerate(g["specs"])]
        m = N.SequentialFlow(layers)
    else:
        raise ValueError(k)
    m.load_state_dict(sd, strict=True)
    m.eval()
    return m


MODULE_KINDS = ("coupling", "spline", "maf", "iaf", "realnvp", "realnvpspline", "splinestack", "mixed", "sequential",
                "arqs")


def assert_close(a, b, atol, rtol, what=""):
    """|a-b| <= atol + rtol*|b| elementwise, identical NaN pattern."""
    a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    assert torch.equal(torch.isnan(a), torch.isnan(b)), f"{what}: NaN pattern differs"
    m = ~torch.isnan(a)
    inf = torch.isinf(b) & m
    assert torch.equal(a[inf], b[inf]), f"{what}: Inf pattern differs"
    m = m & ~torch.isinf(b)
    err = (a[m] - b[m]).abs() - (atol + rtol * b[m].abs())
    if err.numel():
        assert bool((err <= 0).all()), f"{what}: max excess {err.max().item():.3e} (max abs diff {(a[m]-b[m]).abs().max().item():.3e})"
