"""CPU oracle for the normalizing-flow transform hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`normalizing-flows-study_b200/`) may import this module; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs use it, and only as the checker / the timed CPU baseline.

What it is: a functional restatement, on CPU `torch` tensors, of the arithmetic
the reference (itxtx/normalizing-flows-study) performs on its forward / inverse
+ log-det path.  The reference's arithmetic lives entirely in ATen eager ops
(torch>=2.2, `pyproject.toml:20`), so the oracle issues the same ATen ops in
the same order; it is organised as stateless functions over a *state_dict with
the reference's key layout* instead of `nn.Module`s, so one set of weights can
be fed to the reference, the oracle and the CUDA product.

Parity pin: `tests/golden/*.pt` were produced by importing the *unmodified*
reference from /root/reference (`tests/golden/make_golden.py`);
`tests/test_oracle_golden.py` checks every function below against them.

Citations are `path:line` relative to the reference repository root.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

BN_EPS = 1e-5        # nn.BatchNorm1d default
BN_MOMENTUM = 0.1    # nn.BatchNorm1d default


# --------------------------------------------------------------------------- #
# helpers
# --------------------------------------------------------------------------- #
def _scrub_zero(t: Tensor) -> Tensor:
    """NaN/Inf -> 0  (coupling_layer.py:61-66, spline_coupling_layer.py:130-135)."""
    return torch.where(torch.isnan(t) | torch.isinf(t), torch.zeros_like(t), t)


def _bn1d(sd: SD, p: str, h: Tensor, training: bool, update: bool) -> Tensor:
    """nn.BatchNorm1d inside the coupling conditioners (coupling_layer.py:20,23).

    train: biased batch variance normalises, running stats move with momentum
    0.1 using the unbiased variance (torch defaults); eval: running stats.
    `update=False` leaves `sd` untouched in train mode (pure function)."""
    rm, rv = sd[p + ".running_mean"], sd[p + ".running_var"]
    if training and not update:
        rm, rv = rm.clone(), rv.clone()
    out = F.batch_norm(h, rm, rv, sd[p + ".weight"], sd[p + ".bias"],
                       training, BN_MOMENTUM, BN_EPS)
    if training and update and (p + ".num_batches_tracked") in sd:
        sd[p + ".num_batches_tracked"] += 1
    return out


# --------------------------------------------------------------------------- #
# a1/a2  affine coupling  (src/flows/coupling/coupling_layer.py:40-96)
# --------------------------------------------------------------------------- #
def coupling_conditioner(sd: SD, p: str, x_a: Tensor, training=False, update=False) -> Tensor:
    """One of s_net / b_net: Linear-BN-ReLU-Linear-BN-ReLU-Linear (coupling_layer.py:18-35)."""
    h = F.linear(x_a, sd[p + ".0.weight"], sd[p + ".0.bias"])
    h = torch.relu(_bn1d(sd, p + ".1", h, training, update))
    h = F.linear(h, sd[p + ".3.weight"], sd[p + ".3.bias"])
    h = torch.relu(_bn1d(sd, p + ".4", h, training, update))
    return F.linear(h, sd[p + ".6.weight"], sd[p + ".6.bias"])


def affine_coupling(sd: SD, p: str, x: Tensor, inverse: bool,
                    training=False, update=False) -> Tuple[Tensor, Tensor]:
    """CouplingLayer.forward (:40-68) / .inverse (:70-96).  `p` is the layer's
    key prefix including the trailing dot ('' for a bare layer)."""
    m = sd[p + "mask"]
    x_a = x * m
    s = torch.clamp(coupling_conditioner(sd, p + "s_net", x_a, training, update), -10.0, 10.0)
    b = torch.clamp(coupling_conditioner(sd, p + "b_net", x_a, training, update), -10.0, 10.0)
    if not inverse:
        y = x_a + (1 - m) * (x * torch.exp(s) + b)
        ld = ((1 - m) * s).sum(dim=1)
    else:
        y = x_a + (1 - m) * ((x - b) * torch.exp(-s))
        ld = ((1 - m) * -s).sum(dim=1)
    return _scrub_zero(y), _scrub_zero(ld)


# --------------------------------------------------------------------------- #
# a6  bounded rational-quadratic spline with identity tails
#     (src/flows/spline/spline_coupling_layer.py:182-309)
# --------------------------------------------------------------------------- #
def _knots_pm_bound(unnorm: Tensor, K: int, floor: float, bound: float, eps: float):
    """softmax -> floor -> cumsum -> scale to [-B,B] -> pin ends -> re-difference
    (spline_coupling_layer.py:204-215 for widths, :217-228 for heights)."""
    w = F.softmax(unnorm, dim=-1)
    w = floor + (1 - floor * K) * w
    w = torch.clamp(w, min=eps)
    cw = F.pad(torch.cumsum(w, dim=-1), pad=(1, 0), mode="constant", value=0.0)
    cw = (2 * bound) * cw + (-bound)
    cw[..., 0] = -bound
    cw[..., -1] = bound
    w = torch.clamp(cw[..., 1:] - cw[..., :-1], min=eps)
    return w, cw


def rqs_bounded(inputs: Tensor, uw: Tensor, uh: Tensor, ud: Tensor, inverse: bool,
                bound=5.0, min_bin_width=1e-3, min_bin_height=1e-3,
                min_derivative=1e-3) -> Tuple[Tensor, Tensor]:
    """SplineCouplingLayer._rational_quadratic_spline.  inputs [B,Dt]; uw,uh
    [B,Dt,K]; ud [B,Dt,K-1].  Returns per-element outputs and log|dy/dx|."""
    eps = 1e-8                                                     # :189
    K = uw.shape[-1]
    inside = (inputs >= -bound) & (inputs <= bound)                # :192
    outputs = torch.where(~inside, inputs, torch.zeros_like(inputs))
    logabsdet = torch.zeros_like(inputs)
    if not inside.any():                                           # :200-201
        return outputs, logabsdet

    w, cw = _knots_pm_bound(uw, K, min_bin_width, bound, eps)
    h, ch = _knots_pm_bound(uh, K, min_bin_height, bound, eps)
    d = torch.clamp(min_derivative + F.softplus(ud), min=eps)      # :230-231
    d = F.pad(d, pad=(1, 1), mode="constant", value=1.0)           # :232

    flat = inputs.contiguous().view(-1)
    knots = (ch if inverse else cw).contiguous().view(-1, K + 1)   # :236-239
    k = torch.searchsorted(knots, flat.unsqueeze(-1), right=True).squeeze(-1) - 1
    k = torch.clamp(k, 0, K - 1)                                   # :244

    def pick(t, idx):
        return torch.gather(t.contiguous().view(-1, t.shape[-1]), 1, idx.unsqueeze(-1)).squeeze(-1)

    w_k, x_k = pick(w, k), pick(cw, k)
    h_k, y_k = pick(h, k), pick(ch, k)
    d_k = pick(d, k)
    d_k1 = pick(d, (k + 1).clamp(max=K))
    s_k = h_k / torch.clamp(w_k, min=eps)                          # :260

    if inverse:                                                    # :264-281
        y = flat
        t = (y - y_k) * (d_k + d_k1 - 2 * s_k)
        a = t + h_k * (s_k - d_k)
        b = h_k * d_k - t
        c = -s_k * (y - y_k)
        disc = torch.clamp(b.pow(2) - 4 * a * c, min=0.0)
        q = -b - torch.sqrt(disc)
        q = torch.where(q.abs() < eps, torch.full_like(q, eps), q)
        xi = torch.clamp((2 * c) / q, 0, 1)
        out = xi * w_k + x_k
        den = s_k + (d_k1 + d_k - 2 * s_k) * xi * (1 - xi)
        num = s_k.pow(2) * (d_k1 * xi.pow(2) + 2 * s_k * xi * (1 - xi) + d_k * (1 - xi).pow(2))
        lad = -torch.log(torch.clamp(num, min=eps)) + 2 * torch.log(torch.clamp(den, min=eps))
    else:                                                          # :283-293
        xi = torch.clamp((flat - x_k) / torch.clamp(w_k, min=eps), 0, 1)
        den = torch.clamp(s_k + (d_k1 + d_k - 2 * s_k) * xi * (1 - xi), min=eps)
        out = y_k + h_k * (s_k * xi.pow(2) + d_k * xi * (1 - xi)) / den
        num = s_k.pow(2) * (d_k1 * xi.pow(2) + 2 * s_k * xi * (1 - xi) + d_k * (1 - xi).pow(2))
        der = num / torch.clamp(den.pow(2), min=eps)
        lad = torch.log(torch.clamp(der, min=eps))

    sel = inside.view(-1)                                          # :296-303
    outputs = outputs.clone().view(-1)
    logabsdet = logabsdet.clone().view(-1)
    outputs[sel] = out[sel]
    logabsdet[sel] = lad[sel]
    outputs = outputs.view_as(inputs)
    logabsdet = logabsdet.view_as(inputs)
    outputs = torch.where(torch.isnan(outputs) | torch.isinf(outputs), inputs, outputs)   # :306
    logabsdet = _scrub_zero(logabsdet)                                                    # :307
    return outputs, logabsdet


# --------------------------------------------------------------------------- #
# a4/a5  spline coupling layer (spline_coupling_layer.py:66-180)
# --------------------------------------------------------------------------- #
def spline_coupling(sd: SD, p: str, x: Tensor, inverse: bool, num_bins=10, bound=5.0,
                    min_bin_width=1e-3, min_bin_height=1e-3, min_derivative=1e-3,
                    data_min=None, data_max=None) -> Tuple[Tensor, Tensor]:
    m = sd[p + "mask"]
    D = m.numel()
    K = num_bins
    rescale = data_min is not None and data_max is not None
    xr = (2 * bound) / (data_max - data_min) * (x - data_min) - bound if rescale else x   # :78-85
    x_a = xr * m
    h = torch.relu(F.linear(x_a, sd[p + "param_net.0.weight"], sd[p + "param_net.0.bias"]))
    h = torch.relu(F.linear(h, sd[p + "param_net.2.weight"], sd[p + "param_net.2.bias"]))
    params = F.linear(h, sd[p + "param_net.4.weight"], sd[p + "param_net.4.bias"])
    params = params.view(-1, D, 3 * K - 1)                          # :71
    uw, uh, ud = torch.split(params, [K, K, K - 1], dim=-1)
    tr = m == 0                                                     # :108-112
    out_b, ld_b = rqs_bounded(xr[:, tr], uw[:, tr], uh[:, tr], ud[:, tr], inverse,
                              bound, min_bin_width, min_bin_height, min_derivative)
    if rescale:                                                     # :87-94
        out_b = (out_b + bound) * ((data_max - data_min) / (2 * bound)) + data_min
    y = x.clone()
    y[:, tr] = out_b
    return _scrub_zero(y), _scrub_zero(ld_b.sum(dim=1))


# --------------------------------------------------------------------------- #
# a7  public stand-alone spline on [0,1]
#     (src/flows/spline/rational_quadratic_spline.py:4-104)
# --------------------------------------------------------------------------- #
def rqs_unit(inputs: Tensor, widths: Tensor, heights: Tensor, derivatives: Tensor,
             inverse=False, min_bin_width=1e-3, min_bin_height=1e-3,
             min_derivative=1e-3) -> Tuple[Tensor, Tensor]:
    eps = 1e-6                                                      # :19 (argument is overwritten)
    K = widths.shape[-1]
    w = torch.clamp(min_bin_width + (1 - min_bin_width * K) * F.softmax(widths, dim=-1), min=eps)
    h = torch.clamp(min_bin_height + (1 - min_bin_height * K) * F.softmax(heights, dim=-1), min=eps)
    d = torch.clamp(F.softplus(derivatives) + min_derivative, min=eps)
    xk = F.pad(torch.cumsum(w, dim=-1), (1, 0), "constant", 0.0)    # :36-37
    yk = F.pad(torch.cumsum(h, dim=-1), (1, 0), "constant", 0.0)
    d = F.pad(d, (1, 1), "constant", 1.0)                           # :40
    knots = (yk if inverse else xk).contiguous()
    if knots.dim() > 2:
        knots = knots.view(-1, knots.shape[-1])
    k = torch.searchsorted(knots, inputs.unsqueeze(-1), right=True) - 1
    k = torch.clamp(k, 0, K - 1)                                    # :57
    x_k, y_k = torch.gather(xk, -1, k), torch.gather(yk, -1, k)
    w_k, h_k = torch.gather(w, -1, k), torch.gather(h, -1, k)
    d_k, d_k1 = torch.gather(d, -1, k), torch.gather(d, -1, k + 1)
    s_k = h_k / torch.clamp(w_k, min=eps)
    v = inputs.unsqueeze(-1)
    if inverse:                                                     # :72-87
        t = (v - y_k) * (d_k + d_k1 - 2 * s_k)
        a = h_k * (s_k - d_k) + t
        b = h_k * d_k - t
        c = -s_k * (v - y_k)
        disc = torch.clamp(b.pow(2) - 4 * a * c, min=0)
        th = torch.clamp((2 * c) / (-b - torch.sqrt(disc)), 0, 1)
        out = th * w_k + x_k
        tt = th * (1 - th)
        num = s_k.pow(2) * (d_k1 * th.pow(2) + 2 * s_k * tt + d_k * (1 - th).pow(2))
        den = (s_k + (d_k + d_k1 - 2 * s_k) * tt).pow(2)
        ld = -torch.log(torch.clamp(num / torch.clamp(den, min=eps), min=eps))
    else:                                                           # :89-102
        th = torch.clamp((v - x_k) / torch.clamp(w_k, min=eps), 0, 1)
        tt = th * (1 - th)
        den = s_k + (d_k + d_k1 - 2 * s_k) * tt
        out = y_k + h_k * (s_k * th.pow(2) + d_k * tt) / torch.clamp(den, min=eps)
        num = s_k.pow(2) * (d_k1 * th.pow(2) + 2 * s_k * tt + d_k * (1 - th).pow(2))
        ld = torch.log(torch.clamp(num / torch.clamp(den.pow(2), min=eps), min=eps))
    return out.squeeze(-1), ld.squeeze(-1)


# --------------------------------------------------------------------------- #
# a8-a10  MADE  (src/flows/autoregressive/made.py:12-140, masked_linear.py:14-18)
# --------------------------------------------------------------------------- #
def made_degrees(D: int, H: int) -> Tuple[np.ndarray, np.ndarray]:
    """Input degrees 0..D-1 and hidden degrees (made.py:25-41)."""
    m_in = np.arange(D)
    if D > 1:
        if D == 2:
            m_h = np.array([0, 0, 1, 1] * (H // 4 + 1))[:H]
        else:
            m_h = np.floor(np.linspace(0, D - 1, H)).astype(int)
    else:
        m_h = np.zeros(H, dtype=int)
    return m_in, m_h


def made_masks(D: int, H: int, mult: int = 2):
    """[in->h, h->h (shared), h->out] float32 masks (made.py:47-79)."""
    m_in, m_h = made_degrees(D, H)
    m1 = (m_in[:, None] <= m_h[None, :]).T
    mhh = (m_h[:, None] <= m_h[None, :]).T
    m2 = np.zeros((D * mult, H), dtype=np.float32)
    for kk in range(mult):
        for i in range(D):
            m2[kk * D + i, :] = (m_h < m_in[i]).astype(np.float32)
    return [torch.from_numpy(m1.astype(np.float32)), torch.from_numpy(mhh.astype(np.float32)),
            torch.from_numpy(m2)]


def made(sd: SD, p: str, x: Tensor, training=False, update=False) -> Tensor:
    """MADE.forward: 4 masked linears, 3 ReLU (made.py:81-140).  use_batch_norm=False: keys net.{0,2,4,6}.*;
    use_batch_norm=True: a BatchNorm1d after each of the first three masked linears (made.py:93-108), keys
    net.{0,3,6,9}.* for the linears and net.{1,4,7}.* for the BatchNorms -- detected from the state_dict."""
    bn = f"{p}net.1.running_mean" in sd
    idxs = (0, 3, 6, 9) if bn else (0, 2, 4, 6)
    h = x
    for i, idx in enumerate(idxs):
        W = sd[f"{p}net.{idx}.weight"]
        mk = sd[f"{p}net.{idx}.mask"].to(dtype=W.dtype)             # masked_linear.py:17
        h = F.linear(h, W * mk, sd[f"{p}net.{idx}.bias"])           # masked_linear.py:18
        if i < 3:
            if bn:
                h = _bn1d(sd, f"{p}net.{idx + 1}", h, training, update)
            h = torch.relu(h)
    return h


# --------------------------------------------------------------------------- #
# a11/a12  MAF  (masked_autoregressive_flow.py:18-78)
# --------------------------------------------------------------------------- #
def maf_inverse(sd: SD, p: str, x: Tensor, training=False, update=False) -> Tuple[Tensor, Tensor]:
    mu, alpha = made(sd, p + "conditioner.", x, training, update).chunk(2, dim=1)
    alpha = torch.clamp(alpha, min=-3, max=3)
    z = (x - mu) * torch.exp(torch.clamp(-alpha, min=-5, max=5))
    ld = _scrub_zero(-torch.sum(alpha, dim=1))
    return _scrub_zero(z), torch.clamp(ld, min=-100, max=100)


def maf_forward(sd: SD, p: str, z: Tensor) -> Tuple[Tensor, Tensor]:
    """D sequential full-MADE evaluations (:55-67)."""
    B, D = z.shape
    x = torch.zeros(B, D, dtype=z.dtype)
    ld = torch.zeros(B, dtype=z.dtype)
    for i in range(D):
        mu, alpha = made(sd, p + "conditioner.", x).chunk(2, dim=1)
        alpha = torch.clamp(alpha, min=-3, max=3)
        x_new = x.clone()
        x_new[:, i] = z[:, i] * torch.exp(torch.clamp(alpha[:, i], min=-5, max=5)) + mu[:, i]
        x = x_new
        ld += alpha[:, i]
    return _scrub_zero(x), torch.clamp(_scrub_zero(ld), min=-100, max=100)


# --------------------------------------------------------------------------- #
# a13  IAF  (inverse_autoregressive_flow.py:30-103)
# --------------------------------------------------------------------------- #
def iaf_forward(sd: SD, p: str, z: Tensor, training=False, update=False) -> Tuple[Tensor, Tensor]:
    mu, alpha = made(sd, p + "conditioner.", z, training, update).chunk(2, dim=1)
    alpha = torch.clamp(alpha, min=-2, max=2)
    mu = torch.clamp(mu, min=-10, max=10)
    x = z * torch.exp(torch.clamp(alpha, min=-3, max=3)) + mu
    ld = torch.sum(alpha, dim=1)
    x = torch.where(torch.isnan(x) | torch.isinf(x), z, x)         # scrub -> input (:53)
    return x, torch.clamp(_scrub_zero(ld), min=-50, max=50)


def iaf_inverse(sd: SD, p: str, x: Tensor) -> Tuple[Tensor, Tensor]:
    B, D = x.shape
    z = torch.zeros(B, D, dtype=x.dtype)
    ld = torch.zeros(B, dtype=x.dtype)
    for i in range(D):
        mu, alpha = made(sd, p + "conditioner.", z).chunk(2, dim=1)
        alpha = torch.clamp(alpha, min=-2, max=2)
        mu = torch.clamp(mu, min=-10, max=10)
        z_new = z.clone()
        z_new[:, i] = (x[:, i] - mu[:, i]) * torch.exp(torch.clamp(-alpha[:, i], min=-3, max=3))
        z = z_new
        ld -= alpha[:, i]
    z = torch.where(torch.isnan(z) | torch.isinf(z), x, z)         # :93
    return z, torch.clamp(_scrub_zero(ld), min=-50, max=50)


# --------------------------------------------------------------------------- #
# (f2)  ARQS  (src/flows/spline/arqs.py:7-114): MADE conditioner with 3K-1 outputs per dim, viewed [B, D, 3K-1];
#       BOTH directions are D-step sequential loops that re-evaluate the conditioner on the partially filled OUTPUT
#       and apply the public [0,1] spline (rqs_unit) to one column; log-dets accumulate in a float32 vector.
# --------------------------------------------------------------------------- #
def arqs(sd: SD, p: str, v: Tensor, inverse: bool, num_bins=8, data_min=None, data_max=None) -> Tuple[Tensor, Tensor]:
    K = num_bins
    P = 3 * K - 1
    rescale = data_min is not None and data_max is not None
    vr = (v - data_min) / (data_max - data_min) if rescale else v          # :28-34
    cur = torch.zeros_like(vr)                                               # :51 / :91
    ld = torch.zeros(v.size(0))                                              # :52 / :92 (float32 whatever v's dtype)
    B, D = v.shape
    for i in range(D):
        params = made(sd, p + "conditioner.", cur).view(B, D, P)            # :56-59
        out_i, ld_i = rqs_unit(vr[:, i], params[:, i, :K], params[:, i, K:2 * K], params[:, i, 2 * K:],
                               inverse=inverse)                              # :61-70 / :101-110
        new = cur.clone()
        new[:, i] = out_i
        cur = new
        ld += ld_i
    out = cur * (data_max - data_min) + data_min if rescale else cur         # :36-42
    return out, ld


# --------------------------------------------------------------------------- #
# a14-a16  stacks (normalizing_flow_model.py:25-128, sequential_flow.py:15-34)
# --------------------------------------------------------------------------- #
def layer_apply(sd: SD, p: str, spec: dict, x: Tensor, inverse: bool,
                training=False, update=False) -> Tuple[Tensor, Tensor]:
    """Dispatch one layer.  spec = {'kind': 'coupling'|'spline'|'maf'|'iaf', ...ctor kwargs}."""
    kind = spec["kind"]
    if kind == "coupling":
        return affine_coupling(sd, p, x, inverse, training, update)
    if kind == "spline":
        kw = {k: v for k, v in spec.items() if k != "kind"}
        return spline_coupling(sd, p, x, inverse, **kw)
    if kind == "maf":
        return maf_inverse(sd, p, x) if inverse else maf_forward(sd, p, x)
    if kind == "iaf":
        return iaf_inverse(sd, p, x) if inverse else iaf_forward(sd, p, x)
    raise ValueError(kind)


def _bn_between_logdet(sd: SD, p: str) -> Tensor:
    """Scalar sum(log|gamma| - 0.5 log(var+eps))  (normalizing_flow_model.py:87-108)."""
    return (torch.log(torch.abs(sd[p + ".weight"])) - 0.5 * torch.log(sd[p + ".running_var"] + BN_EPS)).sum()


def flow_model(sd: SD, p: str, specs: Sequence[dict], x: Tensor, inverse: bool,
               bn_between=False, training=False, update=False) -> Tuple[Tensor, Tensor]:
    """NormalizingFlowModel.forward (:25-46) / .inverse (:48-65).  `p` is the
    prefix of the model ('flow.' for RealNVP/RealNVPSpline, '' for a bare model)."""
    L = len(specs)
    total = 0
    if not inverse:
        for i, spec in enumerate(specs):
            x, ld = layer_apply(sd, f"{p}flows.{i}.", spec, x, False, training, update)
            total = total + ld
            if bn_between and i < L - 1:
                q = f"{p}batch_norms.{i}"
                if training and update:                              # :74-79
                    sd[q + ".running_mean"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * x.mean(dim=0))
                    sd[q + ".running_var"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * x.var(dim=0, unbiased=False))
                x = (x - sd[q + ".running_mean"].view(1, -1)) / torch.sqrt(sd[q + ".running_var"].view(1, -1) + BN_EPS) \
                    * sd[q + ".weight"].view(1, -1) + sd[q + ".bias"].view(1, -1)
                total = total + _bn_between_logdet(sd, q)
    else:
        for i in reversed(range(L)):
            if bn_between and i < L - 1:
                q = f"{p}batch_norms.{i}"
                x = (x - sd[q + ".bias"].view(1, -1)) / sd[q + ".weight"].view(1, -1) \
                    * torch.sqrt(sd[q + ".running_var"].view(1, -1) + BN_EPS) + sd[q + ".running_mean"].view(1, -1)
                total = total - _bn_between_logdet(sd, q)
            x, ld = layer_apply(sd, f"{p}flows.{i}.", specs[i], x, True, training, update)
            total = total + ld
    return x, total


def sequential_flow(sd: SD, p: str, specs: Sequence[dict], x: Tensor, inverse: bool) -> Tuple[Tensor, Tensor]:
    """SequentialFlow (sequential_flow.py:15-34): float32 zeros accumulator, no BN."""
    total = torch.zeros(x.size(0))
    order = reversed(range(len(specs))) if inverse else range(len(specs))
    for i in order:
        x, ld = layer_apply(sd, f"{p}flows.{i}.", specs[i], x, inverse)
        total += ld
    return x, total


def std_normal_log_prob(z: Tensor) -> Tensor:
    """Flow.log_prob's base term for N(0,I) (flow.py:67-71): sum_d -z^2/2 - D/2 log 2pi."""
    return (-0.5 * z * z).sum(dim=1) - 0.5 * z.shape[1] * math.log(2 * math.pi)


# --------------------------------------------------------------------------- #
# model builders shared by tests / bench (spec lists + reference-layout inits)
# --------------------------------------------------------------------------- #
def realnvp_masks(D: int, L: int):
    """real_nvp.py:27-31 (identical pattern in real_nvp_spline.py:24-31)."""
    out = []
    for i in range(L):
        m = torch.zeros(D)
        if i % 2 == 0:
            m[: D // 2] = 1
        else:
            m[D // 2:] = 1
        out.append(m)
    return out


def init_spline_stack_sd(D: int, L: int, H: int, K: int, seed=0, sigma=0.05, prefix="flow.") -> SD:
    """State dict with the reference's key layout for L SplineCouplingLayers
    (RealNVPSpline masks): xavier-normal hidden layers, zero bias; the final
    layer (zero in the reference, spline_coupling_layer.py:319-323) gets
    N(0, sigma^2) so no layer is the identity."""
    g = torch.Generator().manual_seed(seed)
    P = 3 * K - 1
    sd: SD = {}
    for i, m in enumerate(realnvp_masks(D, L)):
        q = f"{prefix}flows.{i}."
        sd[q + "mask"] = m
        for idx, (o, n) in zip((0, 2, 4), ((H, D), (H, H), (D * P, H))):
            std = math.sqrt(2.0 / (o + n)) if idx != 4 else sigma
            sd[q + f"param_net.{idx}.weight"] = torch.randn(o, n, generator=g) * std
            sd[q + f"param_net.{idx}.bias"] = torch.randn(o, generator=g) * sigma
    return sd


def init_made_sd(D: int, H: int, seed=0, sigma=0.02, prefix="conditioner.", mult=2) -> SD:
    """MADE weights: xavier gain 0.5 hidden, N(0,0.01^2) final (made.py:117-132) + sigma noise."""
    g = torch.Generator().manual_seed(seed)
    masks = made_masks(D, H, mult)
    shapes = ((H, D), (H, H), (H, H), (D * mult, H))
    sd: SD = {}
    for j, (idx, (o, n)) in enumerate(zip((0, 2, 4, 6), shapes)):
        std = 0.5 * math.sqrt(2.0 / (o + n)) if idx != 6 else 0.01
        sd[f"{prefix}net.{idx}.weight"] = torch.randn(o, n, generator=g) * std + torch.randn(o, n, generator=g) * sigma
        sd[f"{prefix}net.{idx}.bias"] = torch.randn(o, generator=g) * sigma
        sd[f"{prefix}net.{idx}.mask"] = masks[min(j, 1) if j < 3 else 2]
    return sd


def init_coupling_stack_sd(D: int, L: int, H: int, seed=0, sigma=0.05, prefix="flow.", bn_between=False) -> SD:
    """RealNVP state dict (real_nvp.py + coupling_layer.py:18-35,98-111) with perturbed
    final layers and non-trivial BN statistics."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for i, m in enumerate(realnvp_masks(D, L)):
        q = f"{prefix}flows.{i}."
        sd[q + "mask"] = m
        for net in ("s_net", "b_net"):
            for idx, (o, n) in zip((0, 3, 6), ((H, D), (H, H), (D, H))):
                std = math.sqrt(2.0 / (o + n)) if idx != 6 else sigma
                sd[q + f"{net}.{idx}.weight"] = torch.randn(o, n, generator=g) * std
                sd[q + f"{net}.{idx}.bias"] = torch.randn(o, generator=g) * sigma
            for idx in (1, 4):
                sd[q + f"{net}.{idx}.weight"] = 1 + 0.1 * torch.randn(H, generator=g)
                sd[q + f"{net}.{idx}.bias"] = 0.1 * torch.randn(H, generator=g)
                sd[q + f"{net}.{idx}.running_mean"] = 0.1 * torch.randn(H, generator=g)
                sd[q + f"{net}.{idx}.running_var"] = 1 + 0.2 * torch.rand(H, generator=g)
                sd[q + f"{net}.{idx}.num_batches_tracked"] = torch.tensor(0)
    if bn_between:
        for i in range(L):
            q = f"{prefix}batch_norms.{i}"
            sd[q + ".weight"] = 1 + 0.1 * torch.randn(D, generator=g)
            sd[q + ".bias"] = 0.1 * torch.randn(D, generator=g)
            sd[q + ".running_mean"] = 0.1 * torch.randn(D, generator=g)
            sd[q + ".running_var"] = 1 + 0.2 * torch.rand(D, generator=g)
            sd[q + ".num_batches_tracked"] = torch.tensor(0)
    return sd
