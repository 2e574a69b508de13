"""Layer stacks: same classes, constructor arguments and state_dict layout as the reference's `src.models`.

    NormalizingFlowModel   src/models/normalizing_flow_model.py:4-128
    RealNVP                src/models/real_nvp.py:6-49
    RealNVPSpline          src/models/real_nvp_spline.py:6-48
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .flows import ChainPlan, CouplingLayer, SplineCouplingLayer, _default_base, flow_log_prob, on_input_device


class _DensityAPI:
    """log_prob / sample on the model containers (north_star: `NormalizingFlowModel.log_prob/sample`; same meaning as
    Flow.log_prob / Flow.sample, flow.py:40-73).  base_dist=None means the standard normal N(0, I); a standard-normal
    base is evaluated by the fused head (one launch for a fused stack, z never written)."""

    def _data_dim(self):
        flows = self.flows if hasattr(self, "flows") else self.flow.flows
        return getattr(flows[0], "data_dim", None)

    def log_prob(self, x, base_dist=None):
        return flow_log_prob(self, x, base_dist, getattr(self, "_log_prob_fused", None))

    def sample(self, num_samples, base_dist=None, device=None):
        if device is None:
            device = next(self.parameters()).device
        if base_dist is None:
            base_dist = _default_base(self._data_dim(), device)
        z = base_dist.sample((num_samples,)).to(device)
        x, _ = self.forward(z)
        return x


class NormalizingFlowModel(_DensityAPI, nn.Module):
    """Chain of flow layers, optionally with an invertible BatchNorm affine between consecutive layers.

    The between-layer BatchNorm always transforms with its *running* statistics (so that forward, inverse and
    log-det share constants); train mode only moves the running statistics towards the batch statistics
    (momentum, biased variance) before using them (normalizing_flow_model.py:67-85).  Its log-det is the scalar
    sum(log|gamma| - 0.5*log(var+eps)), broadcast over rows (:87-108).
    """

    def __init__(self, flows, batch_norm_between_layers=False):
        super().__init__()
        self.batch_norm_between_layers = batch_norm_between_layers
        if self.batch_norm_between_layers:
            data_dim = getattr(flows[0], "data_dim", None)
            if data_dim is None:
                raise ValueError("Cannot use batch_norm_between_layers if flows do not have a 'data_dim' attribute.")
            self.batch_norms = nn.ModuleList([nn.BatchNorm1d(data_dim) for _ in range(len(flows))])
        self.flows = nn.ModuleList(flows)
        self._chain = ChainPlan()

    # -- between-layer BatchNorm pieces: [B,D] work in kernels, [D]-sized constants on the parameter vectors ----
    @staticmethod
    def _bn_sd(bn):
        return torch.sqrt(bn.running_var + bn.eps)

    def _apply_batch_norm(self, bn_layer, x):
        if self.training:
            mean, var = ops.col_stats(x)
            with torch.no_grad():
                momentum = bn_layer.momentum if bn_layer.momentum is not None else 0.1
                bn_layer.running_mean.mul_(1 - momentum).add_(momentum * mean.to(bn_layer.running_mean.dtype))
                bn_layer.running_var.mul_(1 - momentum).add_(momentum * var.to(bn_layer.running_var.dtype))
        return ops.feature_affine(x, bn_layer.running_mean, self._bn_sd(bn_layer), bn_layer.weight, bn_layer.bias)

    def _inverse_batch_norm(self, bn_layer, y):
        return ops.feature_affine(y, bn_layer.bias, bn_layer.weight, self._bn_sd(bn_layer), bn_layer.running_mean)

    def _batch_norm_log_det_jacobian(self, bn_layer, x):
        return (torch.log(torch.abs(bn_layer.weight)) - 0.5 * torch.log(bn_layer.running_var + bn_layer.eps)).sum()

    def _bns(self):
        return self.batch_norms if self.batch_norm_between_layers else None

    def _log_prob_fused(self, x):
        out = self._chain.run(self.flows, self._bns(), self.training, x, True, head=True)
        return None if out is None else out[1]

    @on_input_device
    def forward(self, z):
        """Sampling direction z -> x (:25-46)."""
        fused = self._chain.run(self.flows, self._bns(), self.training, z, False)
        if fused is not None:
            return fused
        log_det_sum = 0
        last = len(self.flows) - 1
        for i, flow in enumerate(self.flows):
            z, log_det = flow(z)
            log_det_sum = log_det_sum + log_det
            if self.batch_norm_between_layers and i < last:
                bn = self.batch_norms[i]
                z = self._apply_batch_norm(bn, z)
                log_det_sum = log_det_sum + self._batch_norm_log_det_jacobian(bn, z)
        return z, log_det_sum

    @on_input_device
    def inverse(self, x):
        """Density direction x -> z (:48-65)."""
        fused = self._chain.run(self.flows, self._bns(), self.training, x, True)
        if fused is not None:
            return fused
        log_det_sum = 0
        last = len(self.flows) - 1
        for i in range(last, -1, -1):
            if self.batch_norm_between_layers and i < last:
                bn = self.batch_norms[i]
                x = self._inverse_batch_norm(bn, x)
                log_det_sum = log_det_sum - self._batch_norm_log_det_jacobian(bn, x)
            x, log_det = self.flows[i].inverse(x)
            log_det_sum = log_det_sum + log_det
        return x, log_det_sum


def _half_masks(data_dim, n_layers):
    """Alternating half masks: even layers condition on the first half, odd layers on the second
    (real_nvp.py:24-33, real_nvp_spline.py:22-31)."""
    first = torch.zeros(data_dim)
    first[: data_dim // 2] = 1
    return [first.clone() if i % 2 == 0 else 1 - first for i in range(n_layers)]


class RealNVP(_DensityAPI, nn.Module):
    def __init__(self, data_dim, n_layers, hidden_dim, batch_norm_between_layers=False):
        super().__init__()
        assert n_layers % 2 == 0, "Number of layers must be even to ensure all dimensions are transformed."
        layers = [CouplingLayer(data_dim, hidden_dim, m) for m in _half_masks(data_dim, n_layers)]
        self.flow = NormalizingFlowModel(layers, batch_norm_between_layers)

    def _log_prob_fused(self, x):
        return self.flow._log_prob_fused(x)

    def forward(self, z):
        return self.flow.forward(z)

    def inverse(self, x):
        return self.flow.inverse(x)


class RealNVPSpline(_DensityAPI, nn.Module):
    def __init__(self, data_dim, n_layers, hidden_dim, batch_norm_between_layers=False):
        super().__init__()
        assert n_layers % 2 == 0, "Number of layers must be even to ensure all dimensions are transformed."
        layers = [SplineCouplingLayer(data_dim, hidden_dim, m) for m in _half_masks(data_dim, n_layers)]
        self.flow = NormalizingFlowModel(layers, batch_norm_between_layers)

    def _log_prob_fused(self, x):
        return self.flow._log_prob_fused(x)

    def forward(self, z):
        return self.flow.forward(z)

    def inverse(self, x):
        return self.flow.inverse(x)
