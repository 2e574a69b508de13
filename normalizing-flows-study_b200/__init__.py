"""B200-native (sm_100a) implementation of the normalizing-flow transform hot path of
itxtx/normalizing-flows-study, behind the reference's own nn.Module API.

The directory name is not a Python identifier; import it through the `nfb200` alias module at the repo root
(or the `src.flows` / `src.models` drop-in shims).
"""
from . import _lib, ops, packing, parallel, graphs  # noqa: F401
from .flows import (Flow, SequentialFlow, CouplingLayer, SplineCouplingLayer, rational_quadratic_spline,  # noqa: F401
                    MaskedLinear, MADE, MaskedAutoregressiveFlow, InverseAutoregressiveFlow, ARQS)
from .models import NormalizingFlowModel, RealNVP, RealNVPSpline  # noqa: F401

__all__ = ["Flow", "SequentialFlow", "CouplingLayer", "SplineCouplingLayer", "rational_quadratic_spline",
           "MaskedLinear", "MADE", "MaskedAutoregressiveFlow", "InverseAutoregressiveFlow", "ARQS",
           "NormalizingFlowModel", "RealNVP", "RealNVPSpline", "set_strict_fp32"]


def set_strict_fp32(flag: bool = True) -> None:
    """Route every dense layer through the FP32-pipe GEMM (round-to-nearest FFMA accumulation) instead of the tcgen05
    3xTF32 kernels.  The tensor core truncates its fp32 accumulator on every MMA, so long contractions (hidden_dim
    >= 1024) drift: one RealNVPSpline(784, ., 1024) layer shows a coherent log-det bias of ~4e-4 (DESIGN.md,
    "fp32 parity on the tensor cores").  Strict mode meets the reference's fp32 error everywhere at ~1/5 of the GEMM
    throughput; the fused D <= 8 stacks (K = 64 contractions) are unaffected and stay on the tensor cores."""
    ops.USE_TENSOR_CORE_GEMM = not flag
