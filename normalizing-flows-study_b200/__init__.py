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
           "NormalizingFlowModel", "RealNVP", "RealNVPSpline"]
