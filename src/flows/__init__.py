import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
from nfb200 import (Flow, SequentialFlow, MaskedLinear, MADE, MaskedAutoregressiveFlow,  # noqa: E402,F401
                    InverseAutoregressiveFlow, CouplingLayer, SplineCouplingLayer, rational_quadratic_spline, ARQS)



class ContinuousFlow(Flow):
    """Placeholder so that `from src.flows import ContinuousFlow` (plots/_common.py:24) keeps importing: the CNF path
    (src/flows/continuous, torchdiffeq) is outside the hot path this package replaces."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("ContinuousFlow is out of scope of the B200 hot-path package; use the reference's")


__all__ = ["ContinuousFlow", "Flow", "SequentialFlow", "MaskedLinear", "MADE", "MaskedAutoregressiveFlow",
           "InverseAutoregressiveFlow", "CouplingLayer", "SplineCouplingLayer", "rational_quadratic_spline", "ARQS"]
