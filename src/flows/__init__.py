import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
from nfb200 import (Flow, SequentialFlow, MaskedLinear, MADE, MaskedAutoregressiveFlow,  # noqa: E402,F401
                    InverseAutoregressiveFlow, CouplingLayer, SplineCouplingLayer, rational_quadratic_spline, ARQS)

__all__ = ["Flow", "SequentialFlow", "MaskedLinear", "MADE", "MaskedAutoregressiveFlow",
           "InverseAutoregressiveFlow", "CouplingLayer", "SplineCouplingLayer", "rational_quadratic_spline", "ARQS"]
