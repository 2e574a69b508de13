import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
from nfb200 import NormalizingFlowModel, RealNVP, RealNVPSpline  # noqa: E402,F401

__all__ = ["NormalizingFlowModel", "RealNVPSpline", "RealNVP"]
