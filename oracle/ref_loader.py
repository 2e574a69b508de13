"""TEST INFRASTRUCTURE -- import the staged, unmodified reference (oracle/_ref, see make_ref.py).

Only bench.py's CPU / eager legs and tests/ may use this.  The reference's package is called `src`, like the repo's
own drop-in shim; `reference_modules()` therefore imports it with oracle/_ref first on sys.path and then REMOVES the
`src*` entries from sys.modules again (returning the module objects), so that a later `import src.flows` in the same
process still resolves to whatever the caller wants (the shim, in the harness tests).
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
STUBS = os.path.join(REF, "_stubs")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "src", "flows", "__init__.py"))


def _have(mod):
    try:
        importlib.import_module(mod)
        return True
    except Exception:
        return False


def stub_missing():
    """Put the torchdiffeq / matplotlib stubs on sys.path only when the real packages are missing."""
    for name in ("torchdiffeq", "matplotlib"):
        if name not in sys.modules and not _have(name):
            if STUBS not in sys.path:
                sys.path.append(STUBS)
            importlib.import_module(name)


def reference_modules():
    """(src.flows, src.models) of the unmodified reference, imported in isolation from the repo's `src` shim."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged (run `python oracle/make_ref.py` in the build container)")
    stub_missing()
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        flows = importlib.import_module("src.flows")
        models = importlib.import_module("src.models")
    finally:
        sys.path.remove(REF)
        mine = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
        for k in mine:
            del sys.modules[k]
        sys.modules.update(saved)
    ns = types.SimpleNamespace(flows=flows, models=models, modules=mine)
    for name in ("Flow", "SequentialFlow", "CouplingLayer", "SplineCouplingLayer", "MaskedLinear", "MADE",
                 "MaskedAutoregressiveFlow", "InverseAutoregressiveFlow", "rational_quadratic_spline"):
        setattr(ns, name, getattr(flows, name))
    for name in ("NormalizingFlowModel", "RealNVP", "RealNVPSpline"):
        setattr(ns, name, getattr(models, name))
    return ns
