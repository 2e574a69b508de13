"""TEST INFRASTRUCTURE -- stages the UNMODIFIED reference for the GPU box.

/root/reference exists only in the build container.  This recipe copies the reference's own Python package (`src/`,
the hot path and everything it imports) and its benchmark harness (`plots/_common.py`) byte for byte into
`oracle/_ref/` -- git-ignored, so no reference source enters the history, but NOT gpurun-ignored, so the files travel to
the GPU box with the snapshot, exactly like the built `.so`.  Two import stubs are generated beside them
(`oracle/_ref/_stubs/`): `torchdiffeq` (src/flows/__init__.py:9 eagerly imports the out-of-scope CNF path) and
`matplotlib` (plots/_common.py:15-18 and src/utils.py:5 import it at module scope); neither is on the measured path.

Users of oracle/_ref (never the product):
  * bench.py --impl reference  -> the reference's own modules timed on the host cores (`kind: "reference"`);
  * bench.py's `eager_cuda` leg  -> the same modules on CUDA tensors = what users of the reference get on a B200 today;
  * tests/test_reference_harness.py -> plots/_common.samples_per_sec / load_cache and FlowProfiler.profile_flow run
    unmodified against the `src.*` drop-in shim.

    python oracle/make_ref.py            # no-op (exit 0) when /root/reference is absent
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("NF_REFERENCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")

_TORCHDIFFEQ = '''"""stub: the CNF path (out of scope) is imported eagerly by src/flows/__init__.py:9"""


def odeint(*a, **k):
    raise RuntimeError("torchdiffeq stub: continuous flows are out of scope")


odeint_adjoint = odeint
'''

_MPL_INIT = '''"""stub: plots/_common.py and src/utils.py import matplotlib at module scope; nothing on the measured path draws"""


def use(*a, **k):
    return None


class _Anything:
    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


rcParams = {}
'''

_MPL_PYPLOT = '''from . import _Anything, rcParams  # noqa: F401


def __getattr__(name):
    return _Anything()
'''

_MPL_COLORS = '''from . import _Anything


class LinearSegmentedColormap:
    @staticmethod
    def from_list(*a, **k):
        return _Anything()


def __getattr__(name):
    return _Anything()
'''


def stage(verbose=True):
    if not os.path.isdir(os.path.join(REF_SRC, "src")):
        if verbose:
            print(f"  oracle/_ref: {REF_SRC} not present, nothing staged (prebuilt copy is used if it exists)")
        return False
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc")
    shutil.copytree(os.path.join(REF_SRC, "src"), os.path.join(DEST, "src"), ignore=ignore)
    os.makedirs(os.path.join(DEST, "plots"))
    shutil.copy2(os.path.join(REF_SRC, "plots", "_common.py"), os.path.join(DEST, "plots", "_common.py"))
    stubs = os.path.join(DEST, "_stubs")
    os.makedirs(os.path.join(stubs, "torchdiffeq"))
    os.makedirs(os.path.join(stubs, "matplotlib"))
    with open(os.path.join(stubs, "torchdiffeq", "__init__.py"), "w") as f:
        f.write(_TORCHDIFFEQ)
    with open(os.path.join(stubs, "matplotlib", "__init__.py"), "w") as f:
        f.write(_MPL_INIT)
    with open(os.path.join(stubs, "matplotlib", "pyplot.py"), "w") as f:
        f.write(_MPL_PYPLOT)
    with open(os.path.join(stubs, "matplotlib", "colors.py"), "w") as f:
        f.write(_MPL_COLORS)
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DEST))
        print(f"  oracle/_ref: staged {n} files from {REF_SRC}")
    return True


if __name__ == "__main__":
    stage()
    sys.exit(0)
