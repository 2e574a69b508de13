"""Golden vectors for ARQS (src/flows/spline/arqs.py) from the UNMODIFIED reference; same conventions as
make_golden.py (run in the build container only: python tests/golden/make_golden_arqs.py)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _import_reference, perturb, clone_sd, run_both  # noqa: E402


def main():
    RF, _ = _import_reference()
    g = torch.Generator().manual_seed(4321)
    n = 0
    for D, H, K, rescale in ((1, 8, 4, False), (2, 16, 8, False), (3, 16, 8, False), (5, 32, 10, False),
                             (8, 32, 8, True), (4, 24, 6, False)):
        torch.manual_seed(11)
        kw = dict(data_min=-2.0, data_max=3.0) if rescale else {}
        layer = RF.ARQS(D, hidden_dim=H, num_bins=K, **kw)
        perturb(layer, g, 0.3)
        layer.eval()
        x = torch.rand(40, D, generator=g)                       # the public spline lives on [0,1]
        x[0] = 0.0
        x[1] = 1.0
        x[2] = 0.5
        x[3] = -0.25                                             # outside: clamped theta (no tails in this spline)
        x[4] = 1.75
        if rescale:
            x = x * 5.0 - 2.0
        blob = dict(kind="arqs", D=D, H=H, K=K, extra=kw, sd=clone_sd(layer), x=x, **run_both(layer, x))
        name = f"arqs_D{D}_H{H}_K{K}" + ("_rescale" if rescale else "")
        torch.save(blob, os.path.join(HERE, name + ".pt"))
        n += 1
    print(f"wrote {n} ARQS golden files")


if __name__ == "__main__":
    main()
