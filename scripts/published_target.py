#!/usr/bin/env python
"""The reference's four published configs (plots/_common.py:158-170) at n rows, both directions, once warm and --reps
times timed: the target of a launch-list capture (which kernels the sampling / density passes run, and for how long).

    python scripts/published_target.py [--n 1048576] [--reps 3] [--only maf]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nfb200 as N  # noqa: E402

BUILD = {
    "realnvp": lambda: N.RealNVP(2, 10, 128),
    "spline": lambda: N.RealNVPSpline(2, 8, 64),
    "maf": lambda: N.NormalizingFlowModel([N.MaskedAutoregressiveFlow(2, 64) for _ in range(6)]),
    "iaf": lambda: N.NormalizingFlowModel([N.InverseAutoregressiveFlow(2, 64) for _ in range(6)]),
    # the reference's notebooks (1_Basics_Coupling_Flow, 2_Autoregressive_Flows, 4_Neural_Spline_Flows; plots/fig_gif.py)
    "nb_realnvp256": lambda: N.RealNVP(2, 8, 256),
    "nb_maf128": lambda: N.NormalizingFlowModel([N.MaskedAutoregressiveFlow(2, 128) for _ in range(8)]),
    "nb_iaf128": lambda: N.NormalizingFlowModel([N.InverseAutoregressiveFlow(2, 128) for _ in range(8)]),
    "nb_spline128": lambda: N.RealNVPSpline(2, 8, 128),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    for name in [k for k in BUILD if (k in a.only.split(",") if a.only else not k.startswith("nb_"))]:
        torch.manual_seed(0)
        m = BUILD[name]()
        with torch.no_grad():
            for p in m.parameters():
                p.add_(torch.randn_like(p) * 0.05)
        m.to(dev).eval()
        z = torch.randn(a.n, 2, device=dev)
        rec = {"flow": name, "n": a.n}
        with torch.no_grad():
            for what, fn in (("forward", m.forward), ("inverse", m.inverse)):
                fn(z)
                torch.cuda.synchronize()
                l0 = N._lib.launch_count()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(a.reps):
                    fn(z)
                e.record()
                torch.cuda.synchronize()
                ms = s.elapsed_time(e) / a.reps
                rec[what] = {"ms": ms, "samples_per_s": a.n / ms * 1e3, "launches": (N._lib.launch_count() - l0) / a.reps}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
