"""Initial-weight parity fixture: constructs reference modules under a fixed seed and stores their state_dicts
(tests/golden/init_parity.pt).  Build container only (needs /root/reference); see make_golden.py."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _import_reference  # noqa: E402


def main():
    RF, RM = _import_reference()
    seed = 2024
    ctors = {
        "coupling": lambda: RF.CouplingLayer(4, 16, torch.tensor([1., 0., 1., 0.])),
        "spline": lambda: RF.SplineCouplingLayer(4, 16, torch.tensor([1., 1., 0., 0.]), num_bins=6),
        "maf": lambda: RF.MaskedAutoregressiveFlow(5, 32),
        "iaf": lambda: RF.InverseAutoregressiveFlow(5, 32),
        "made_bn": lambda: RF.MADE(3, 8, 2, use_batch_norm=True),
        "realnvp_bn": lambda: RM.RealNVP(4, 4, 16, batch_norm_between_layers=True),
        "realnvpspline": lambda: RM.RealNVPSpline(6, 2, 32),
    }
    blob = {"kind": "init", "seed": seed}
    for k, c in ctors.items():
        torch.manual_seed(seed)
        blob[k] = {n: v.detach().clone() for n, v in c().state_dict().items()}
    torch.save(blob, os.path.join(HERE, "init_parity.pt"))
    print("wrote init_parity.pt", os.path.getsize(os.path.join(HERE, "init_parity.pt")))


if __name__ == "__main__":
    main()
