"""Golden vectors for MADE / MAF / IAF built with use_batch_norm=True (made.py:93-108), from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_bn.py

Eval mode (running statistics, both directions) and one train-mode pass of the parallel direction (batch statistics +
running-stat side effects).  Same conventions as make_golden.py, which this script imports its helpers from.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def main():
    RF, _ = MG._import_reference()
    g = torch.Generator().manual_seed(77)
    cases = {}
    for D, H in ((5, 32), (16, 64)):
        for name, cls in (("maf", RF.MaskedAutoregressiveFlow), ("iaf", RF.InverseAutoregressiveFlow)):
            torch.manual_seed(9)
            layer = cls(D, H, use_batch_norm=True)
            MG.perturb(layer, g, 0.25)
            layer.eval()
            x = torch.randn(40, D, generator=g)
            sd = MG.clone_sd(layer)
            cases[f"{name}bn_D{D}_H{H}"] = dict(kind=name, D=D, H=H, use_batch_norm=True, sd=sd, x=x,
                                                **MG.run_both(layer, x),
                                                degrees=torch.as_tensor(layer.conditioner.m[0]))
            # train mode, parallel direction only (the sequential one would update the statistics D times)
            layer.train()
            sd0 = MG.clone_sd(layer)
            with torch.no_grad():
                y, ld = layer.inverse(x) if name == "maf" else layer.forward(x)
            cases[f"{name}bn_train_D{D}_H{H}"] = dict(kind=name + "_train", D=D, H=H, use_batch_norm=True, sd=sd0,
                                                      sd_after=MG.clone_sd(layer), x=x, out=y.clone(), out_ld=ld.clone())
    for k, v in cases.items():
        torch.save(v, os.path.join(HERE, k + ".pt"))
        print("wrote", k)


if __name__ == "__main__":
    main()
