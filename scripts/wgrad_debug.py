import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nfb200 as N
torch.set_printoptions(linewidth=250, sci_mode=False, precision=1)
B, Nn, K = 32, 128, 128
def run(g, x):
    out = torch.full((Nn, K), float("nan"), device="cuda")
    return N.ops.linear_wgrad_tc(g.cuda().contiguous(), x.cuda().contiguous(), out=out).cpu()
bb = torch.arange(B, dtype=torch.float32)[:, None]
kk = torch.arange(K, dtype=torch.float32)[None, :]
nn_ = torch.arange(Nn, dtype=torch.float32)[None, :]
for variant in (0, 1, 2, 3, 4):
    os.environ["NF_WGRAD_VARIANT"] = str(variant)
    print("=== variant", variant)
    x = torch.zeros(B, K); x[0] = kk[0] + 1
    d = run(torch.ones(B, Nn), x); print("exp1 expect k+1:", d[0, :40]); print(d[5, :8], d[127, :8], "nonzero:", int((d != 0).sum()), "nan:", int(d.isnan().sum()))
    for b0 in (0, 1, 9):
        g = torch.zeros(B, Nn); g[b0] = 1
        x = bb * 1000 + kk + 1
        d = run(g, x)
        print(f"exp2 b0={b0}: expect {b0*1000}+k+1:", d[0, :10], d[0, 30:36], d[0, 64:68], "rows equal:", bool((d == d[0]).all()))
    for b0 in (0, 9):
        x = torch.zeros(B, K); x[b0] = 1
        g = bb * 1000 + nn_ + 1
        d = run(g, x)
        print(f"exp4 b0={b0}: expect {b0*1000}+n+1:", d[:10, 0], d[30:36, 0], d[64:68, 0], "cols equal:", bool((d == d[:, :1]).all()))
