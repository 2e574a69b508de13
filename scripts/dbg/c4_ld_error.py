import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import nfb200 as N
from oracle import flows_oracle as O
D, H, B = 784, 1024, 300
mask = torch.tensor([1.0 if i % 2 == 0 else 0.0 for i in range(D)])
m = N.SplineCouplingLayer(D, H, mask, num_bins=10)
g = torch.Generator().manual_seed(D + H)
with torch.no_grad():
    for p in m.parameters(): p.add_(0.3 / H ** 0.5 * torch.randn(p.shape, generator=g))
m.eval()
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
x = torch.randn(B, D, generator=torch.Generator().manual_seed(B)) * 1.5
ry, rld = O.spline_coupling(sd, "", x, False, num_bins=10)
y64, ld64 = O.spline_coupling(sd64, "", x.double(), False, num_bins=10)
print("reference fp32 ld err: mean %+.2e max %.2e" % ((rld.double() - ld64).mean().item(), (rld.double() - ld64).abs().max().item()))
m = m.cuda()
def rep(name):
    with torch.no_grad():
        y, ld = m.forward(x.cuda())
    e = ld.cpu().double() - ld64
    ez = (y.cpu().double() - y64).abs() / (1 + y64.abs())
    print("%-34s ld err mean %+.2e max %.2e | z err max %.2e frac>1e-5 %.4f" % (name, e.mean().item(), e.abs().max().item(), ez.max().item(), (ez > 1e-5).double().mean().item()))
rep("tcgen05 wide route")
N.ops.USE_TENSOR_CORE_GEMM = False; N.flows.USE_TENSOR_CORES = False
rep("FP32-pipe GEMMs (layered)")
# params from the fp64 oracle conditioner through our transform kernel only
