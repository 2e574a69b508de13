"""Small driver for compute-sanitizer runs (memcheck / racecheck / synccheck): every warp-specialised mbarrier / TMEM
kernel of the library on small shapes -- smoke() (fused tcgen05 spline stack incl. the log-prob head, layered training step
with the tensor-core forward / dgrad / wgrad GEMMs, MAF chain + blocked sequential direction), the affine coupling stack on
tcgen05, the K > 128 short-chain GEMM, and the bf16 fused MADE chain.  Exit code 0 = all finished and matched.
    compute-sanitizer --tool memcheck python scripts/sanitizer_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as G  # noqa: E402
import nfb200 as N  # noqa: E402

dev = "cuda:0"
G.smoke()
torch.manual_seed(0)
# affine coupling stack on tcgen05 (coupling_stack_tc_kernel)
m = N.RealNVP(2, 4, 64).to(dev).eval()
with torch.no_grad():
    for p in m.parameters():
        p.add_(0.05 * torch.randn_like(p))
    x = torch.randn(1500, 2, device=dev)
    z, ld = m.inverse(x)
    xr, ld2 = m.forward(z)
    assert torch.allclose(xr, x, atol=1e-4) and torch.allclose(ld + ld2, torch.zeros_like(ld), atol=1e-3)
# K > 128 dense layers (gemm_tc2_kernel<false>, wgrad) through a wide MAF training step
maf = N.MaskedAutoregressiveFlow(32, 256).to(dev).train()
xm = torch.randn(1024, 32, device=dev)
zz, l = maf.inverse(xm)
(-(N.ops.std_normal_log_prob(zz, l)).mean()).backward()
assert all(torch.isfinite(p.grad).all() for p in maf.parameters())
# bf16 fused chain (made_chain_bf16_kernel): several tiles per CTA would need > 148 * 128 rows; 3 tiles here
maf.eval()
N.set_gemm_precision("bf16")
with torch.no_grad():
    xb = torch.randn(300, 32, device=dev)
    zb, lb = maf.inverse(xb)
    lp = maf.log_prob(xb)
N.set_gemm_precision("fp32")
with torch.no_grad():
    z32, l32 = maf.inverse(xb)
assert torch.allclose(zb, z32, atol=2e-2, rtol=2e-2) and torch.allclose(lb, l32, atol=1e-1)
torch.cuda.synchronize()
print("sanitizer target ok")
