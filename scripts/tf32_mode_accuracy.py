"""Measured effect of the reduced-precision dense-layer mode (set_gemm_precision("tf32"), one TF32 pass) against the
default 3xTF32 mode on the same weights and inputs: whole-model density passes and gradients.
    python scripts/tf32_mode_accuracy.py > gpurun_out/tf32_mode_accuracy.log"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import nfb200 as N  # noqa: E402

torch.manual_seed(0)
CASES = [("MaskedAutoregressiveFlow(64, 512)", lambda: N.MaskedAutoregressiveFlow(64, 512), 64, 8192, 0.02),
         ("4 x MaskedAutoregressiveFlow(256, 1024)", lambda: N.NormalizingFlowModel([N.MaskedAutoregressiveFlow(256, 1024) for _ in range(4)]), 256, 8192, 0.02),
         ("RealNVP(256, 8, 512)", lambda: N.RealNVP(256, 8, 512), 256, 8192, 0.02),
         ("RealNVPSpline(784, 2, 1024)", lambda: N.RealNVPSpline(784, 2, 1024), 784, 2048, 0.01),
         ("RealNVPSpline(784, 16, 1024)", lambda: N.RealNVPSpline(784, 16, 1024), 784, 2048, 0.01)]
for name, make, D, B, sigma in CASES:
    m = make().cuda().eval()
    with torch.no_grad():
        for p in m.parameters():
            p.add_(sigma * torch.randn_like(p))
    x = torch.randn(B, D, device="cuda")
    out = {}
    for mode in ("fp32", "tf32"):
        N.set_gemm_precision(mode)
        for p in m.parameters():
            p.grad = None
        z, ld = m.inverse(x)
        loss = -N.ops.std_normal_log_prob(z, ld).mean()
        loss.backward()
        g = torch.cat([p.grad.flatten() for p in m.parameters() if p.grad is not None])
        out[mode] = (z.detach(), ld.detach(), loss.item(), g)
    N.set_gemm_precision("fp32")
    (z0, l0, n0, g0), (z1, l1, n1, g1) = out["fp32"], out["tf32"]
    dz = ((z1 - z0).abs() / z0.abs().clamp_min(1)).max().item()
    print(f"{name}: B={B}  max|dz|/max(1,|z|) = {dz:.3e}   max|d log_det| = {(l1 - l0).abs().max().item():.3e}   "
          f"mean|d log_det| = {(l1 - l0).abs().mean().item():.3e}   NLL {n0:.6f} -> {n1:.6f}   "
          f"|dgrad| / |grad| = {((g1 - g0).norm() / g0.norm()).item():.3e}", flush=True)
    del m, out
    torch.cuda.empty_cache()
