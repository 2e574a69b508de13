"""Error of the tcgen05 3xTF32 forward GEMM vs float64, next to cuBLAS fp32 and the FP32-pipe GEMM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nfb200 as N
torch.backends.cuda.matmul.allow_tf32 = False
for M, Nn, K in [(4096, 512, 512), (4096, 1024, 1024), (4096, 512, 64), (2048, 2048, 4096)]:
    for kind in ("randn", "relu-act", "positive"):
        gen = torch.Generator(device="cuda").manual_seed(1)
        x = torch.randn(M, K, device="cuda", generator=gen); w = torch.randn(Nn, K, device="cuda", generator=gen) / K ** 0.5
        if kind == "relu-act": x = x.clamp_min(0)
        if kind == "positive": x, w = x.abs(), w.abs()
        ref = x.double() @ w.double().T
        rms = ref.pow(2).mean().sqrt().item()
        hi, lo = N.ops.split_tf32(w)
        def rep(name, y):
            e = y.double() - ref
            print(f"  {name:22s} max|err|/rms {e.abs().max().item()/rms:.2e}  mean err/rms {e.mean().item()/rms:+.2e}  rms err/rms {e.pow(2).mean().sqrt().item()/rms:.2e}", flush=True)
        print(f"M={M} N={Nn} K={K} {kind}: rms(y)={rms:.3g}")
        rep("torch fp32 (cuBLAS)", x @ w.T)
        rep("FP32-pipe nf_gemm", N.ops.linear_raw(x, w))
        rep("tcgen05 3xTF32", N.ops.linear_tc(x, hi, lo))
