"""One-pass dense layer: converters + TMEM A operand (default) against the SS form (nf_set_option(10, 1))."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nfb200 as N
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
N.set_gemm_precision("tf32")
for M, Nn, K in [(262144, 512, 512), (4096, 22736, 1024), (262144, 512, 64)]:
    x = torch.randn(M, K, device="cuda"); w = torch.randn(Nn, K, device="cuda") / K ** 0.5; b = torch.randn(Nn, device="cuda")
    hi, lo = N.ops.split_tf32(w)
    ref = torch.relu(x.double() @ w.double().T + b.double()) if M * Nn <= 1 << 27 else None
    for ss in (0, 1):
        N._lib.call("nf_set_option", 10, ss)
        ms = t(lambda: N.ops.linear_tc(x, hi, lo, b, relu=True))
        y = N.ops.linear_tc(x, hi, lo, b, relu=True)
        err = "" if ref is None else f"  rms err / rms y = {((y.double() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item():.3e}  mean err / rms y = {((y.double() - ref).mean() / ref.pow(2).mean().sqrt()).item():+.3e}"
        print(f"[{M}x{Nn}x{K}] ss={ss}: {ms:.3f} ms  {2.0 * M * Nn * K / ms / 1e9:.1f} TFLOP/s{err}", flush=True)
N._lib.call("nf_set_option", 10, 0)
N.set_gemm_precision("fp32")
