"""A/B of the two tcgen05 dense-layer kernels (nf_set_option(5, v)): time and error vs float64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nfb200 as N
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for M, Nn, K in [(262144, 512, 512), (262144, 512, 64), (262144, 128, 512), (65536, 1024, 1024), (65536, 1024, 256), (4096, 1024, 1024),
                 (4096, 11368, 1024), (1048576, 64, 64), (5000, 64, 64)]:
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(M, K, device="cuda", generator=gen).clamp_min(0); w = torch.randn(Nn, K, device="cuda", generator=gen) / K ** 0.5
    b = torch.randn(Nn, device="cuda", generator=gen)
    hi, lo = N.ops.split_tf32(w)
    sub = slice(0, min(M, 8192))
    ref = (x[sub].double() @ w.double().T + b.double()).clamp_min(0)
    rms = ref.pow(2).mean().sqrt().item()
    line = f"M={M} N={Nn} K={K}:"
    for v in (0, 1):
        N._lib.call("nf_set_option", 5, v)
        ms = t(lambda: N.ops.linear_tc(x, hi, lo, b, relu=True))
        y = N.ops.linear_tc(x, hi, lo, b, relu=True)
        e = y[sub].double() - ref
        line += f"  v{v}: {ms:.3f} ms {2.0*M*Nn*K/ms/1e9:6.1f} TF/s  rms {e.pow(2).mean().sqrt().item()/rms:.1e} max {e.abs().max().item()/rms:.1e} mean {e.mean().item()/rms:+.1e} |"
    N._lib.call("nf_set_option", 5, 1)
    print(line, flush=True)
