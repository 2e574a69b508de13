"""Time the tcgen05 weight-gradient kernel against the FP32-pipe GEMM on the training shapes (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nfb200 as N

def t(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

SHAPES = [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]]
for B, Nn, K in SHAPES or [(262144, 512, 512), (262144, 512, 64), (262144, 128, 512), (65536, 512, 512), (65536, 1024, 1024),
                 (65536, 512, 256), (4096, 1024, 1024), (4096, 11368, 1024), (4096, 1024, 784)]:
    g = torch.randn(B, Nn, device="cuda"); x = torch.randn(B, K, device="cuda")
    a = t(lambda: N.ops.linear_wgrad_tc(g, x))
    b = t(lambda: N.ops.gemm(g, x, Nn, K, B, 1, Nn, K, 1), n=2)
    fl = 2.0 * B * Nn * K
    print(f"wgrad B={B} N={Nn} K={K}: tc {a:.3f} ms ({fl / a / 1e9:.1f} TFLOP/s)   fp32 pipe {b:.3f} ms ({fl / b / 1e9:.1f} TFLOP/s)", flush=True)
