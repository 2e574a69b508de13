"""Run the tcgen05 linear kernel on one shape a few times (ncu target): python scripts/gemm_one.py M N K [fp32|tf32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nfb200 as N
M, Nn, K = (int(a) for a in sys.argv[1:4])
N.set_gemm_precision(sys.argv[4] if len(sys.argv) > 4 else "fp32")
x = torch.randn(M, K, device="cuda"); w = torch.randn(Nn, K, device="cuda") / K ** 0.5; b = torch.randn(Nn, device="cuda")
hi, lo = N.ops.split_tf32(w)
for _ in range(6):
    y = N.ops.linear_tc(x, hi, lo, b, relu=True)
torch.cuda.synchronize()
print("ok", float(y[0, 0]))
