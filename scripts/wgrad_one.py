"""Run the tcgen05 weight-gradient kernel on one shape a few times (ncu target): python scripts/wgrad_one.py B N K"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nfb200 as N
B, Nn, K = (int(a) for a in sys.argv[1:4])
g = torch.randn(B, Nn, device="cuda"); x = torch.randn(B, K, device="cuda")
for _ in range(6):
    dw = N.ops.linear_wgrad_tc(g, x)
torch.cuda.synchronize()
print("ok", float(dw[0, 0]))
