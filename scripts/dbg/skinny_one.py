import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import nfb200 as N
B, H, P = 1 << 20, 64, 23
g = torch.randn(B, P, device="cuda"); w = torch.randn(P, H, device="cuda"); x = torch.randn(B, H, device="cuda")
for _ in range(3):
    dx = N.ops.gemm(g, w, B, H, P, P, 1, H, 1)          # dX = dY W  (K = 23)
    dw = N.ops.gemm(g, x, P, H, B, 1, P, H, 1)          # dW = dY^T X (M = 23)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); dx = N.ops.gemm(g, w, B, H, P, P, 1, H, 1); e1.record(); torch.cuda.synchronize(); print("dX ms", e0.elapsed_time(e1))
e0.record(); dw = N.ops.gemm(g, x, P, H, B, 1, P, H, 1); e1.record(); torch.cuda.synchronize(); print("dW ms", e0.elapsed_time(e1))
