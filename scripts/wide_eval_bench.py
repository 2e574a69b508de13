"""Eval-mode throughput of the wide (large data_dim) models of BASELINE configs 4/5: log_prob + sample passes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nfb200 as N
dev = torch.device("cuda:0")
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
MODELS = {"realnvp256": (lambda: N.RealNVP(256, 8, 512), 256, 262144),
          "maf256x4": (lambda: N.NormalizingFlowModel([N.MaskedAutoregressiveFlow(256, 1024) for _ in range(4)]), 256, 262144),
          "spline784": (lambda: N.RealNVPSpline(784, 16, 1024), 784, 16384)}
for name, (mk, D, B) in MODELS.items():
    torch.manual_seed(0)
    m = mk().to(dev).eval()
    with torch.no_grad():
        for p in m.parameters(): p.add_(0.01 * torch.randn_like(p))
        x = torch.randn(B, D, device=dev)
        l0 = N._lib.launch_count()
        ms_inv = timeit(lambda: m.inverse(x))
        res = {"model": name, "rows": B, "inverse_ms": ms_inv, "inverse_rows_per_s": B / ms_inv * 1e3,
               "launches_per_inverse": (N._lib.launch_count() - l0) / 7}
        if name != "maf256x4":
            ms_fwd = timeit(lambda: m.forward(x))
            res.update(forward_ms=ms_fwd, forward_rows_per_s=B / ms_fwd * 1e3)
    print(json.dumps(res), flush=True)
