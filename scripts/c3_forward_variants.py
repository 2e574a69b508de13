"""C3 sequential direction (MaskedAutoregressiveFlow(64, 512).forward, 262144 rows): time against the dense-layer
kernel selection for the K <= 128 slice GEMMs (nf_set_option(6, v))."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nfb200 as N
torch.manual_seed(0)
m = N.MaskedAutoregressiveFlow(64, 512).cuda().eval()
with torch.no_grad():
    for p in m.parameters():
        p.add_(0.02 * torch.randn_like(p))
z = torch.randn(262144, 64, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
with torch.no_grad():
    for small_k in (1, 0, 1, 0):
        N._lib.call("nf_set_option", 6, small_k)
        print(f"small_k={small_k}: forward {t(lambda: m.forward(z)):.3f} ms   inverse {t(lambda: m.inverse(z)):.3f} ms", flush=True)
N._lib.call("nf_set_option", 6, 1)
