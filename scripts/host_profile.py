"""cProfile of the eager training step at the C1 size (host overhead per launch)."""
import cProfile, pstats, os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nfb200 as N
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = N.RealNVP(2, 8, 64).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
x = torch.randn(5000, 2, device=dev)
def step():
    opt.zero_grad(set_to_none=True)
    z, ld = model.inverse(x)
    loss = -N.ops.std_normal_log_prob(z, ld).mean()
    loss.backward()
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
import time
t = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t) / 20 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
