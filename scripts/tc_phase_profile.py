"""Debug: per-phase cycle breakdown of the tcgen05 stack kernel (needs a library built with -DNF_TC_PROFILE,
NFB200_LIB pointing at it)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nfb200 as N
import bench
m = bench.build_model("c2", N).cuda().eval()
x = torch.randn(1 << 20, 2, device="cuda")
with torch.no_grad():
    for _ in range(3):
        m.inverse(x)
    torch.cuda.synchronize()
    out = (ctypes.c_longlong * 8)()
    raw = ctypes.CDLL(N._lib.LIB_PATH)
    raw.nf_debug_tc_profile(out)
    names = ["layer1+st+sync", "mma2 wait", "hidden2+st+sync", "head mma wait", "spline", "other(load x, bn, rescale)", "total"]
    tot = out[6]
    for n, v in zip(names, out):
        print(f"{n:28s} {v:12d} cycles  {100.0 * v / tot:5.1f}%")
