"""Debug: issue / completion cycles of tcgen05.mma chains (TS mode, kind::tf32, M=128) vs chain length, N and
number of independent accumulators."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nfb200 as N
a = torch.randn(128, 64, device="cuda")
for n_out in (32, 64, 128):
    img = torch.from_numpy(N.packing.umma_sw128_images(np.random.randn(n_out, 64).astype(np.float32))).cuda()
    d = torch.empty(128, n_out, device="cuda")
    t = torch.zeros(2, dtype=torch.int64, device="cuda")
    for nacc in (1, 2, 3):
        for passes in (1, 3, 6, 12):
            for _ in range(3):
                N._lib.call("nf_debug_tc_gemm128", a.data_ptr(), img.data_ptr(), d.data_ptr(), n_out, passes, t.data_ptr(), nacc, N._lib.stream())
            torch.cuda.synchronize()
            i, c = t.tolist()
            print(f"N={n_out:3d} accumulators={nacc} mmas={passes*8:3d}: issue {i:6d} cyc, issue->complete {c:6d} cyc ({c/(passes*8):6.1f} per MMA)")
