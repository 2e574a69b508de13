"""Pull products of the blocked sequential direction in isolation: Y[:, u0:u0+N] = X[:, :K] W[u0:u0+N, :K]^T on column
slices of [M, H] buffers (nf_linear_tc with row pitches H), per K.  usage: slice_gemm_bench.py [M] [H]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nfb200 as N
from nfb200 import _lib as L

M = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
H = int(sys.argv[2]) if len(sys.argv) > 2 else 536
dev = "cuda"
x = torch.randn(M, H, device=dev)
y = torch.zeros(M, H, device=dev)
w = torch.randn(H, H, device=dev) / H ** 0.5
hi, lo = N.ops.split_tf32(w)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
st = L.stream()
for (K, Nn) in ([tuple(int(v) for v in a.split("x")) for a in sys.argv[3:]] or [(68, 68), (136, 68), (204, 68), (272, 68), (340, 68), (408, 68), (476, 60), (476, 68), (448, 64), (512, 24), (408, 136)]):
    u0 = K if K + Nn <= H else H - Nn
    u0 -= u0 % 4
    def run():
        rc = L.lib().nf_linear_tc(L.ptr(x), hi.data_ptr() + 4 * u0 * H, lo.data_ptr() + 4 * u0 * H, None, y.data_ptr() + 4 * u0,
                                  M, Nn, K, H, H, H, 0, None, st)
        assert rc == 0, rc
    for _ in range(2):
        run()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    ref = (x[:4096, :K].double() @ w[u0:u0 + Nn, :K].double().T)
    err = float((y[:4096, u0:u0 + Nn].double() - ref).abs().max())
    print(json.dumps({"M": M, "K": K, "N": Nn, "u0": u0, "ms": round(ms, 4), "read_GBps": round(M * K * 4 / ms / 1e6, 1),
                      "tflops_fp32_equiv": round(2.0 * M * K * Nn / ms / 1e9, 1), "max_abs_err": err}))
