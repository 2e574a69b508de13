"""ncu target: bench.py's C3 model (MaskedAutoregressiveFlow(64, 512), 262144 rows), model.log_prob(x) three times in the
fp32-parity mode -- the chain of four tcgen05 GEMMs + affine_ar (+ fused N(0,I) head) that bench.py's also.c3 roofline block
describes.  usage: c3_logprob_target.py [--precision fp32|bf16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import nfb200 as N

prec = sys.argv[sys.argv.index("--precision") + 1] if "--precision" in sys.argv else "fp32"
N.set_gemm_precision(prec)
model = bench.build_model("c3", N).cuda().eval()
torch.manual_seed(3)
x = torch.randn(bench.WORKLOADS["c3"]["rows"], bench.WORKLOADS["c3"]["D"], device="cuda")
with torch.no_grad():
    for _ in range(3):
        before = N._lib.launch_count()
        lp = model.log_prob(x)
    torch.cuda.synchronize()
print("launches per pass", N._lib.launch_count() - before, float(lp.mean()))
