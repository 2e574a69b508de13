"""MAF(64, 512).inverse at 262144 rows in the three GEMM precision modes (C3 log_prob direction): ms per pass, dense
TFLOP/s (SURVEY 8d accounting: 1 245 184 FLOP per row) and the fraction of the measured bf16 peak.
    python scripts/chain_bench.py [rows] [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import nfb200 as N  # noqa: E402
from bench import build_model, make_inputs  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
peak = 1620.0
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
except Exception:
    pass
model = build_model("c3", N).to(dev).eval()
sets = [make_inputs("c3", rows, s)[0].to(dev) for s in range(4)]          # 4 x 67 MB > L2
flops = 2 * (64 * 512 + 2 * 512 * 512 + 512 * 128) * rows
for mode in ("fp32", "tf32", "bf16"):
    N.set_gemm_precision(mode)
    for what, fn in (("inverse", lambda x: model.inverse(x)), ("log_prob", lambda x: model.log_prob(x))):
        with torch.no_grad():
            for i in range(3):
                fn(sets[i % 4])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = N._lib.launch_count()
            e0.record()
            for i in range(reps):
                fn(sets[i % 4])
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"mode": mode, "pass": what, "rows": rows, "ms": ms, "launches_per_pass": (N._lib.launch_count() - l0) / reps,
                          "dense_tflops": flops / ms / 1e9, "frac_of_bf16_peak": flops / ms / 1e9 / peak,
                          "rows_per_s": rows / ms * 1e3}), flush=True)
N.set_gemm_precision("fp32")
