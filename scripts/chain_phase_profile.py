"""Debug: where CTA 0 of made_chain_bf16_kernel waits, per warp role (needs lib/libnfb200_prof.so, scripts/build_profile_lib.py).
    NFB200_LIB=normalizing-flows-study_b200/lib/libnfb200_prof.so python scripts/chain_phase_profile.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NFB200_LIB", os.path.join(ROOT, "normalizing-flows-study_b200", "lib", "libnfb200_prof.so"))
import torch  # noqa: E402

import nfb200 as N  # noqa: E402
import bench  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
m = bench.build_model("c3", N).cuda().eval()
x = bench.make_inputs("c3", rows, 0)[0].cuda()
N.set_gemm_precision("bf16")
with torch.no_grad():
    for _ in range(3):
        m.inverse(x)
    torch.cuda.synchronize()
out = (ctypes.c_longlong * 16)()
raw = ctypes.CDLL(N._lib.LIB_PATH)
assert raw.nf_debug_mc_profile(out) == 0
names = {0: "producer: wait free weight stage", 1: "producer total",
         2: "MMA warp: wait x", 3: "MMA warp: wait accumulator empty", 4: "MMA warp: wait input block ready",
         5: "MMA warp: wait weight stage full", 6: "MMA warp total",
         8: "epilogue warp 2: wait hidden-layer accumulator full", 9: "epilogue warp 2: wait last-layer accumulator full",
         10: "epilogue warp 2 total"}
tiles = -(-rows // 128)
mine = len(range(0, tiles, 148))
for i, n in names.items():
    tot = out[6] if 2 <= i <= 6 else (out[1] if i < 2 else out[10])
    print(f"{n:52s} {out[i]:12d} cycles  {100.0 * out[i] / max(tot, 1):5.1f}%   ({out[i] / mine:9.0f} per tile)")
