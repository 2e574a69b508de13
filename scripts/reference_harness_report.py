"""The reference's published benchmark (plots/fig_benchmark.py via plots/_common.samples_per_sec: sampling direction,
n=4000, dim=2, eval, 1 warm-up + 3 reps, time.time()) for its four published configs (plots/_common.py:158-170), run
UNMODIFIED from oracle/_ref/plots/_common.py against (a) the src.* shim on CUDA and (b) the unmodified reference
modules on the host cores and on CUDA.  The harness does not synchronise the device, so a synchronised figure at the
same n and at n = 2^20 is printed beside it.  JSON lines on stdout.

    python scripts/reference_harness_report.py > gpurun_out/r02_reference_harness.jsonl
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import ref_loader  # noqa: E402
from tests.test_reference_harness import load_reference_plots_common  # noqa: E402

PUBLISHED_CPU = {"realnvp": 186e3, "spline": 334e3, "maf": 602e3, "iaf": 1121e3}


def synced_sps(model, n, dev, reps=5):
    z = torch.randn(n, 2, device=dev)
    with torch.no_grad():
        model.forward(z)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            model.forward(z)
        torch.cuda.synchronize()
    return n / ((time.perf_counter() - t) / reps)


def main():
    dev = "cuda:0"
    C = load_reference_plots_common()
    R = ref_loader.reference_modules()
    ref_build = {"realnvp": lambda: R.RealNVP(2, 10, 128), "spline": lambda: R.RealNVPSpline(2, 8, 64),
                 "maf": lambda: R.NormalizingFlowModel([R.MaskedAutoregressiveFlow(2, 64) for _ in range(6)]),
                 "iaf": lambda: R.NormalizingFlowModel([R.InverseAutoregressiveFlow(2, 64) for _ in range(6)])}
    torch.set_num_threads(os.cpu_count() or 1)
    for flow in ("realnvp", "spline", "maf", "iaf"):
        torch.manual_seed(0)
        ref_cpu = ref_build[flow]().eval()
        sd = ref_cpu.state_dict()
        rec = {"flow": flow, "published_cpu_sps": PUBLISHED_CPU[flow], "params": C.count_params(ref_cpu)}
        rec["reference_cpu_harness_sps"] = float(C.samples_per_sec(ref_cpu))           # host cores, n = 4000
        with torch.device(dev):
            mine = C.build_model(flow).to(dev)
            mine.load_state_dict(sd)
            mine.eval()
            rec["b200_harness_sps_n4000"] = float(C.samples_per_sec(mine))             # the published recipe, unmodified
        rec["b200_synced_sps_n4000"] = synced_sps(mine, 4000, dev)
        rec["b200_synced_sps_n1M"] = synced_sps(mine, 1 << 20, dev)
        ref_gpu = ref_build[flow]()
        ref_gpu.load_state_dict(sd)
        ref_gpu = ref_gpu.to(dev).eval()
        rec["reference_eager_cuda_synced_sps_n4000"] = synced_sps(ref_gpu, 4000, dev)
        rec["reference_eager_cuda_synced_sps_n1M"] = synced_sps(ref_gpu, 1 << 20, dev, reps=2)
        with torch.no_grad():
            z = torch.randn(4000, 2, device=dev)
            a, _ = mine.forward(z)
            b, _ = ref_gpu.forward(z)
        rec["max_abs_diff_vs_reference_eager"] = float((a - b).abs().max())
        rec["host_cores"] = torch.get_num_threads()
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
