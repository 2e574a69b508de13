"""GPU: the reference's OWN benchmark harness, unmodified, against the `src.*` drop-in shim.

  * plots/_common.py (loaded from oracle/_ref/plots/_common.py, byte-identical copy staged by oracle/make_ref.py):
    build_model for the four published configs (plots/_common.py:158-170), samples_per_sec (:264-274), save_cache /
    load_cache (:279-303), model_samples, reconstruction_error, train -- with `from src.models import ...` resolving to
    the B200 implementation and torch's default device set to CUDA (the harness creates its tensors itself);
  * src/flows/utils/profiling.py::FlowProfiler.profile_flow (:63-100), loaded from oracle/_ref with its relative import
    `..flow.flow.Flow` bound to the shim's Flow.
Skipped when oracle/_ref has not been staged."""
import importlib.util
import os
import sys
import types

import pytest
import torch

import nfb200 as N
from oracle import ref_loader

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _need_ref():
    if not ref_loader.available():
        pytest.skip("oracle/_ref not staged (python oracle/make_ref.py in the build container)")


def load_reference_plots_common(cache_dir=None):
    """plots/_common.py, executed unmodified; `src.*` is the repo's drop-in shim."""
    ref_loader.stub_missing()                   # matplotlib stub when matplotlib is not installed
    import src.flows  # noqa: F401  (the shim)
    import src.models  # noqa: F401
    assert "normalizing-flows-study_b200" in N.__file__ and src.models.RealNVP is N.RealNVP
    path = os.path.join(ref_loader.REF, "plots", "_common.py")
    spec = importlib.util.spec_from_file_location("reference_plots_common", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if cache_dir is not None:
        mod.CACHE = str(cache_dir)
    return mod


def load_reference_profiler():
    """src/flows/utils/profiling.py, executed unmodified inside a synthetic package whose `flow.flow.Flow` is the shim's."""
    pkg = types.ModuleType("refharness"); pkg.__path__ = []
    flow_pkg = types.ModuleType("refharness.flow"); flow_pkg.__path__ = []
    flow_mod = types.ModuleType("refharness.flow.flow"); flow_mod.Flow = N.Flow
    utils_pkg = types.ModuleType("refharness.utils"); utils_pkg.__path__ = []
    for m in (pkg, flow_pkg, flow_mod, utils_pkg):
        sys.modules[m.__name__] = m
    path = os.path.join(ref_loader.REF, "src", "flows", "utils", "profiling.py")
    spec = importlib.util.spec_from_file_location("refharness.utils.profiling", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("flow", ["realnvp", "spline", "maf", "iaf"])
def test_reference_harness_samples_per_sec_and_cache_roundtrip(flow, tmp_path):
    _need_ref()
    C = load_reference_plots_common(tmp_path)
    with torch.device(DEV):                     # the harness builds models / base distributions on the default device
        torch.manual_seed(0)
        model = C.build_model(flow).to(DEV)
        assert type(model).__module__.startswith("nfb200")
        with torch.no_grad():
            for p in model.parameters():
                p.add_(0.05 * torch.randn_like(p))
        model.eval()
        before = N._lib.launch_count()
        sps = C.samples_per_sec(model)          # plots/_common.py:264-274, n=4000
        assert N._lib.launch_count() > before and sps > 0
        x = C.model_samples(model, n=1000)
        assert x.shape == (1000, 2) and bool(torch.isfinite(x).all())
        data = C.get_dataset("moons", n=512).to(DEV)
        z, ld = model.inverse(data)
        x_rec, _ = model.forward(z)
        assert float((data - x_rec).abs().sum(1).mean()) < 1e-3
        # save_cache / load_cache: state_dict round trip through the harness (load_cache builds a fresh model)
        C.save_cache("moons", flow, model, [1.0, 0.9], 0.0)
        blob = C.load_cache("moons", flow)
        m2 = blob["model"]
        assert blob["params"] == C.count_params(model)
        z2, ld2 = m2.inverse(data)
        assert torch.equal(z, z2) and torch.equal(ld, ld2)
        # a few full-batch training steps of the harness's own loop
        model.train()
        curve = C.train(model, data, epochs=3, lr=1e-3)
        assert len(curve) == 3 and all(c == c for c in curve)


def test_reference_flow_profiler_runs_unmodified():
    _need_ref()
    P = load_reference_profiler()
    prof = P.FlowProfiler(warmup_iterations=2, measurement_iterations=3)
    # the reference's CUDA timing returns 0.0 for both stamps (profiling.py:183-186) and then divides by the mean time;
    # with its start_event unset the same unmodified code path times with perf_counter around torch.cuda.synchronize()
    prof.start_event = None
    flow = N.MaskedAutoregressiveFlow(8, 32)
    res = prof.profile_flow(flow, (8,), batch_sizes=[64, 4096], device=DEV, include_backward=False)
    for b, m in res.items():
        assert m.forward_throughput > 0 and m.inverse_throughput > 0 and m.parameters == sum(p.numel() for p in flow.parameters())
