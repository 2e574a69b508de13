"""GPU: the tcgen05 tile primitive (TMEM operands, SWIZZLE_128B weight images, 3xTF32) against a float64 product."""
import numpy as np
import pytest
import torch

import nfb200 as N

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_out", [16, 32, 64, 128])
def test_tc_gemm128_3xtf32(n_out):
    rng = np.random.default_rng(n_out)
    a = rng.standard_normal((128, 64)).astype(np.float32) * 2
    w = rng.standard_normal((n_out, 64)).astype(np.float32)
    img = torch.from_numpy(N.packing.umma_sw128_images(w)).cuda()
    ad = torch.from_numpy(a).cuda()
    ref = a.astype(np.float64) @ w.astype(np.float64).T
    scale = np.abs(a).astype(np.float64) @ np.abs(w).astype(np.float64).T
    for passes, tol in ((3, 2e-6), (1, 2e-3)):
        d = torch.full((128, n_out), float("nan"), device="cuda")
        N._lib.call("nf_debug_tc_gemm128", ad.data_ptr(), img.data_ptr(), d.data_ptr(), n_out, passes, None, 1,
                    N._lib.stream())
        torch.cuda.synchronize()
        err = np.abs(d.cpu().numpy().astype(np.float64) - ref) / scale
        assert np.isfinite(err).all()
        assert err.max() < tol, f"passes={passes}: max scaled error {err.max():.3e}"
        if passes == 1:
            assert err.max() > 1e-6      # a single TF32 pass must NOT be fp32-accurate (sanity of the test itself)
