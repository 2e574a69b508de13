"""GPU: the tcgen05 tile primitive (TMEM operands, SWIZZLE_128B weight images, 3xTF32) against a float64 product."""
import numpy as np
import pytest
import torch

import nfb200 as N
N_ = N

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
    for passes, tol in ((3, 1e-6), (1, 2e-3)):
        d = torch.full((128, n_out), float("nan"), device="cuda")
        N._lib.call("nf_debug_tc_gemm128", ad.data_ptr(), img.data_ptr(), d.data_ptr(), n_out, passes, None, 1,
                    N._lib.stream())
        torch.cuda.synchronize()
        err = np.abs(d.cpu().numpy().astype(np.float64) - ref) / scale
        assert np.isfinite(err).all()
        assert err.max() < tol, f"passes={passes}: max scaled error {err.max():.3e}"
        if passes == 1:
            assert err.max() > 1e-6      # a single TF32 pass must NOT be fp32-accurate (sanity of the test itself)


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (1000, 512, 64), (4096, 512, 512), (777, 130, 100), (300, 64, 1024),
                                   (257, 22736 // 8, 128)])
@pytest.mark.parametrize("relu", [False, True])
def test_linear_tc_matches_float64(M, N, K, relu):
    gen = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=gen) * 1.5
    w = torch.randn(N, K, generator=gen) / K ** 0.5
    b = torch.randn(N, generator=gen)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    hi, lo = N_.ops.split_tf32(wd)
    assert torch.equal(hi.cpu().view(torch.int32) & 0x1FFF, torch.zeros(N, K, dtype=torch.int32))
    assert torch.equal(lo.cpu().view(torch.int32) & 0x1FFF, torch.zeros(N, K, dtype=torch.int32))
    assert bool((((hi + lo).cpu().double() - w.double()).abs() <= 2.0 ** -23 * w.double().abs()).all())
    hn, ln = N_.packing.split_tf32(w.numpy())
    assert torch.equal(hi.cpu(), torch.from_numpy(hn)) and torch.equal(lo.cpu(), torch.from_numpy(ln))
    y = N_.ops.linear_tc(xd, hi, lo, bd, relu)
    assert y is not None
    ref = x.double() @ w.double().T + b.double()
    scale = x.double().abs() @ w.double().abs().T + b.double().abs()
    if relu:
        ref = ref.clamp_min(0)
    err = ((y.cpu().double() - ref).abs() / scale).max().item()
    assert err < 2e-6, f"scaled error {err:.3e}"      # K up to 1024: fp32 accumulation in TMEM adds ~sqrt(K) ulp
    # and against the FP32-pipe GEMM (same inputs): both are fp32-accurate
    y2 = N_.ops.linear_raw(xd, wd, bd, relu)
    assert ((y - y2).abs() / scale.cuda().float()).max().item() < 2e-6


def test_linear_tc_k_extent_matches_masked_dense():
    gen = torch.Generator().manual_seed(0)
    M, N, K = 512, 256, 256
    x = torch.randn(M, K, generator=gen)
    w = torch.randn(N, K, generator=gen)
    ext = torch.tensor([64, 100, 200, 256], dtype=torch.int32)
    for t in range(4):
        w[t * 64:(t + 1) * 64, int(ext[t]):] = 0
    hi, lo = N_.ops.split_tf32(w.cuda())
    y = N_.ops.linear_tc(x.cuda(), hi, lo, None, False, ext.cuda())
    ref = x.double() @ w.double().T
    scale = x.double().abs() @ w.double().abs().T + 1e-30
    err = ((y.cpu().double() - ref).abs() / scale).max().item()
    assert err < 2e-6, f"scaled error {err:.3e}"
