"""Evaluation metrics of the reference's internal/math.py (PSNR / SSIM / sRGB), ported from internal/math_test.py:52-55,
117-181.  The reference checks SSIM against tf.image.ssim (not installed here): the independent check below is a direct
2-D Gaussian-window evaluation in numpy float64."""
import numpy as np
import torch

from durf_b200 import math as M


def _ssim_direct(img0, img1, max_val, fs, sigma, k1, k2):
    """Plain-loop SSIM with a full 2-D window (no separability), float64."""
    hw = fs // 2
    shift = (2 * hw - fs + 1) / 2
    f = np.exp(-0.5 * ((np.arange(fs) - hw + shift) / sigma) ** 2)
    f /= f.sum()
    w2 = np.outer(f, f)
    B, H, W, C = img0.shape
    oh, ow = H - fs + 1, W - fs + 1
    out = np.zeros((B, oh, ow, C))
    c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
    for b in range(B):
        for c in range(C):
            for y in range(oh):
                for x in range(ow):
                    p0 = img0[b, y:y + fs, x:x + fs, c]
                    p1 = img1[b, y:y + fs, x:x + fs, c]
                    m0, m1 = (w2 * p0).sum(), (w2 * p1).sum()
                    s00 = max(0.0, (w2 * p0 * p0).sum() - m0 * m0)
                    s11 = max(0.0, (w2 * p1 * p1).sum() - m1 * m1)
                    s01 = (w2 * p0 * p1).sum() - m0 * m1
                    s01 = np.sign(s01) * min(np.sqrt(s00 * s11), abs(s01))
                    out[b, y, x, c] = ((2 * m0 * m1 + c1) * (2 * s01 + c2)) / ((m0 * m0 + m1 * m1 + c1) * (s00 + s11 + c2))
    return out


def test_psnr_round_trip():
    """math_test.py:52-55."""
    mse = torch.tensor(0.07)
    assert torch.allclose(M.psnr_to_mse(M.mse_to_psnr(mse)), mse)
    assert abs(float(M.mse_to_psnr(torch.tensor(0.01))) - 20.0) < 1e-5


def test_ssim_matches_direct_evaluation():
    """math_test.py:117-161 with the independent evaluation standing in for tf.image.ssim."""
    rng = np.random.default_rng(0)
    for _ in range(4):
        max_val = rng.uniform(0.1, 3.0)
        img0 = max_val * rng.uniform(-1, 1, (2, 12, 12, 3))
        img1 = max_val * rng.uniform(-1, 1, (2, 12, 12, 3))
        fs = int(rng.integers(1, 10))
        sigma = rng.uniform(0.1, 10.0)
        k1, k2 = rng.uniform(0.001, 0.1), rng.uniform(0.001, 0.1)
        want = _ssim_direct(img0, img1, max_val, fs, sigma, k1, k2)
        got_map = M.compute_ssim(torch.from_numpy(img0), torch.from_numpy(img1), max_val, fs, sigma, k1, k2, return_map=True)
        got = M.compute_ssim(torch.from_numpy(img0), torch.from_numpy(img1), max_val, fs, sigma, k1, k2)
        np.testing.assert_allclose(got_map.numpy(), want, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(got.numpy(), want.mean(axis=(1, 2, 3)), rtol=1e-9, atol=1e-10)
        assert got_map.max() <= 1.0 + 1e-12 and got_map.min() >= -1.0 - 1e-12
        got32 = M.compute_ssim(torch.from_numpy(img0).float(), torch.from_numpy(img1).float(), max_val, fs, sigma, k1, k2)
        np.testing.assert_allclose(got32.numpy(), want.mean(axis=(1, 2, 3)), rtol=2e-4, atol=2e-5)


def test_ssim_lowerbound():
    """math_test.py:163-170: the corner case where SSIM is -1."""
    sz = 11
    img = np.meshgrid(*([np.linspace(-1, 1, sz)] * 2))[0][None, ..., None]
    ssim = M.compute_ssim(torch.from_numpy(img), torch.from_numpy(-img), 1., filter_size=sz, filter_sigma=1.5, k1=1e-5, k2=1e-5)
    np.testing.assert_allclose(ssim.numpy(), -np.ones_like(ssim.numpy()), rtol=1e-6, atol=1e-6)


def test_srgb_linearize():
    """math_test.py:172-181: round trips and finite gradients."""
    x = torch.linspace(-1, 3, 10000, dtype=torch.float64)
    assert torch.allclose(M.linear_to_srgb(M.srgb_to_linear(x)), x, rtol=1e-6, atol=1e-6)
    assert torch.allclose(M.srgb_to_linear(M.linear_to_srgb(x)), x, rtol=1e-6, atol=1e-6)
    for fn in (M.linear_to_srgb, M.srgb_to_linear):
        xr = x.clone().requires_grad_(True)
        fn(xr).sum().backward()
        assert torch.isfinite(xr.grad).all()


def test_avg_error_is_geometric_mean():
    """math.py:59-63."""
    psnr, ssim, lpips = 25.0, 0.8, 0.2
    want = (float(M.psnr_to_mse(psnr)) * (1 - ssim) ** 0.5 * lpips) ** (1 / 3)
    assert abs(float(M.compute_avg_error(psnr, ssim, lpips)) - want) < 1e-9
