"""The reference's own tests for this path (internal/math_test.py), run against the oracle.

These are the only known-answer / property tests the reference holds for the hot path
(SURVEY.md §4, §8c): they pin `sorted_piecewise_constant_pdf`, `safe_sin/safe_cos`,
`learning_rate_decay` and the PSNR round trip.  The JAX threefry draws are replaced by
numpy draws with the same distributions (the properties do not depend on the stream).
"""
import numpy as np
import scipy.special
import scipy.stats
import torch

from oracle import durf_oracle as O


def _trig_harness(fn, max_exp, dtype):
    # math_test.py:31-36
    x = 10 ** np.linspace(-30, max_exp, 10000)
    x = np.concatenate([-x[::-1], np.array([0]), x])
    xt = torch.from_numpy(x).to(dtype)
    y_true = getattr(np, fn)(xt.double().numpy())
    y = getattr(O, 'safe_' + fn)(xt).double().numpy()
    return y_true, y


def test_sin():
    """math_test.py:41-50: accurate on +-[1e-30, 1e10], never NaN up to 1e60."""
    for fn in ['sin', 'cos']:
        y_true, y = _trig_harness(fn, 10, torch.float64)
        assert np.max(np.abs(y - y_true)) < 1e-4
        assert not np.any(np.isnan(y))
    for fn in ['sin', 'cos']:
        _, y = _trig_harness(fn, 60, torch.float64)
        assert not np.any(np.isnan(y))
        _, y = _trig_harness(fn, 38, torch.float32)
        assert not np.any(np.isnan(y))


def test_sin_fp32_matches_reduced_argument():
    """fp32: below 100*pi the argument is untouched; above, it is the exact floored remainder by fl32(100*pi)."""
    x = torch.linspace(-5000, 5000, 200001, dtype=torch.float32)
    t32 = np.float32(100 * np.pi)
    xn = x.numpy().astype(np.float64)
    red = np.where(np.abs(xn) < t32, xn, xn - np.floor(xn / float(t32)) * float(t32))
    np.testing.assert_allclose(O.safe_sin(x).numpy(), np.sin(red), atol=2e-6)
    np.testing.assert_allclose(O.safe_cos(x).numpy(), np.cos(red), atol=2e-6)


def test_psnr_round_trip():
    """math_test.py:52-55."""
    mse = 0.07
    np.testing.assert_allclose(float(O.psnr_to_mse(O.mse_to_psnr(mse))), mse, rtol=1e-6)


def test_learning_rate_decay():
    """math_test.py:57-80."""
    np.random.seed(0)
    for _ in range(10):
        lr_init = np.exp(np.random.normal() - 3)
        lr_final = lr_init * np.exp(np.random.normal() - 5)
        max_steps = int(np.ceil(100 + 100 * np.exp(np.random.normal())))
        f = lambda s: O.learning_rate_decay(s, lr_init, lr_final, max_steps)
        np.testing.assert_allclose(f(0), lr_init, rtol=1e-6)
        np.testing.assert_allclose(f(max_steps), lr_final, rtol=1e-6)
        np.testing.assert_allclose(f(max_steps / 2), np.sqrt(lr_init * lr_final), rtol=1e-6)
        np.testing.assert_allclose(f(max_steps + 100), lr_final, rtol=1e-6)


def test_delayed_learning_rate_decay():
    """math_test.py:82-115."""
    np.random.seed(0)
    for _ in range(10):
        lr_init = np.exp(np.random.normal() - 3)
        lr_final = lr_init * np.exp(np.random.normal() - 5)
        max_steps = int(np.ceil(100 + 100 * np.exp(np.random.normal())))
        lr_delay_steps = int(np.random.uniform(low=0.1, high=0.4) * max_steps)
        lr_delay_mult = np.exp(np.random.normal() - 3)
        f = lambda s: O.learning_rate_decay(s, lr_init, lr_final, max_steps, lr_delay_steps, lr_delay_mult)
        np.testing.assert_allclose(f(0), lr_delay_mult * lr_init, rtol=1e-6)
        np.testing.assert_allclose(f(max_steps), lr_final, rtol=1e-6)
        np.testing.assert_allclose(f(lr_delay_steps), O.learning_rate_decay(lr_delay_steps, lr_init, lr_final, max_steps), rtol=1e-6)
        np.testing.assert_allclose(f(max_steps / 2), np.sqrt(lr_init * lr_final), rtol=1e-6)
        np.testing.assert_allclose(f(max_steps + 100), lr_final, rtol=1e-6)


def test_sorted_piecewise_constant_pdf_train_mode():
    """math_test.py:183-268: sampling reproduces its distribution (angle <= 0.5 deg, JS <= 1e-5)."""
    batch_size, num_bins, num_samples, precision = 4, 16, 1000000, 1e5
    rng = np.random.default_rng(20202020)
    data = []
    for _ in range(batch_size):
        bins_delta = np.round(precision * np.exp(rng.uniform(-3, 3, size=num_bins + 1)))
        keep = rng.uniform(size=bins_delta.shape) < 0.9
        keep[-1] = True  # jnp clamps the out-of-range index the reference's merge loop would hit; avoid that case
        bins_delta *= keep
        bins = np.cumsum(bins_delta) / precision
        bins += rng.normal() * num_bins / 2
        weights = np.maximum(0, rng.uniform(-0.5, 1.0, size=num_bins))
        data.append((bins, weights, weights / weights.sum()))
    data.append((bins, np.zeros_like(weights), np.ones_like(weights) / num_bins))
    bins, weights, gt_hist = [np.stack(x).astype(np.float32) for x in zip(*data)]

    for randomized in [True, False]:
        u_rand = torch.from_numpy(rng.uniform(size=(bins.shape[0], num_samples)).astype(np.float32))
        samples = O.sorted_piecewise_constant_pdf(torch.from_numpy(bins), torch.from_numpy(weights),
                                                  num_samples, randomized, u_rand=u_rand, chunk=1).numpy()
        assert samples.shape[-1] == num_samples
        assert np.all(samples[..., 1:] >= samples[..., :-1])
        for i_samples, i_bins, i_gt in zip(samples, bins, gt_hist):
            i_hist = np.float32(np.histogram(i_samples, i_bins)[0]) / num_samples
            i_gt = np.array(i_gt)
            while np.any(i_bins[:-1] == i_bins[1:]):
                j = int(np.where(i_bins[:-1] == i_bins[1:])[0][0])
                i_hist = np.concatenate([i_hist[:j], [i_hist[j] + i_hist[j + 1]], i_hist[j + 2:]])
                i_gt = np.concatenate([i_gt[:j], [i_gt[j] + i_gt[j + 1]], i_gt[j + 2:]])
                i_bins = np.concatenate([i_bins[:j], i_bins[j + 1:]])
            angle = 180 / np.pi * np.arccos(np.minimum(
                1., np.mean((i_hist * i_gt) / np.sqrt(np.mean(i_hist ** 2) * np.mean(i_gt ** 2)))))
            m = (i_hist + i_gt) / 2
            js = np.sum(scipy.special.kl_div(i_hist, m) + scipy.special.kl_div(i_gt, m)) / 2
            assert angle <= 0.5
            assert js <= 1e-5


def _large(delta: bool):
    num_samples, num_bins = 100, 100000
    rng = np.random.default_rng(0)
    bins = np.arange(num_bins, dtype=np.float32)
    weights = np.ones(num_bins - 1, np.float32)
    delta_idx = len(weights) // 2
    if delta:
        weights[delta_idx] = len(weights) - 1
    u_rand = torch.from_numpy(rng.uniform(size=(1, num_samples)).astype(np.float32))
    samples = O.sorted_piecewise_constant_pdf(torch.from_numpy(bins)[None], torch.from_numpy(weights)[None],
                                              num_samples, True, u_rand=u_rand)[0].numpy()
    assert np.all(samples >= bins[0]) and np.all(samples <= bins[-1])
    assert scipy.stats.kstest(np.mod(samples, 1), 'uniform', (0, 1)).statistic <= 0.2
    return samples, bins, delta_idx


def test_sorted_piecewise_constant_pdf_large_flat():
    """math_test.py:270-295."""
    samples, bins, _ = _large(False)
    assert scipy.stats.kstest(samples, 'uniform', (bins[0], bins[-1])).statistic <= 0.2


def test_sorted_piecewise_constant_pdf_sparse_delta():
    """math_test.py:297-325: the delta bin holds ~half the samples."""
    samples, bins, delta_idx = _large(True)
    in_delta = (samples >= bins[delta_idx]) & (samples <= bins[delta_idx + 1])
    np.testing.assert_allclose(np.mean(in_delta), 0.5, atol=0.05)


def test_sorted_piecewise_constant_pdf_single_bin():
    """math_test.py:327-346: the reference's only concrete vector, bins=[0,1,3,6,10] one-hot weights."""
    num_samples = 625
    rng = np.random.default_rng(0)
    bins = torch.tensor([0, 1, 3, 6, 10], dtype=torch.float32)
    for randomized in [False, True]:
        for i in range(len(bins) - 1):
            weights = np.zeros(len(bins) - 1, np.float32)
            weights[i] = 1.
            u_rand = torch.from_numpy(rng.uniform(size=(1, num_samples)).astype(np.float32))
            samples = O.sorted_piecewise_constant_pdf(bins[None], torch.from_numpy(weights)[None], num_samples,
                                                      randomized, u_rand=u_rand)[0]
            assert torch.all(samples >= bins[i]) and torch.all(samples <= bins[i + 1])
