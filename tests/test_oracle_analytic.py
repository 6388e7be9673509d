"""Analytic / Monte-Carlo / self-consistency checks of the oracle for the functions the reference
does not test (nothing tests mip.py, mip360.py, box_helpers.py, the model or the losses)."""
import math

import numpy as np
import pytest
import torch

from durf_b200 import synthetic as S
from oracle import durf_oracle as O


def test_conical_frustum_moments_monte_carlo():
    """The Gaussian of mip.py:99-130 matches sampled moments of a conical frustum."""
    rng = np.random.default_rng(0)
    d = np.array([0.3, -0.5, 1.1])
    t0, t1, r = 1.3, 2.1, 0.05
    n = 2_000_000
    # uniform in the frustum: t with density ~ t^2, disc radius r*t (in units of |d|-scaled perpendicular plane)
    u = rng.uniform(size=n)
    t = (t0 ** 3 + u * (t1 ** 3 - t0 ** 3)) ** (1 / 3)
    rho = np.sqrt(rng.uniform(size=n)) * r * t
    phi = rng.uniform(0, 2 * np.pi, size=n)
    dn = d / np.linalg.norm(d)
    a = np.cross(dn, [1, 0, 0]); a /= np.linalg.norm(a)
    b = np.cross(dn, a)
    # mip-NeRF measures t along the un-normalised d; the disc radius r*t is in world units
    pts = t[:, None] * d + rho[:, None] * (np.cos(phi)[:, None] * a + np.sin(phi)[:, None] * b)
    mean, cov = O.conical_frustum_to_gaussian(torch.tensor(d)[None], torch.tensor([[t0]], dtype=torch.float64),
                                              torch.tensor([[t1]], dtype=torch.float64), torch.tensor([[r]], dtype=torch.float64))
    np.testing.assert_allclose(mean[0, 0].numpy(), pts.mean(0), rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(cov[0, 0].numpy(), np.cov(pts.T), rtol=3e-2, atol=2e-4)


def test_new_space_closed_form():
    """mip360.py:63-79: cov' = cov * v_j^2 with v = (2/n - 1/n^2) + x * sum(x) * (-2/n^3 + 2/n^4) for n > 0.1, else 1."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(64, 16, 3, generator=g, dtype=torch.float64) * torch.logspace(-3, 2, 16, dtype=torch.float64)[None, :, None]
    a = torch.randn(64, 16, 3, 3, generator=g, dtype=torch.float64)
    cov = a @ a.transpose(-1, -2)
    mc, cc = O.new_space((x, cov))
    n = O.safe_norm(x)
    big = n > 0.1
    sx = x.sum(-1, keepdim=True)
    v = torch.where(big, (2 / n - 1 / n ** 2) + x * sx * (-2 / n ** 3 + 2 / n ** 4), torch.ones_like(x))
    np.testing.assert_allclose(cc.numpy(), (cov * (v ** 2)[..., None, :]).numpy(), rtol=1e-10, atol=1e-12)
    xc = torch.where(big, (2 - 1 / n) * x / n, x)
    np.testing.assert_allclose(mc.numpy(), xc.numpy(), rtol=1e-12)


def test_ipe_uses_only_cov_diagonal_and_barf_index_quirk():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(4, 8, 3, generator=g)
    a = torch.randn(4, 8, 3, 3, generator=g) * 0.1
    cov = a @ a.transpose(-1, -2)
    full = O.integrated_pos_enc((x, cov), 0, 10)
    diag = O.integrated_pos_enc((x, torch.diag_embed(torch.diagonal(cov, dim1=-2, dim2=-1))), 0, 10)
    np.testing.assert_allclose(full.numpy(), diag.numpy(), rtol=0, atol=0)
    l, dd = 3, 1
    want = torch.exp(-0.5 * 4.0 ** l * cov[..., dd, dd]) * torch.sin(2.0 ** l * x[..., dd])
    np.testing.assert_allclose(full[..., l * 3 + dd].numpy(), want.numpy(), rtol=1e-5, atol=1e-6)
    w = O.weighted_ipe((x, cov), 0, 10, alpha=3.5)
    bw = O.barf_weights(3.5, 10)
    np.testing.assert_allclose(w[..., :3].numpy(), x.numpy())
    for i in (0, 5, 6, 29, 30, 59):
        np.testing.assert_allclose(w[..., 3 + i].numpy(), (bw[i // 6] * full[..., i]).numpy(), rtol=1e-6, atol=1e-7)
    assert float(bw[2]) == 1.0 and abs(float(bw[3]) - 0.5) < 1e-6 and float(bw[4]) == 0.0


def test_aa2matrix_is_rotation_and_slab_test():
    aa = torch.tensor([[0.3, -0.2, 0.9], [0.0, 0.0, 0.0], [1e-7, 0, 0]], dtype=torch.float64)
    R = O.aa2matrix(aa)
    np.testing.assert_allclose((R @ R.transpose(-1, -2)).numpy(), np.tile(np.eye(3), (3, 1, 1)), atol=1e-9)
    o = torch.tensor([[[0.0, 0.0, -3.0]], [[0.0, 5.0, -3.0]], [[0.0, 0.0, 3.0]]])
    d = torch.tensor([[[1e-3, 1e-3, 1.0]], [[1e-3, 1e-3, 1.0]], [[1e-3, 1e-3, 1.0]]])
    d = d / d.norm(dim=-1, keepdim=True)
    zi, zo, hit = O.ray_box_intersection(o, d, -torch.ones(3, 1, 3), torch.ones(3, 1, 3))
    assert hit[:, 0].tolist() == [1, 0, 0] and hit.dtype == torch.int32
    assert abs(float(zi[0, 0]) - 2.0) < 1e-3 and abs(float(zo[0, 0]) - 4.0) < 1e-3


def _tiny_scene(B=96, K=2, dtype=torch.float32, seed=3, bias_scale=0.05):
    rng = np.random.default_rng(seed)
    rays, c2w = S.random_rays(rng, B, far=40.0)
    centers, ext = S.boxes_in_view(rng, c2w, K)
    params = dict(
        mlp=[(torch.from_numpy(k).to(dtype), torch.from_numpy(b).to(dtype)) for k, b in S.glorot_mlp(rng, 60, 256, bias_scale)],
        box_mlps=[[(torch.from_numpy(k).to(dtype), torch.from_numpy(b).to(dtype)) for k, b in S.glorot_mlp(rng, 63, 128, bias_scale)]
                  for _ in range(K)],
        box_centers=torch.from_numpy(centers).to(dtype))
    tr = O.Rays(*[torch.from_numpy(np.asarray(a)).to(dtype) for a in rays])
    tg = {k: torch.from_numpy(v).to(dtype) for k, v in S.targets(rng, B).items()}
    t_rand = torch.from_numpy(rng.uniform(size=(B, 129)).astype(np.float32)).to(dtype)
    u_rand = torch.from_numpy(rng.uniform(size=(B, 129)).astype(np.float32)).to(dtype)
    return params, tr, torch.from_numpy(ext).to(dtype), tg, t_rand, u_rand


def test_model_forward_shapes_masks_and_compaction_identity():
    params, rays, ext, tg, t_rand, u_rand = _tiny_scene()
    ret = O.model_forward(params, rays, ext, ts=2, randomized=True, rand_bkgd=False, white_bkgd=False, alpha=10.0,
                          t_rand=t_rand, u_rand=u_rand)
    assert len(ret) == 2
    B = rays.origins.shape[0]
    for lv in ret:
        assert lv.comp_rgb.shape == (B, 3) and lv.weights.shape == (B, 128) and lv.t_vals.shape == (B, 129)
        assert lv.dyn_mask.shape == (B, 1) and lv.zo.shape == (B,)
        assert torch.all(lv.t_vals[:, 1:] >= lv.t_vals[:, :-1])
        assert torch.isfinite(lv.comp_rgb).all()
    hits = int(ret[0].dyn_mask.sum())
    assert 0 < hits < B, "scene must exercise both object and background rays"
    assert float(ret[0].dyn_mask.max()) == 1.0, "no ray may cross two boxes (the reference sums the hits -> NaN)"


def test_loss_and_gradients_fp32_vs_fp64():
    out = {}
    for dtype in (torch.float32, torch.float64):
        params, rays, ext, tg, t_rand, u_rand = _tiny_scene(dtype=dtype)
        leaves = [t for kb in params['mlp'] for t in kb] + [params['box_centers']]
        for t in leaves:
            t.requires_grad_(True)
        cfg = O.ModelConfig(no_pose_opt=False, no_yaw_opt=False)
        ret = O.model_forward(params, rays, ext, ts=1, randomized=True, rand_bkgd=False, white_bkgd=False, alpha=4.5,
                              cfg=cfg, t_rand=t_rand, u_rand=u_rand)
        loss, stats = O.loss_fn(ret, rays, tg['pixels'], tg['depth'], tg['sky'], eps=3.0)
        grads = torch.autograd.grad(loss, leaves, allow_unused=True)
        out[dtype] = (float(loss), [g.double() for g in grads])
        assert math.isfinite(float(loss))
    l32, g32 = out[torch.float32]
    l64, g64 = out[torch.float64]
    assert abs(l32 - l64) < 2e-4 * max(1.0, abs(l64))
    # first-layer kernel gradient: cosine similarity between fp32 and fp64 oracles
    a, b = g32[0].flatten(), g64[0].flatten()
    assert float(a @ b / (a.norm() * b.norm())) > 0.999
    assert g64[-1].abs().sum() > 0, "pose gradient must be non-zero when pose/yaw optimisation is on"
    assert torch.all(g64[-1][[0, 2, 3, 4]] == 0), "only the selected timestep row receives gradient"


def test_distortion_prefix_sum_identity():
    """The O(N) form used by the CUDA loss kernel equals the dense [B,N,N] form of train_boxpose.py:146-151."""
    g = torch.Generator().manual_seed(5)
    w = torch.rand(7, 128, generator=g, dtype=torch.float64)
    s = torch.sort(torch.rand(7, 128, generator=g, dtype=torch.float64) * 40, dim=-1).values
    dense = (w[:, :, None] * w[:, None, :] * (s[:, :, None] - s[:, None, :]).abs()).sum()
    W = torch.cumsum(w, -1) - w
    WS = torch.cumsum(w * s, -1) - w * s
    fast = (2 * w * (s * W - WS)).sum()
    np.testing.assert_allclose(float(fast), float(dense), rtol=1e-12)


def test_adam_and_grad_postprocess():
    g = [torch.tensor([float('nan'), float('inf'), -float('inf'), 0.5, -0.01])]
    gs, norm = O.postprocess_grads(g)
    want = torch.tensor([0.0, 0.0, -0.1, 0.1, -0.01])
    n = want.norm()
    np.testing.assert_allclose(gs[0].numpy(), (want * min(1.0, 1.0 / (1e-7 + float(n)))).numpy(), rtol=1e-6)
    p, m, v = O.adam_step([torch.ones(3)], [torch.full((3,), 0.1)], [torch.zeros(3)], [torch.zeros(3)], step=0, lr=1e-3)
    np.testing.assert_allclose(p[0].numpy(), np.full(3, 1 - 1e-3), rtol=1e-5)


def test_ray_generation_oracle_matches_synthetic_frame_rays():
    """oracle.generate_rays (obbpose_dataset.py:613-661) and the benchmark's synthetic.frame_rays are the same rays."""
    from durf_b200 import synthetic as S
    rng = np.random.default_rng(3)
    c2w = S.random_c2w(rng)
    want = O.generate_rays(c2w, 96, 40, 103.5, 0.0, 40.0)
    got = S.frame_rays(c2w, width=96, height=40, focal=103.5, far=40.0)
    for a, b, name in zip(got, want, O.Rays._fields):
        assert np.array_equal(np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)), name
    assert float(np.abs(np.linalg.norm(want.viewdirs, axis=-1) - 1).max()) < 1e-6
    # quirk kept: `np.concatenate([v, v[-2:-1]])` gives the last image row the spacing of row H-3, not H-2 (:643)
    assert np.array_equal(want.radii[-1], want.radii[-3])
