"""Parity of every CUDA kernel against the CPU oracle, through the C ABI (ctypes).  Run on the B200 box: -m gpu.

Tolerances: fp32 kernels |a-b| <= 1e-5 * max(|b|, 1); IPE features additionally get the conditioning
term 2^l * ulp32(x) * exp(-0.5 * 4^l * var) (sin(2^l x) amplifies a 1-ulp difference in x by 2^l)."""
import numpy as np
import pytest
import torch

from oracle import durf_oracle as O
import durf_test_helpers as H

pytestmark = pytest.mark.gpu


def _ops():
    from durf_b200 import ops
    return ops


def test_library_reports_version_and_counts_launches():
    ops = _ops()
    from durf_b200 import _lib
    assert b"sm_100a" in _lib.load().durf_version()
    ops.reset_launch_count()
    ops.viewdir_enc(torch.randn(8, 3, device='cuda'))
    assert ops.launch_count() == 1


def test_errors_are_reported_not_swallowed():
    ops = _ops()
    from durf_b200._lib import DurfError
    with pytest.raises(DurfError):
        ops.viewdir_enc(torch.randn(8, 3))                       # CPU tensor: no CPU path
    with pytest.raises(AssertionError):
        ops.raymarch(torch.zeros(1, 3, device='cuda'), torch.ones(1, 3, device='cuda'), torch.ones(1, device='cuda'), 128,
                     t_vals=torch.zeros(1, 129, device='cuda'), ray_shape='sphere')
    with pytest.raises(DurfError):
        ops.resample(torch.zeros(2, 300, device='cuda'), torch.zeros(2, 299, device='cuda'))   # N > 128


def test_viewdir_enc():
    ops = _ops()
    v = torch.nn.functional.normalize(torch.randn(1000, 3), dim=-1)
    H.assert_close(ops.viewdir_enc(v.cuda(), 4), O.pos_enc(v, 0, 4, True), what="pos_enc")


def test_obb_frontend_and_aa2matrix():
    ops = _ops()
    sc = H.scene(B=4096, K=3, seed=11)
    rays = H.oracle_rays(sc)
    box = torch.from_numpy(sc['centers'][2])
    ext = torch.from_numpy(sc['ext'])
    B, K = 4096, 3
    R = O.aa2matrix(box[:, 3:])
    H.assert_close(ops.aa2matrix(box[:, 3:].cuda()), R, what="aa2matrix")
    oo, do = O.world2object_rpy(rays.origins, rays.directions, box[:, :3].expand(B, K, 3), R.expand(B, K, 3, 3))
    zi, zo, hit = O.ray_box_intersection(oo, do, -ext.expand(B, K, 3), ext.expand(B, K, 3))
    got = ops.obb_frontend(rays.origins.cuda(), rays.directions.cuda(), box.cuda(), ext.cuda(), want_object_rays=True)
    H.assert_close(got['origins_o'], oo, what="origins_o")
    H.assert_close(got['dirs_o'], do, what="dirs_o")
    assert torch.equal(got['hit'].cpu(), hit), "intersection mask must be bit-exact"
    assert 0 < int(hit.sum()) < B
    H.assert_close(got['zi'], zi, what="zi"); H.assert_close(got['zo'], zo, what="zo")
    hf = hit.float()
    bk = (hit.sum(-1) == 0).float()
    H.assert_close(got['origins_s'], (oo * hf[..., None]).sum(-2) + bk[:, None] * rays.origins, what="origins_s")
    H.assert_close(got['dirs_s'], (do * hf[..., None]).sum(-2) + bk[:, None] * rays.directions, what="dirs_s")
    H.assert_close(got['zo_ret'], (hf * zo).sum(-1), what="zo_ret")
    for k in range(K):
        idx, cnt = ops.compact_hits(got['hit'], k)
        n = int(cnt.item())
        assert n == int(hit[:, k].sum())
        assert sorted(idx[:n].cpu().tolist()) == torch.nonzero(hit[:, k]).flatten().tolist()


def _ipe_tolerance(means, covd, min_deg, max_deg, weighted):
    """per-feature tolerance: 1e-5 + 2 * 2^l * ulp32(x) * exp(-0.5 * 4^l var)."""
    D = max_deg - min_deg
    sc = 2.0 ** torch.arange(min_deg, max_deg, dtype=torch.float64)
    ulp = torch.from_numpy(np.spacing(np.abs(means.numpy()).astype(np.float32)).astype(np.float64))
    ulp = torch.maximum(ulp, torch.tensor(np.spacing(np.float32(1.57))).double())        # y + pi/2 rounds at ulp(pi/2) too
    cond = (sc[:, None] * ulp[..., None, :]).reshape(*means.shape[:-1], 3 * D)
    damp = torch.exp(-0.5 * (sc[:, None] ** 2 * covd.double()[..., None, :]).reshape(*means.shape[:-1], 3 * D))
    t = 1e-5 + 2.0 * torch.cat([cond * damp, cond * damp], -1)
    if weighted:
        t = torch.cat([torch.full((*means.shape[:-1], 3), 1e-5, dtype=torch.float64), t], -1)
    return t


@pytest.mark.parametrize("randomized", [False, True])
@pytest.mark.parametrize("contract", [False, True])
def test_raymarch_sample_cast_contract_ipe(randomized, contract):
    ops = _ops()
    sc = H.scene(B=512, seed=5, far=200.0)
    rays = H.oracle_rays(sc)
    t_rand = torch.from_numpy(sc['t_rand'])
    t_vals, (mean, cov) = O.sample_along_rays(rays.origins, rays.directions, rays.radii, 128, rays.near, rays.far, randomized,
                                              t_rand=t_rand)
    if contract:
        mean, cov = O.new_space((mean, cov))
    want = O.integrated_pos_enc((mean, cov), 0, 10)
    cr = H.cuda_rays(sc)
    got = ops.raymarch(cr.origins, cr.directions, cr.radii, 128, near=cr.near, far=cr.far,
                       t_rand=t_rand.cuda() if randomized else None, contract=contract, want_gaussians=True)
    H.assert_close(got['t_vals'], t_vals, what="t_vals")
    H.assert_close(got['means'], mean, what="means")
    covd = torch.diagonal(cov, dim1=-2, dim2=-1)
    H.assert_close(got['cov_diag'], covd, rtol=2e-5, atol_scale=1e-3, what="cov_diag")
    tol = _ipe_tolerance(mean, covd, 0, 10, False)
    err = (got['features'].double().cpu() - want.double()).abs()
    assert bool((err <= tol).all()), f"IPE: worst excess {float((err - tol).max()):.3e}"
    # bf16 tile-image output carries the same values rounded to bf16
    tiles = ops.raymarch(cr.origins, cr.directions, cr.radii, 128, t_vals=got['t_vals'], contract=contract, bf16_tiles=True)
    dec = H.unswizzle_tiles(tiles['features'], 512 * 128, 60).reshape(512, 128, 60)
    assert float((dec - got['features'].cpu()).abs().max()) <= 2 ** -8 * 1.01 + 1e-6


def test_raymarch_weighted_ipe_compacted_and_masked():
    ops = _ops()
    sc = H.scene(B=300, seed=9)
    rays = H.oracle_rays(sc)
    t_vals = O.sample_t_vals(rays.near, rays.far, 128, True, torch.from_numpy(sc['t_rand']))
    mask = (torch.arange(300) % 3 == 0).float()
    mean, cov = O.cast_rays(t_vals, rays.origins, rays.directions, rays.radii)
    m3 = mask[:, None, None]
    for alpha in (0.0, 3.5, 10.0):
        want = O.weighted_ipe((m3 * mean, m3[..., None] * cov), 0, 10, alpha)
        cr = H.cuda_rays(sc)
        got = ops.raymarch(cr.origins, cr.directions, cr.radii, 128, t_vals=t_vals.cuda(), weighted=True, alpha=alpha,
                           ray_mult=mask.cuda(), want_gaussians=True)
        tol = _ipe_tolerance(m3 * mean, torch.diagonal(m3[..., None] * cov, dim1=-2, dim2=-1), 0, 10, True)
        err = (got['features'].double().cpu() - want.double()).abs()
        assert bool((err <= tol).all()), f"weighted IPE alpha={alpha}: worst excess {float((err - tol).max()):.3e}"
    # compaction: rows follow ray_index
    idx = torch.nonzero(mask).flatten().int().cuda()
    comp = ops.raymarch(cr.origins, cr.directions, cr.radii, 128, t_vals=t_vals.cuda(), weighted=True, alpha=10.0,
                        ray_index=idx, rows=idx.numel())
    full = ops.raymarch(cr.origins, cr.directions, cr.radii, 128, t_vals=t_vals.cuda(), weighted=True, alpha=10.0)
    assert torch.equal(comp['features'], full['features'][idx.long()])


def test_raymarch_cylinder_and_no_integration_and_ragged_n():
    ops = _ops()
    sc = H.scene(B=64, seed=4, N=48)
    rays = H.oracle_rays(sc)
    t_vals = O.sample_t_vals(rays.near, rays.far, 48, False)
    mean, cov = O.cast_rays(t_vals, rays.origins, rays.directions, rays.radii, 'cylinder')
    cr = H.cuda_rays(sc)
    got = ops.raymarch(cr.origins, cr.directions, cr.radii, 48, near=cr.near, far=cr.far, ray_shape='cylinder', want_gaussians=True)
    H.assert_close(got['t_vals'], t_vals, what="t_vals N=48")
    H.assert_close(got['means'], mean, what="cyl means")
    H.assert_close(got['cov_diag'], torch.diagonal(cov, dim1=-2, dim2=-1), rtol=2e-5, atol_scale=1e-3, what="cyl cov")
    got0 = ops.raymarch(cr.origins, cr.directions, cr.radii, 48, t_vals=t_vals.cuda(), integrate=False, want_gaussians=True)
    assert float(got0['cov_diag'].abs().max()) == 0.0
    empty = ops.raymarch(cr.origins[:0], cr.directions[:0], cr.radii[:0], 48, near=cr.near[:0], far=cr.far[:0])
    assert empty['features'].shape == (0, 48, 60)


@pytest.mark.parametrize("white,rand", [(False, False), (True, False), (False, True)])
def test_composite_fwd_bwd(white, rand):
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    B, N = 777, 128
    raw_rgb = torch.randn(B, N, 3, generator=g)
    raw_den = torch.randn(B, N, generator=g) * 2
    raw_den[5] = 80.0          # saturating density: transmittance underflows
    raw_den[6] = -60.0         # empty ray
    t = torch.sort(torch.rand(B, N + 1, generator=g) * 30, dim=-1).values
    dirs = torch.randn(B, 3, generator=g)
    rr, rd = raw_rgb.clone().requires_grad_(True), raw_den.clone().requires_grad_(True)
    dd = dirs.clone().requires_grad_(True)
    out = O.volumetric_rendering(torch.sigmoid(rr), torch.nn.functional.softplus(rd[..., None] - 1.0), t, dd, white, rand)
    got = ops.composite(raw_rgb.cuda(), raw_den.cuda(), t.cuda(), dirs.cuda(), white_bkgd=white, rand_bkgd=rand)
    for name, w in zip(('comp_rgb', 'depth', 'acc', 'weights'), out[:4]):
        H.assert_close(got[name], w, what=name)
    H.assert_close(got['t_mids'], out[5], what="t_mids"); H.assert_close(got['t_dists'], out[6], what="t_dists")
    # already-activated signature of mip.volumetric_rendering
    from durf_b200 import mip
    got2 = mip.volumetric_rendering(torch.sigmoid(raw_rgb).cuda(), torch.nn.functional.softplus(raw_den - 1.0).cuda()[..., None],
                                    t.cuda(), dirs.cuda(), white, rand, None)
    H.assert_close(got2[0], out[0], what="activated comp_rgb")
    # backward against autograd
    g_rgb, g_dep, g_w = torch.randn(B, 3, generator=g), torch.randn(B, generator=g), torch.randn(B, N, generator=g)
    loss = (out[0] * g_rgb).sum() + (out[1] * g_dep).sum() + (out[3] * g_w).sum()
    loss.backward()
    b_rgb, b_den, b_dirs = ops.composite_bwd(raw_rgb.cuda(), raw_den.cuda(), t.cuda(), dirs.cuda(), g_rgb.cuda(), g_dep.cuda(),
                                             g_w.cuda(), white_bkgd=white, rand_bkgd=rand, want_d_dirs=True)
    scale = float(rd.grad.abs().max())
    H.assert_close(b_den, rd.grad, rtol=2e-5, atol_scale=scale, what="d_raw_density")
    H.assert_close(b_rgb, rr.grad, rtol=2e-5, atol_scale=float(rr.grad.abs().max()), what="d_raw_rgb")
    H.assert_close(b_dirs, dd.grad, rtol=5e-5, atol_scale=float(dd.grad.abs().max()), what="d_dirs")


@pytest.mark.parametrize("randomized", [False, True])
def test_resample(randomized):
    ops = _ops()
    g = torch.Generator().manual_seed(8)
    B, N = 1024, 128
    t = torch.sort(torch.rand(B, N + 1, generator=g) * 40, dim=-1).values
    w = torch.rand(B, N, generator=g) ** 4
    w[0] = 0.0                       # all-zero weights: the eps padding path
    w[1] = 0.0; w[1, 77] = 1.0       # delta
    u = torch.rand(B, N + 1, generator=g)
    want = O.resample_t_vals(t, w, randomized, 0.01, u_rand=u if randomized else None)
    got = ops.resample(t.cuda(), w.cuda(), u_rand=u.cuda() if randomized else None).cpu()
    assert bool((got[:, 1:] >= got[:, :-1]).all()), "resampled fenceposts must be sorted"
    assert bool((got >= t[:, :1]).all() and (got <= t[:, -1:]).all())
    # parity in position or, where the bin mass makes the position ill-conditioned, in CDF space (see the helper)
    wp = torch.cat([w[:, :1], w, w[:, -1:]], -1)
    wmax = torch.maximum(wp[:, :-1], wp[:, 1:])
    wblur = 0.5 * (wmax[:, :-1] + wmax[:, 1:]) + 0.01
    frac_pos = H.assert_samples_close(got, want, t, wblur, what="resample")
    assert frac_pos > 0.999, f"only {frac_pos:.5f} of the samples agree to 1e-5 in position"


def test_sampler_reference_properties_on_gpu():
    """math_test.py:327-346 (single_bin) and :183-268 (sortedness) run against the CUDA sampler."""
    from durf_b200 import math as dmath
    bins = torch.tensor([[0, 1, 3, 6, 10]], dtype=torch.float32, device='cuda')
    for randomized in (False, True):
        for i in range(4):
            w = torch.zeros(1, 4, device='cuda'); w[0, i] = 1.0
            s = dmath.sorted_piecewise_constant_pdf(None, bins, w, 625, randomized)[0]
            assert bool((s >= bins[0, i]).all() and (s <= bins[0, i + 1]).all())
    g = torch.Generator().manual_seed(1)
    b = torch.sort(torch.randn(64, 17, generator=g) * 3, dim=-1).values
    w = torch.clamp(torch.rand(64, 16, generator=g) - 0.3, min=0)
    u = torch.rand(64, 4000, generator=g)
    for randomized in (False, True):
        want = O.sorted_piecewise_constant_pdf(b, w, 4000, randomized, u_rand=u)
        got = dmath.sorted_piecewise_constant_pdf(u.cuda(), b.cuda(), w.cuda(), 4000, randomized).cpu()
        assert bool((got[:, 1:] >= got[:, :-1]).all())
        H.assert_samples_close(got, want, b, w, what=f"sorted_piecewise_constant_pdf(randomized={randomized})")


def _mlp_inputs(topo, M, N, seed, bias_scale=0.1):
    rng = np.random.default_rng(seed)
    from durf_b200 import synthetic as S
    layers = S.glorot_mlp(rng, topo[0], topo[1], bias_scale)
    x = rng.uniform(-1, 1, size=(M, N, topo[0])).astype(np.float32)
    cond = rng.uniform(-1, 1, size=(M, 27)).astype(np.float32)
    return layers, torch.from_numpy(x), torch.from_numpy(cond)


def _blob(topo, layers):
    ops = _ops()
    blob = torch.zeros(ops.mlp_param_count(topo), device='cuda')
    for (w, b), (kw, kb) in zip(ops.mlp_layer_views(topo, blob), layers):
        w.copy_(torch.from_numpy(kw)); b.copy_(torch.from_numpy(kb))
    return blob


@pytest.mark.parametrize("topo", [(60, 256, 8, 4, 27, 128), (63, 128, 8, 4, 27, 128)])
def test_mlp_fp32_forward_backward(topo):
    ops = _ops()
    from durf_b200 import _lib
    M, N = 37, 16
    layers, x, cond = _mlp_inputs(topo, M, N, 21)
    ot = O.MLPTopology(*topo)
    params = [(torch.from_numpy(k).requires_grad_(True), torch.from_numpy(b).requires_grad_(True)) for k, b in layers]
    xr = x.clone().requires_grad_(True)
    want_rgb, want_den = O.mlp_apply(params, ot, xr, cond)
    blob = _blob(topo, layers)
    rgb, den, saved = ops.mlp_fwd(topo, x.cuda().reshape(M * N, -1), cond.cuda(), blob, M=M, N=N, precision=_lib.PREC_FP32, save=True)
    H.assert_close(rgb, want_rgb, what="raw_rgb"); H.assert_close(den, want_den[..., 0], what="raw_density")
    g_rgb, g_den = torch.randn(M, N, 3), torch.randn(M, N)
    ((want_rgb * g_rgb).sum() + (want_den[..., 0] * g_den).sum()).backward()
    d_blob = torch.zeros_like(blob)
    dfeat = ops.mlp_bwd(topo, x.cuda().reshape(M * N, -1), cond.cuda(), blob, saved, g_rgb.cuda(), g_den.cuda(), d_blob, M=M, N=N,
                        want_d_features=True)
    H.assert_close(dfeat.reshape(M, N, -1), xr.grad, rtol=2e-5, atol_scale=float(xr.grad.abs().max()), what="d_features")
    for i, ((dw, db), (pk, pb)) in enumerate(zip(ops.mlp_layer_views(topo, d_blob), params)):
        H.assert_close(dw, pk.grad, rtol=3e-5, atol_scale=float(pk.grad.abs().max()), what=f"dW{i}")
        H.assert_close(db, pb.grad, rtol=3e-5, atol_scale=float(pb.grad.abs().max()), what=f"db{i}")


@pytest.mark.parametrize("topo,M", [((60, 256, 8, 4, 27, 128), 5), ((60, 256, 8, 4, 27, 128), 300), ((63, 128, 8, 4, 27, 128), 301)])
def test_mlp_tensor_core_forward(topo, M):
    """tcgen05 chain vs the fp32 oracle: relative Frobenius error of the raw outputs <= 2e-2 (bf16 operands, fp32
    accumulation), and vs an oracle fed the same bf16-rounded inputs/weights much tighter."""
    ops = _ops()
    from durf_b200 import _lib
    N = 128
    layers, x, cond = _mlp_inputs(topo, M, N, 33)
    ot = O.MLPTopology(*topo)
    want_rgb, want_den = O.mlp_apply([(torch.from_numpy(k), torch.from_numpy(b)) for k, b in layers], ot, x, cond)
    blob = _blob(topo, layers)
    packed = ops.mlp_pack(topo, blob)
    # features as tile images: reuse the ray-march packer by writing them through torch (same swizzle as helpers.unswizzle)
    xb = torch.zeros(M * N, 64)
    xb[:, :topo[0]] = x.reshape(M * N, -1)
    tiles = torch.empty(M, 128 * 64, dtype=torch.bfloat16)
    r = torch.arange(128)
    xt = xb.reshape(M, 128, 64).to(torch.bfloat16)
    for c in range(8):
        off = (r // 8) * 512 + (r % 8) * 64 + ((c ^ (r % 8)) * 8)
        idx = (off[:, None] + torch.arange(8)[None, :]).reshape(-1)
        tiles[:, idx] = xt[:, :, c * 8:(c + 1) * 8].reshape(M, -1)
    rgb, den, _ = ops.mlp_fwd(topo, tiles.cuda(), cond.cuda(), blob, M=M, N=N, precision=_lib.PREC_BF16, packed=packed)
    torch.cuda.synchronize()
    rgb, den = rgb.cpu(), den.cpu()
    assert torch.isfinite(rgb).all() and torch.isfinite(den).all()
    rel = lambda a, b: float((a - b).norm() / b.norm())
    assert rel(rgb, want_rgb) <= 2e-2, f"raw_rgb rel Frobenius {rel(rgb, want_rgb):.3e}"
    assert rel(den, want_den[..., 0]) <= 2e-2, f"raw_density rel Frobenius {rel(den, want_den[..., 0]):.3e}"
    # against the oracle with the kernel's bf16 rounding points made explicit (operands bf16, accumulation / biases / heads
    # fp32): what is left is the summation order of the fp32 accumulation and rare 1-ulp bf16 rounding flips -> 2e-3
    emu_rgb, emu_den = H.mlp_apply_bf16_emulated([(torch.from_numpy(k), torch.from_numpy(b)) for k, b in layers], ot, x, cond)
    assert rel(rgb, emu_rgb) <= 2e-3, f"raw_rgb vs bf16-emulated oracle: rel Frobenius {rel(rgb, emu_rgb):.3e}"
    assert rel(den, emu_den[..., 0]) <= 2e-3, f"raw_density vs bf16-emulated oracle: rel Frobenius {rel(den, emu_den[..., 0]):.3e}"


def test_mlp_tensor_core_forward_edge_counts():
    """CTA pairs walk the tiles in lock step: a lone tile (the pair's second CTA has nothing to store), a device-side
    count below M, and count == 0 (nothing may be written) must all behave; tiles are independent of their batch."""
    ops = _ops()
    from durf_b200 import _lib
    topo, N, M = (60, 256, 8, 4, 27, 128), 128, 7
    layers, x, cond = _mlp_inputs(topo, M, N, 77)
    blob = _blob(topo, layers)
    packed = ops.mlp_pack(topo, blob)
    tiles = _tile_images(x, M, topo[0]).cuda()
    rgb_all, den_all, _ = ops.mlp_fwd(topo, tiles, cond.cuda(), blob, M=M, N=N, precision=_lib.PREC_BF16, packed=packed)
    rgb_1, den_1, _ = ops.mlp_fwd(topo, tiles[:1].contiguous(), cond[:1].cuda(), blob, M=1, N=N, precision=_lib.PREC_BF16, packed=packed)
    assert torch.equal(rgb_1[0], rgb_all[0]) and torch.equal(den_1[0], den_all[0])
    for cnt in (0, 3):
        count = torch.tensor([cnt], dtype=torch.int32, device="cuda")
        idx = torch.arange(M, dtype=torch.int32, device="cuda")
        rgb = torch.full((M, N, 3), 7.0, device="cuda")
        den = torch.full((M, N), 7.0, device="cuda")
        ops.mlp_fwd(topo, tiles, cond.cuda(), blob, M=M, N=N, precision=_lib.PREC_BF16, packed=packed, ray_index=idx, count=count,
                    raw_rgb=rgb, raw_density=den, num_rays_out=M)
        torch.cuda.synchronize()
        assert torch.equal(rgb[:cnt], rgb_all[:cnt]) and torch.equal(den[:cnt], den_all[:cnt])
        assert bool((rgb[cnt:] == 7.0).all()) and bool((den[cnt:] == 7.0).all())


def _tile_images(x, M, F):
    """[M,128,F] float features -> bf16 128x64 SWIZZLE_128B tile images (what the ray-march kernel writes)."""
    xb = torch.zeros(M * 128, 64)
    xb[:, :F] = x.reshape(M * 128, -1)
    tiles = torch.empty(M, 128 * 64, dtype=torch.bfloat16)
    r = torch.arange(128)
    xt = xb.reshape(M, 128, 64).to(torch.bfloat16)
    for c in range(8):
        off = (r // 8) * 512 + (r % 8) * 64 + ((c ^ (r % 8)) * 8)
        idx = (off[:, None] + torch.arange(8)[None, :]).reshape(-1)
        tiles[:, idx] = xt[:, :, c * 8:(c + 1) * 8].reshape(M, -1)
    return tiles


@pytest.mark.parametrize("topo,M", [((60, 256, 8, 4, 27, 128), 3), ((60, 256, 8, 4, 27, 128), 311), ((63, 128, 8, 4, 27, 128), 150)])
def test_mlp_tensor_core_backward(topo, M):
    """tcgen05 dgrad chain + wgrad kernel vs autograd of the oracle.
    (a) against the oracle with the tensor-core path's bf16 rounding points emulated (H.mlp_apply_bf16_emulated): the ReLU
        masks then agree, what is left is the bf16 rounding of dZ -> cosine >= 0.9995, relative Frobenius <= 3e-2;
    (b) against the plain fp32 oracle: a bf16 forward flips the ReLU mask of pre-activations near zero, and a fraction f
        of flipped entries costs ~sqrt(f) in Frobenius norm, accumulating down the chain -> cosine >= 0.985 (measured
        0.991 at the first layer, 0.999 at the last)."""
    ops = _ops()
    from durf_b200 import _lib
    N = 128
    layers, x, cond = _mlp_inputs(topo, M, N, 41)
    ot = O.MLPTopology(*topo)
    g = torch.Generator().manual_seed(3)
    d_rgb = torch.randn(M, N, 3, generator=g) * 0.1
    d_den = torch.randn(M, N, generator=g) * 0.1
    blob = _blob(topo, layers)
    packed = ops.mlp_pack(topo, blob)
    tiles = _tile_images(x, M, topo[0]).cuda()
    _, _, saved = ops.mlp_fwd(topo, tiles, cond.cuda(), blob, M=M, N=N, precision=_lib.PREC_BF16, packed=packed, save=True)
    d_blob = torch.zeros_like(blob)
    ops.mlp_bwd(topo, tiles, cond.cuda(), blob, saved, d_rgb.cuda(), d_den.cuda(), d_blob, M=M, N=N, precision=_lib.PREC_BF16,
                packed=packed)
    torch.cuda.synchronize()
    assert torch.isfinite(d_blob).all()
    for apply_fn, min_cos, max_rel, tag in ((H.mlp_apply_bf16_emulated, 0.9995, 3e-2, "bf16-emulated oracle"),
                                            (O.mlp_apply, 0.985, 0.2, "fp32 oracle")):
        params = [(torch.from_numpy(k).requires_grad_(True), torch.from_numpy(b).requires_grad_(True)) for k, b in layers]
        rgb, den = apply_fn(params, ot, x, cond)
        (rgb * d_rgb).sum().add((den[..., 0] * d_den).sum()).backward()
        for i, ((dw, db), (pk, pb)) in enumerate(zip(ops.mlp_layer_views(topo, d_blob), params)):
            for got, want, what in ((dw.cpu(), pk.grad, f"dW{i}"), (db.cpu(), pb.grad, f"db{i}")):
                cos = float((got * want).sum() / (got.norm() * want.norm() + 1e-30))
                rel = float((got - want).norm() / (want.norm() + 1e-30))
                assert cos >= min_cos and rel <= max_rel, f"{what} vs {tag}: cosine {cos:.5f}, rel Frobenius {rel:.3e}"


def test_generate_rays_matches_oracle_bit_exact():
    """Device ray generation vs the numpy restatement of _generate_rays_multi: float32 arithmetic in the same order, so the
    comparison is bit-exact (division and sqrt are IEEE, FMA contraction is off)."""
    ops = _ops()
    from durf_b200 import synthetic as S
    rng = np.random.default_rng(5)
    for (w, h, f) in ((97, 33, 120.25), (1920, 64, 2058.72)):
        c2w = S.random_c2w(rng)
        want = O.generate_rays(c2w, w, h, f, 0.5, 200.0)
        got = ops.generate_rays(c2w, w, h, f, 0.5, 200.0)
        for a, b, name in zip(got, want, O.Rays._fields):
            assert torch.equal(a.cpu().reshape(-1), torch.from_numpy(np.ascontiguousarray(b)).reshape(-1)), name
        pp = (w * 0.5 + 3.25, h * 0.5 - 1.75)                                        # Waymo variant: explicit principal point
        want_pp = O.generate_rays(c2w, w, h, f, 0.5, 200.0, principal_point=pp)
        got_pp = ops.generate_rays(c2w, w, h, f, 0.5, 200.0, principal_point=pp)
        for a, b, name in zip(got_pp, want_pp, O.Rays._fields):
            assert torch.equal(a.cpu().reshape(-1), torch.from_numpy(np.ascontiguousarray(b)).reshape(-1)), "principal point: " + name
        part = ops.generate_rays(c2w, w, h, f, 0.5, 200.0, row0=h - 3, row1=h)       # a row range incl. the last row
        assert torch.equal(part.radii.cpu().reshape(-1), torch.from_numpy(want.radii[h - 3:]).reshape(-1))
    empty = ops.generate_rays(c2w, 8, 8, 10.0, 0.0, 1.0, row0=4, row1=4)
    assert empty.origins.shape[0] == 0


def test_mlp_tensor_core_input_gradient():
    """BoxMLP (width 128) on the tensor cores: the dgrad chain's extra stage dX = dZ_0 W_0^T + dZ_skip W_skip[width:]^T
    vs autograd of the bf16-emulated oracle (cosine >= 0.9995, relative Frobenius <= 3e-2)."""
    ops = _ops()
    from durf_b200 import _lib
    topo, M, N = (63, 128, 8, 4, 27, 128), 37, 128
    layers, x, cond = _mlp_inputs(topo, M, N, 43)
    ot = O.MLPTopology(*topo)
    g = torch.Generator().manual_seed(5)
    d_rgb = torch.randn(M, N, 3, generator=g) * 0.1
    d_den = torch.randn(M, N, generator=g) * 0.1
    params = [(torch.from_numpy(k), torch.from_numpy(b)) for k, b in layers]
    xr = x.clone().requires_grad_(True)
    rgb, den = H.mlp_apply_bf16_emulated(params, ot, xr, cond)
    (rgb * d_rgb).sum().add((den[..., 0] * d_den).sum()).backward()
    blob = _blob(topo, layers)
    packed = ops.mlp_pack(topo, blob)
    tiles = _tile_images(x, M, topo[0]).cuda()
    _, _, saved = ops.mlp_fwd(topo, tiles, cond.cuda(), blob, M=M, N=N, precision=_lib.PREC_BF16, packed=packed, save=True)
    d_blob = torch.zeros_like(blob)
    dfeat = ops.mlp_bwd(topo, tiles, cond.cuda(), blob, saved, d_rgb.cuda(), d_den.cuda(), d_blob, M=M, N=N,
                        precision=_lib.PREC_BF16, packed=packed, want_d_features=True)
    torch.cuda.synchronize()
    got, want = dfeat.reshape(M, N, -1).cpu(), xr.grad
    assert torch.isfinite(got).all()
    cos = float((got * want).sum() / (got.norm() * want.norm()))
    rel = float((got - want).norm() / want.norm())
    assert cos >= 0.9995 and rel <= 3e-2, f"d_features: cosine {cos:.5f}, rel Frobenius {rel:.3e}"
    # the weight gradients are unchanged by asking for the input gradient
    d2 = torch.zeros_like(blob)
    ops.mlp_bwd(topo, tiles, cond.cuda(), blob, saved, d_rgb.cuda(), d_den.cuda(), d2, M=M, N=N, precision=_lib.PREC_BF16, packed=packed)
    assert float((d2 - d_blob).abs().max()) <= 1e-3 * float(d2.abs().max())
