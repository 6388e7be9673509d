"""The CUDA path against fixtures produced by EXECUTING THE REFERENCE'S OWN CODE (tests/golden/ref_*.npz, generator
tests/golden/make_ref_golden.py, see tests/_refshim/README.md).  Every SURVEY §8(a) row has a case here whose expected
values came from reference code.  Tolerances: fp32 kernels |a-b| <= 1e-5 * max(|b|, 1) (north_star), IPE features with
the 2^l ulp(x) conditioning term, the inverse-CDF sampler in position-or-CDF space, bf16 MLP path by relative Frobenius
error of the composited outputs and PSNR (stated per test).  Nothing here reads /root/reference.
"""
import numpy as np
import pytest
import torch

import durf_test_helpers as H
import ref_cases as C
from test_ref_golden import load, T, GRAD_TOL

pytestmark = pytest.mark.gpu


def _ops():
    from durf_b200 import ops
    return ops


def close(got, want, rtol=1e-5, scale=1.0, what=''):
    """|a-b| <= rtol * max(|b|, scale); NaN patterns must agree."""
    got = np.asarray(got.detach().cpu() if torch.is_tensor(got) else got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    ng, nw = np.isnan(got), np.isnan(want)
    assert np.array_equal(ng, nw), f"{what}: NaN pattern differs ({ng.sum()} vs {nw.sum()} NaNs)"
    err = np.abs(np.where(nw, 0, got - want))
    tol = rtol * np.maximum(np.abs(np.where(nw, 0, want)), scale)
    assert (err <= tol).all(), f"{what}: {(err > tol).sum()}/{err.size} beyond tol, worst excess {np.max(err - tol):.3e}"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------------------ R15: math.py sampler
def test_sampler_against_reference():
    from durf_b200 import math as dmath
    g, I = load('math'), C.math_inputs()
    bins, w, u = cu(I['bins']), cu(I['weights']), cu(I['u'])
    for randomized, key in ((False, 'pdf_det'), (True, 'pdf_rand')):
        got = dmath.sorted_piecewise_constant_pdf(u if randomized else None, bins, w, C.N + 1, randomized).cpu()
        assert bool((got[:, 1:] >= got[:, :-1]).all())
        frac = H.assert_samples_close(got, T(g[key]), T(I['bins']), T(I['weights']), what=key)
        assert frac > 0.995
    sb = torch.tensor([[0, 1, 3, 6, 10]], dtype=torch.float32, device='cuda')
    for i in range(4):
        sw = torch.zeros(1, 4, device='cuda'); sw[0, i] = 1.0
        got = dmath.sorted_piecewise_constant_pdf(None, sb, sw, 625, False)
        close(got, g[f'pdf_single_bin_{i}'], rtol=2e-6, what=f'single bin {i}')


# ------------------------------------------------------------------------------------------------ R1-R4: box_helpers.py
def test_obb_functions_against_reference():
    """aa2matrix / world2object_rpy / ray_box_intersection with the reference's own call shapes, and the fused front-end,
    on a 4-box scene with rays that cross two boxes and a zero-rotation box."""
    from durf_b200 import box_helpers as bh
    ops = _ops()
    g = load('obb')
    sc = C._scene(B=512, K=4, seed=111, overlap=True)
    assert str(g['inputs_sha256']) == C.scene_digest(sc)
    B, K = 512, 4
    box = cu(g['box'])
    R = bh.aa2matrix(box[:, 3:])
    close(R, g['aa2matrix'], what='aa2matrix')
    rays = H.cuda_rays(sc)
    pose = box[:, :3].expand(B, K, 3)
    oo, do = bh.world2object_rpy(rays.origins, rays.directions, pose, R.expand(B, K, 3, 3))
    close(oo, g['origins_o'], what='origins_o'); close(do, g['dirs_o'], what='dirs_o')
    oo2, do2 = bh.world2object_rpy(rays.origins, rays.directions, pose.contiguous(), R.expand(B, K, 3, 3).contiguous())
    assert torch.equal(oo, oo2) and torch.equal(do, do2)                                   # per-ray layout, same values
    ext = cu(sc['ext']).expand(B, K, 3)
    zi, zo, hit = bh.ray_box_intersection(cu(g['origins_o']), cu(g['dirs_o']), -ext, ext)
    assert np.array_equal(hit.cpu().numpy(), g['hit']), "intersection mask must be bit-exact"
    close(zi, g['zi'], what='zi'); close(zo, g['zo'], what='zo')
    # fused front-end (the model's path): same numbers from ONE launch, plus the scene-graph merge (obbpose_model.py:113-131)
    fe = ops.obb_frontend(rays.origins, rays.directions, box, cu(sc['ext']), want_object_rays=True)
    close(fe['origins_o'], g['origins_o'], what='fused origins_o'); close(fe['dirs_o'], g['dirs_o'], what='fused dirs_o')
    # a hit decided on the fused kernel's own object rays may flip where t_far - t_near is within rounding of 0
    flips = int((fe['hit'].cpu().numpy() != g['hit']).sum())
    assert flips <= 2, f"{flips} intersection flips between the fused front-end and the reference"
    hf = torch.from_numpy(g['hit']).float()
    nhit = hf.sum(-1)
    assert float(nhit.max()) >= 2.0
    if flips == 0:
        close(fe['zi'], g['zi'], what='fused zi'); close(fe['zo'], g['zo'], what='fused zo')
        close(fe['nhit'], nhit.numpy(), rtol=0, scale=1, what='nhit')
        bk = (nhit == 0).float()
        want_os = (T(g['origins_o']) * hf[..., None]).sum(-2) + bk[:, None] * T(sc['rays'].origins)
        want_ds = (T(g['dirs_o']) * hf[..., None]).sum(-2) + bk[:, None] * T(sc['rays'].directions)
        close(fe['origins_s'], want_os.numpy(), what='origins_s (multi-hit rays are SUMMED like the reference)')
        close(fe['dirs_s'], want_ds.numpy(), what='dirs_s')
        close(fe['zo_ret'], (hf * T(g['zo'])).sum(-1).numpy(), what='zo_ret')


# ------------------------------------------------------------------------------------------------ R5-R10: mip.py / mip360.py
def _ipe_tol(means, covd, weighted):
    sc_l = 2.0 ** torch.arange(0, 10, dtype=torch.float64)
    shp = means.shape[:-1]
    ulp = T(np.spacing(np.abs(means.numpy()).astype(np.float32)).astype(np.float64))
    ulp = torch.maximum(ulp, torch.tensor(np.spacing(np.float32(1.57))).double())
    cond = (sc_l[:, None] * ulp[..., None, :]).reshape(*shp, 30)
    damp = torch.exp(-0.5 * (sc_l[:, None] ** 2 * covd.double()[..., None, :]).reshape(*shp, 30))
    t = 1e-5 + 2.0 * torch.cat([cond * damp, cond * damp], -1)
    if weighted:
        t = torch.cat([torch.full((*shp, 3), 1e-5, dtype=torch.float64), t], -1)
    return t


def test_mip_functions_against_reference():
    """sample_along_rays (both modes), cast_rays (cone / cylinder), mip360.new_space, integrated_pos_enc, weighted_ipe
    (alpha = 0 / 3.7 / 10), pos_enc, through the reference-named host functions."""
    from durf_b200 import mip, mip360
    g = load('mip')
    sc = C._scene(B=24, K=1, seed=121, far=200.0)
    assert str(g['inputs_sha256']) == C.scene_digest(sc)
    r = H.cuda_rays(sc)
    t_det, _ = mip.sample_along_rays(None, r.origins, r.directions, r.radii, C.N, r.near, r.far, False, False, 'cone')
    t_rnd, samples = mip.sample_along_rays(cu(sc['t_rand']), r.origins, r.directions, r.radii, C.N, r.near, r.far, True, False, 'cone')
    close(t_det, g['t_det'], what='t_det'); close(t_rnd, g['t_rnd'], what='t_rnd')
    mean, covd = samples.gaussians()
    close(mean, g['cone_mean'], what='cone mean')
    want_covd = np.diagonal(g['cone_cov'], axis1=-2, axis2=-1)
    close(covd, want_covd, rtol=2e-5, scale=1e-3, what='cone cov diag')
    cyl = mip.cast_rays(cu(g['t_rnd']), r.origins, r.directions, r.radii, 'cylinder')
    m2, c2 = cyl.gaussians()
    close(m2, g['cyl_mean'], what='cyl mean')
    close(c2, np.diagonal(g['cyl_cov'], axis1=-2, axis2=-1), rtol=2e-5, scale=1e-3, what='cyl cov diag')
    con = mip360.new_space(mip.cast_rays(cu(g['t_rnd']), r.origins, r.directions, r.radii, 'cone'))
    m3, c3 = con.gaussians()
    close(m3, g['contract_mean'], what='contracted mean')
    close(c3, np.diagonal(g['contract_cov'], axis1=-2, axis2=-1), rtol=2e-5, scale=1e-3, what='contracted cov diag')
    enc = mip.integrated_pos_enc(con, 0, 10).cpu()
    tol = _ipe_tol(T(g['contract_mean']), T(np.ascontiguousarray(np.diagonal(g['contract_cov'], axis1=-2, axis2=-1))), False)
    assert bool(((enc.double() - T(g['ipe_contracted']).double()).abs() <= tol).all()), 'IPE (contracted)'
    plain = mip.cast_rays(cu(g['t_det']), r.origins, r.directions, r.radii, 'cone')
    pm, pc = plain.gaussians()
    enc = mip.integrated_pos_enc(plain, 0, 10).cpu()
    tol = _ipe_tol(pm.cpu(), pc.cpu(), False)
    assert bool(((enc.double() - T(g['ipe_plain']).double()).abs() <= tol).all()), 'IPE (world space, |2^9 x| ~ 1e5)'
    obj = mip.cast_rays(cu(g['t_obj']), r.origins, r.viewdirs, r.radii, 'cone')
    om, oc = obj.gaussians()
    for a in (0.0, 3.7, 10.0):
        enc = mip.weighted_ipe(obj, 0, 10, a).cpu()[:8]
        tol = _ipe_tol(om.cpu()[:8], oc.cpu()[:8], True)
        assert bool(((enc.double() - T(g[f'wipe_alpha_{a}']).double()).abs() <= tol).all()), f'weighted_ipe alpha={a}'
    close(mip.pos_enc(r.viewdirs, 0, 4, True), g['pos_enc'], what='pos_enc')


def test_bf16_encoder_bound_at_object_frame_magnitudes():
    """The fast bf16 tile encoder (3 sincosf + angle doubling, no 100*pi wrap) against the reference-executed weighted_ipe:
    error <= one bf16 ulp of a unit-range value (2^-8) + the conditioning term, at object-frame magnitudes (t up to 12,
    |2^9 x| ~ 5e3) -- stated here, not in a comment."""
    ops = _ops()
    g = load('mip')
    sc = C._scene(B=24, K=1, seed=121, far=200.0)
    r = H.cuda_rays(sc)
    for a in (3.7, 10.0):
        tiles = ops.raymarch(r.origins, r.viewdirs, r.radii, C.N, t_vals=cu(g['t_obj']), weighted=True, alpha=a, bf16_tiles=True)
        dec = H.unswizzle_tiles(tiles['features'], 24 * C.N, 63).reshape(24, C.N, 63)[:8]
        want = T(g[f'wipe_alpha_{a}'])
        err = (dec.double() - want.double()).abs()
        bound = 2.0 ** -8 * torch.clamp(want.abs().double(), min=1.0) * 1.01 + 1e-4
        bound[..., :3] = 2.0 ** -8 * torch.clamp(want[..., :3].abs().double(), min=1.0) * 1.01      # identity part: bf16 rounding only
        assert bool((err <= bound).all()), f"bf16 weighted tiles alpha={a}: worst excess {float((err - bound).max()):.3e}"


def test_volumetric_rendering_and_resampling_against_reference():
    from durf_b200 import mip
    g = load('mip')
    sc = C._scene(B=24, K=1, seed=121, far=200.0)
    r = H.cuda_rays(sc)
    rgb = torch.sigmoid(cu(g['raw_rgb'])); den = torch.nn.functional.softplus(cu(g['raw_den']) - 1.0)
    t = cu(g['t_rnd'])
    for tag, white, rand in (('grey', False, False), ('white', True, False), ('randbg', False, True)):
        out = mip.volumetric_rendering(rgb, den, t, r.directions, white, rand, None)
        close(out[0], g[f'vr_{tag}_comp_rgb'], what=f'{tag} comp_rgb')
        if tag == 'grey':
            for i, nm in enumerate(('comp_rgb', 'depth', 'acc', 'weights', 't_vals', 't_mids', 't_dists')):
                close(out[i], g[f'vr_grey_{nm}'], what=nm)
    w = cu(g['vr_grey_weights'])
    wc = T(g['vr_grey_weights'])
    wp = torch.cat([wc[:, :1], wc, wc[:, -1:]], -1)
    wmax = torch.maximum(wp[:, :-1], wp[:, 1:])
    wblur = 0.5 * (wmax[:, :-1] + wmax[:, 1:]) + 0.01
    for randomized, key in ((False, 'resample_det'), (True, 'resample_rnd')):
        new_t, _ = mip.resample_along_rays(cu(sc['u_rand']) if randomized else None, r.origins, r.directions, r.radii, t, w,
                                           randomized, 'cone', True, 0.01)
        frac = H.assert_samples_close(new_t.cpu(), T(g[key]), T(g['t_rnd']), wblur, what=key)
        assert frac > 0.99


# ------------------------------------------------------------------------------------------------ R11-R17: obbpose_model.py
def cuda_model_case(name, precision):
    from durf_b200.obbpose_model import MipNerfModel
    skw, mover, akw = C.MODEL_CASES[name]
    sc = C._scene(**skw)
    fields = {k: v for k, v in mover.items() if k in MipNerfModel.__dataclass_fields__}
    model = MipNerfModel(precision=precision, **fields)
    v = H.cuda_variables(sc, model)
    rng = dict(t_rand=cu(sc['t_rand']), u_rand=cu(sc['u_rand']),
               density_noise=[cu(sc['noise'][i].reshape(sc['B'], C.N)) for i in range(2)])
    ret = model.apply(v, rng, H.cuda_rays(sc), None, cu(sc['ext']), torch.tensor([akw['ts']]), akw['randomized'],
                      akw.get('rand_bkgd', False), akw.get('white_bkgd', False), akw['alpha'])
    return sc, ret


@pytest.mark.parametrize("name", list(C.MODEL_CASES))
def test_model_fp32_against_reference(name):
    """MipNerfModel.apply (fp32 MLP kernels) vs the reference's MipNerfModel.__call__ on the same rays / weights / draws:
    static C1, contracted C2, dynamic C3, 8 objects with multi-hit rays (C4), BARF alpha (C5), density noise + cylinder +
    white background, and the gain-3 network."""
    g = load('model')
    sc, ret = cuda_model_case(name, 'fp32')
    for lvl, r in enumerate(ret):
        rt = 1e-5 if lvl == 0 else 2e-4          # level 1 sits behind the ill-conditioned inverse CDF of level-0 weights
        if name == 'c7_gain3':
            rt *= 4
        for i, nm in enumerate(('comp_rgb', 'distance', 'acc', 'weights', 't_vals')):
            close(r[i], g[f'{name}/L{lvl}/{nm}'], rtol=rt if nm != 'weights' else 5 * rt, scale=1.0 if nm != 'weights' else 0.2,
                  what=f'{name} L{lvl} {nm}')
        close(r[8], g[f'{name}/L{lvl}/dyn_mask'], rtol=0, what='dyn_mask')
        close(r[9], g[f'{name}/L{lvl}/zo'], what='zo')
    close(ret[-1][7][0], g[f'{name}/off_pose'], rtol=0, scale=1, what='off pose'); close(ret[-1][7][1], g[f'{name}/off_rot'], rtol=0, what='off rot')


# bf16 tolerance, stated: relative Frobenius error of the composited rgb / acc / distance and of the weights, plus PSNR of
# the composited image, per case.  c7_gain3 is the meaningful one (saturating sigmoid / softplus, dense ReLU flips).
BF16_TOL = {name: dict(rgb=1e-2, weights=3e-2, psnr=40.0) for name in C.MODEL_CASES}
BF16_TOL['c7_gain3'] = dict(rgb=3e-2, weights=8e-2, psnr=32.0)


@pytest.mark.parametrize("name", list(C.MODEL_CASES))
def test_model_bf16_against_reference(name):
    g = load('model')
    sc, ret = cuda_model_case(name, 'bf16')
    tol = BF16_TOL[name]
    for lvl, r in enumerate(ret):
        want_rgb = T(g[f'{name}/L{lvl}/comp_rgb'])
        got_rgb = r[0].cpu()
        ok = torch.isfinite(want_rgb).all(-1)                       # multi-hit rays may be non-finite on both sides
        assert bool(torch.isfinite(got_rgb[ok]).all())
        rel = float((got_rgb[ok] - want_rgb[ok]).norm() / want_rgb[ok].norm())
        mse = float(((got_rgb[ok] - want_rgb[ok]) ** 2).mean())
        psnr = -10.0 * np.log10(max(mse, 1e-20))
        assert rel <= tol['rgb'] and psnr >= tol['psnr'], f"{name} L{lvl}: rgb rel Frobenius {rel:.3e}, PSNR {psnr:.1f} dB"
        if lvl == 0:                                                # level 0 shares t_vals exactly: weights comparable 1:1
            ww, gw = T(g[f'{name}/L0/weights'])[ok], r[3].cpu()[ok]
            relw = float((gw - ww).norm() / ww.norm())
            assert relw <= tol['weights'], f"{name}: coarse weights rel Frobenius {relw:.3e}"
            close(r[4], g[f'{name}/L0/t_vals'], what='t_vals')


# ------------------------------------------------------------------------------------------------ R18-R19: train_boxpose.py
def cuda_train_case(name, precision):
    from durf_b200.obbpose_model import MipNerfModel
    from durf_b200.train import TrainState, train_step
    from durf_b200.utils import Config
    skw, mover, cover, st = C.TRAIN_CASES[name]
    sc = C._scene(**skw)
    model = MipNerfModel(precision=precision, **{k: v for k, v in mover.items() if k in MipNerfModel.__dataclass_fields__})
    config = Config(**cover)
    v = H.cuda_variables(sc, model)
    before = v.flat.clone()
    state = TrainState.create(v)
    tg, ts = sc['targets'], st['ts']
    batch = dict(rays=H.cuda_rays(sc), ext=cu(sc['ext']), ts=torch.tensor([ts]), pixels=cu(tg['pixels']), depth=cu(tg['depth']),
                 sky=cu(tg['sky']))
    rng = dict(t_rand=cu(sc['t_rand']), u_rand=cu(sc['u_rand']),
               density_noise=[cu(sc['noise'][i].reshape(sc['B'], C.N)) for i in range(2)])
    prev = cu(sc['centers'][ts + 1 if ts == 0 else ts - 1])[None]
    state, stats = train_step(model, config, rng, state, batch, lr=st['lr'], eps=st['eps'], alpha=st['alpha'], prev=prev)
    torch.cuda.synchronize()
    return sc, v, before, stats


def _named_slices(v, sc):
    """name -> (offset, size) of every reference leaf inside Variables.flat."""
    out = {}
    for net in ['MLP_0'] + [f'BoxMLP_{k}' for k in range(sc['K'])]:
        base = v.slots[net][0]
        off = base
        for i, (w, b) in enumerate(v.layers(net)):
            out[f'{net}/Dense_{i}/kernel'] = (off, w.numel()); off += w.numel()
            out[f'{net}/Dense_{i}/bias'] = (off, b.numel()); off += b.numel()
    out['box_centers'] = v.slots['box_centers']
    return out


@pytest.mark.parametrize("name", list(C.TRAIN_CASES))
def test_train_step_fp32_against_reference(name):
    """train_step (fp32 kernels) vs the reference's train_boxpose.train_step: every loss term, the RAW gradient of every
    parameter tensor (norm + 32 random projections, tests/ref_cases.py), the clipped gradient norm and the Adam update.
    Cases: gin defaults; pose + yaw optimisation with BARF alpha 2.5; box_loss_mult / tv_loss_mult / weight_decay_mult /
    density_noise / coarse_loss_mult all non-default; single object without clipping and without multiscale loss."""
    g = load('train')
    pre = name + '/'
    skw, mover, cover, st = C.TRAIN_CASES[name]
    # raw gradients need clipping off; run twice: (a) clipping off for the gradient, (b) the case's own config for the update
    from durf_b200.utils import Config
    C.TRAIN_CASES['_raw'] = (skw, mover, dict(cover, grad_max_val=0.0, grad_max_norm=0.0), st)
    try:
        sc, v, before, stats = cuda_train_case('_raw', 'fp32')
    finally:
        del C.TRAIN_CASES['_raw']
    want_loss = float(g[pre + 'stats/loss'])
    assert abs(float(stats['loss']) - want_loss) <= 2e-4 * max(1.0, abs(want_loss)), (float(stats['loss']), want_loss)
    for f in ('losses', 'd_losses', 'n_losses', 'e_losses', 's_losses', 'tv_losses'):
        close(stats[f], g[pre + 'stats/' + f], rtol=5e-4, scale=1e-2, what=f)
    close(stats['distr_losses'], g[pre + 'stats/distr_losses'], rtol=2e-3, what='distr_losses')
    close(stats['obj_losses'], g[pre + 'stats/obj_losses'], rtol=5e-4, scale=1e-2, what='obj_losses')
    close(stats['weight_l2'], g[pre + 'stats/weight_l2'], rtol=1e-5, scale=1e-12, what='weight_l2')
    grad = stats['grad'].double().cpu().numpy()
    names = C.param_names(sc['K'])
    slices = _named_slices(v, sc)
    tol = 2e-2         # SURVEY §7 asks cosine >= 0.999 per tensor, i.e. |dg|/|g| <= 4.5e-2; we hold 2e-2 (cosine >= 0.9998);
    #                    the large tensors agree to ~1e-3, the loosest are first-layer biases of object MLPs with |g| ~ 4e-5
    for i, n in enumerate(names):
        off, cnt = slices[n]
        gr = grad[off:off + cnt]
        want_norm = float(g[pre + 'grad_norms'][i])
        if want_norm == 0.0:
            assert np.linalg.norm(gr) == 0.0, f"{n}: expected an exactly zero gradient"
            continue
        R = C.projections(n, cnt).astype(np.float64)
        err = np.sqrt(np.mean((R @ gr - g[pre + 'grad_projs'][i]) ** 2))
        assert err <= tol * want_norm, f"{n}: |dg| ~ {err:.3e} vs |g| {want_norm:.3e}"
        assert abs(np.linalg.norm(gr) / want_norm - 1.0) <= tol, n
    off, cnt = slices['box_centers']
    gb = grad[off:off + cnt].reshape(g[pre + 'grad/box_centers'].shape)
    scale = float(np.abs(g[pre + 'grad/box_centers']).max())
    if scale > 0:
        # pose gradients run through sin(2^l x) up to l = 9: 2e-3 of the largest entry (the fp32 oracle itself sits ~1e-4 off fp64)
        assert float(np.abs(gb - g[pre + 'grad/box_centers']).max()) <= 2e-3 * scale, 'd box_centers'
    # (b) the case's own clipping + Adam
    sc, v, before, stats = cuda_train_case(name, 'fp32')
    close(stats['grad_norm'], g[pre + 'stats/grad_norm'], rtol=2e-3, scale=1e-6, what='grad_norm (after value clip)')
    delta = (v.flat - before).double().cpu().numpy()
    for i, n in enumerate(names):
        off, cnt = slices[n]
        d = delta[off:off + cnt]
        want_norm = float(g[pre + 'update_norms'][i])
        if want_norm == 0.0:
            assert np.linalg.norm(d) == 0.0, n
            continue
        R = C.projections(n, cnt).astype(np.float64)
        err = np.sqrt(np.mean((R @ d - g[pre + 'update_projs'][i]) ** 2))
        # first Adam step = lr * g / (|g| + 1e-8): sign-like, every entry whose gradient is smaller than the float32 noise of
        # the two implementations may flip -> 5 % of the update norm (15 % for the eps = 0.5 case, see GRAD_TOL)
        upd_tol = 0.15 if name == 't_noclip_single' else 5e-2
        assert err <= upd_tol * want_norm, f"{n}: Adam update differs by ~{err:.3e} of {want_norm:.3e}"
    newb = v.box_centers.cpu().numpy()
    close(newb, g[pre + 'new/box_centers'], rtol=2e-5, what='box_centers after Adam')


@pytest.mark.parametrize("name", ['t_default', 't_pose', 't_extras'])
def test_train_step_bf16_against_reference(name):
    """The tensor-core train step vs the reference: loss within 1e-2, per-tensor gradient cosine >= 0.97 estimated from the
    projections (bf16 forward flips ReLU masks near zero; see test_mlp_tensor_core_backward for the split of that error)."""
    g = load('train')
    pre = name + '/'
    skw, mover, cover, st = C.TRAIN_CASES[name]
    C.TRAIN_CASES['_raw'] = (skw, mover, dict(cover, grad_max_val=0.0, grad_max_norm=0.0), st)
    try:
        sc, v, before, stats = cuda_train_case('_raw', 'bf16')
    finally:
        del C.TRAIN_CASES['_raw']
    want_loss = float(g[pre + 'stats/loss'])
    assert abs(float(stats['loss']) - want_loss) <= 1e-2 * max(1.0, abs(want_loss)), (float(stats['loss']), want_loss)
    grad = stats['grad'].double().cpu().numpy()
    assert np.isfinite(grad).all()
    names = C.param_names(sc['K'])
    slices = _named_slices(v, sc)
    for i, n in enumerate(names):
        if n == 'box_centers':
            continue
        off, cnt = slices[n]
        gr = grad[off:off + cnt]
        want_norm = float(g[pre + 'grad_norms'][i])
        if want_norm == 0.0:
            assert np.linalg.norm(gr) == 0.0, n
            continue
        R = C.projections(n, cnt).astype(np.float64)
        err = np.sqrt(np.mean((R @ gr - g[pre + 'grad_projs'][i]) ** 2))
        # |dg|/|g| = sqrt(2 - 2 cos) for equal norms: cosine 0.97 <-> 0.245
        assert err <= 0.25 * want_norm, f"{n}: |dg| ~ {err:.3e} vs |g| {want_norm:.3e}"
        assert abs(np.linalg.norm(gr) / want_norm - 1.0) <= 0.1, n
