"""Parity against fixtures produced by EXECUTING THE REFERENCE'S OWN CODE (tests/golden/ref_*.npz; generator:
tests/golden/make_ref_golden.py; how the reference runs without JAX: tests/_refshim/README.md).

CPU half (this file, not `gpu`): the oracle (oracle/durf_oracle.py) must reproduce the reference-executed values.
Tolerances are a few float32 ulps: both sides are float32 on the CPU and differ only in summation order / libm.
The GPU half is tests/test_ref_golden_gpu.py (CUDA path vs the same fixtures).
"""
import os

import numpy as np
import pytest
import torch

import durf_test_helpers as H
import ref_cases as C
from oracle import durf_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return np.load(os.path.join(GOLD, f'ref_{name}.npz'))


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def close(got, want, rtol=2e-6, atol=2e-6, what=''):
    got = np.asarray(got.detach() if hasattr(got, 'detach') else got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), f"{what}: NaN pattern differs ({nan_g.sum()} vs {nan_w.sum()})"
    err = np.abs(np.where(nan_w, 0, got - want))
    tol = atol + rtol * np.abs(np.where(nan_w, 0, want))
    assert (err <= tol).all(), f"{what}: {(err > tol).sum()}/{err.size} beyond tol, worst {np.max(err - tol):.3e} (|want| max {np.nanmax(np.abs(want)):.3e})"


def cov_atol(cov):
    """Off-diagonal covariance entries are differences of nearly equal products: absolute tolerance 3e-6 of the
    sample's largest entry."""
    return 3e-6 * np.abs(cov).max(axis=(-1, -2), keepdims=True)


# ------------------------------------------------------------------------------------------------ internal/math.py
def test_fixture_inputs_are_reproducible():
    """The seeded inputs the fixtures were generated from are the ones this checkout produces."""
    assert str(load('math')['inputs_sha256']) == C.digest(C.math_inputs().values())
    g = load('model')
    for name, (skw, _, _) in C.MODEL_CASES.items():
        assert str(g[f'{name}/inputs_sha256']) == C.scene_digest(C._scene(**skw)), name
    g = load('train')
    for name, (skw, _, _, _) in C.TRAIN_CASES.items():
        assert str(g[f'{name}/inputs_sha256']) == C.scene_digest(C._scene(**skw)), name


def test_math_safe_trig_and_schedules():
    g, I = load('math'), C.math_inputs()
    x = T(I['trig_x'])
    # sin/cos of the SAME float32 argument after the reference's own `x mod 100 pi` reduction: libm-level agreement
    close(O.safe_sin(x), g['safe_sin'], rtol=0, atol=3e-7, what='safe_sin')
    close(O.safe_cos(x), g['safe_cos'], rtol=0, atol=3e-7, what='safe_cos')
    cfg = dict(lr_init=5e-4, lr_final=5e-6, max_steps=200000, lr_delay_steps=2500, lr_delay_mult=0.01)
    close(np.array([float(O.learning_rate_decay(int(s), **cfg)) for s in I['steps']]), g['lr'], rtol=2e-6, atol=0, what='lr')
    ecfg = dict(lr_init=3.0, lr_final=0.2, max_steps=200000, lr_delay_steps=0, lr_delay_mult=0.01)
    close(np.array([float(O.learning_rate_decay(int(s), **ecfg)) for s in I['steps']]), g['eps'], rtol=2e-6, atol=0, what='eps')
    close(np.array([float(O.freq_alpha_rate(int(s), 0.0, 10.0, 1000, 100000)) for s in I['steps']]), g['alpha'], rtol=1e-7, what='alpha')
    close(O.mse_to_psnr(T(np.array([1e-4, 0.07, 0.5], np.float32))), g['mse_to_psnr'], what='psnr')
    close(O.psnr_to_mse(T(g['mse_to_psnr'])), g['psnr_to_mse'], what='mse')


def test_math_sampler():
    g, I = load('math'), C.math_inputs()
    bins, w, u = T(I['bins']), T(I['weights']), T(I['u'])
    det = O.sorted_piecewise_constant_pdf(bins, w, C.N + 1, False)
    rnd = O.sorted_piecewise_constant_pdf(bins, w, C.N + 1, True, u_rand=u)
    for got, want, tag in ((det, g['pdf_det'], 'det'), (rnd, g['pdf_rand'], 'rand')):
        frac = H.assert_samples_close(got, T(want), bins, w, what=f'sampler {tag}', pos_rtol=2e-6, cdf_ulps=8)
        assert frac > 0.999
    sb = torch.tensor([[0, 1, 3, 6, 10]], dtype=torch.float32)
    for i in range(4):
        sw = torch.zeros(1, 4); sw[0, i] = 1.0
        close(O.sorted_piecewise_constant_pdf(sb, sw, 625, False), g[f'pdf_single_bin_{i}'], rtol=2e-6, atol=2e-6, what=f'single bin {i}')


def test_eval_metrics_against_reference():
    """durf_b200.math (the product's host-side metrics) vs the reference-executed compute_ssim / sRGB."""
    from durf_b200 import math as dmath
    g, I = load('math'), C.math_inputs()
    close(dmath.compute_ssim(T(I['img0']), T(I['img1']), 1.0), g['ssim'], rtol=1e-5, atol=1e-6, what='ssim')
    close(dmath.compute_ssim(T(I['img0']), T(I['img1']), 1.0, return_map=True), g['ssim_map'], rtol=1e-4, atol=1e-5, what='ssim map')
    close(dmath.linear_to_srgb(T(I['lin'])), g['linear_to_srgb'], rtol=1e-6, atol=1e-6, what='linear_to_srgb')
    close(dmath.srgb_to_linear(T(I['lin'])), g['srgb_to_linear'], rtol=1e-6, atol=1e-6, what='srgb_to_linear')
    close(dmath.compute_avg_error(27.5, 0.83, 0.21), g['avg_error'], rtol=1e-5, what='avg_error')


# ------------------------------------------------------------------------------------------------ internal/box_helpers.py
def test_obb_frontend_functions():
    g = load('obb')
    sc = C._scene(B=512, K=4, seed=111, overlap=True)
    assert str(g['inputs_sha256']) == C.scene_digest(sc)
    box = T(g['box'])
    B, K = 512, 4
    R = O.aa2matrix(box[:, 3:])
    close(R, g['aa2matrix'], what='aa2matrix')
    rays = O.Rays(*[T(a) for a in sc['rays']])
    oo, do = O.world2object_rpy(rays.origins, rays.directions, box[:, :3].expand(B, K, 3), R.expand(B, K, 3, 3))
    close(oo, g['origins_o'], rtol=3e-6, atol=3e-6, what='origins_o'); close(do, g['dirs_o'], what='dirs_o')
    ext = T(sc['ext'])
    zi, zo, hit = O.ray_box_intersection(T(g['origins_o']), T(g['dirs_o']), -ext.expand(B, K, 3), ext.expand(B, K, 3))
    assert np.array_equal(hit.numpy().astype(np.int32), g['hit']), "intersection mask must be bit-exact"
    close(zi, g['zi'], what='zi'); close(zo, g['zo'], what='zo')
    assert (g['hit'].sum(-1) >= 2).sum() > 10


# ------------------------------------------------------------------------------------------------ internal/mip.py, mip360.py
def test_mip_sampling_gaussians_contraction_encodings():
    g = load('mip')
    sc = C._scene(B=24, K=1, seed=121, far=200.0)
    assert str(g['inputs_sha256']) == C.scene_digest(sc)
    r = O.Rays(*[T(a) for a in sc['rays']])
    t_det, (mean_d, cov_d) = O.sample_along_rays(r.origins, r.directions, r.radii, C.N, r.near, r.far, False)
    t_rnd, (mean_r, cov_r) = O.sample_along_rays(r.origins, r.directions, r.radii, C.N, r.near, r.far, True, t_rand=T(sc['t_rand']))
    close(t_det, g['t_det'], what='t_det'); close(t_rnd, g['t_rnd'], what='t_rnd')
    close(mean_r, g['cone_mean'], rtol=3e-6, atol=3e-6, what='cone mean')
    close(cov_r, g['cone_cov'], rtol=2e-5, atol=cov_atol(g['cone_cov']), what='cone cov')
    mc, cc = O.cast_rays(T(g['t_rnd']), r.origins, r.directions, r.radii, 'cylinder')
    close(mc, g['cyl_mean'], rtol=3e-6, atol=3e-6, what='cyl mean'); close(cc, g['cyl_cov'], rtol=2e-5, atol=cov_atol(g['cyl_cov']), what='cyl cov')
    cm, ccov = O.new_space((T(g['cone_mean']), T(g['cone_cov'])))
    close(cm, g['contract_mean'], what='contracted mean')
    close(ccov, g['contract_cov'], rtol=2e-5, atol=cov_atol(g['contract_cov']), what='contracted cov')
    # encodings from the reference's own gaussians: only the sin/exp evaluation differs
    close(O.integrated_pos_enc((T(g['contract_mean']), T(g['contract_cov'])), 0, 10), g['ipe_contracted'], rtol=0, atol=2e-6, what='IPE contracted')
    mean_d, cov_d = O.cast_rays(T(g['t_det']), r.origins, r.directions, r.radii)
    want = g['ipe_plain']
    got = O.integrated_pos_enc((mean_d, cov_d), 0, 10)
    # un-contracted world-space means reach |2^9 x| ~ 1e5: a 1-ulp difference of x moves sin by 2^l ulp(x) (damped by the variance)
    covd = torch.diagonal(cov_d, dim1=-2, dim2=-1)
    sc_l = 2.0 ** torch.arange(0, 10, dtype=torch.float64)
    ulp = T(np.spacing(np.abs(mean_d.numpy())).astype(np.float64))
    cond = (sc_l[:, None] * ulp[..., None, :]).reshape(24, C.N, 30)
    damp = torch.exp(-0.5 * (sc_l[:, None] ** 2 * covd.double()[..., None, :]).reshape(24, C.N, 30))
    tol = 2e-6 + 2.0 * torch.cat([cond * damp, cond * damp], -1)
    assert bool(((got.double() - T(want).double()).abs() <= tol).all()), 'IPE plain'
    t_obj = T(g['t_obj'])
    mo, co = O.cast_rays(t_obj, r.origins, r.viewdirs, r.radii)
    for a in (0.0, 3.7, 10.0):
        got = O.weighted_ipe((mo[:8], co[:8]), 0, 10, a)
        want = T(g[f'wipe_alpha_{a}'])
        ulp = T(np.spacing(np.abs(mo[:8].numpy())).astype(np.float64))
        cond = (sc_l[:, None] * ulp[..., None, :]).reshape(8, C.N, 30)
        tol = torch.cat([torch.full((8, C.N, 3), 3e-6, dtype=torch.float64), 2e-6 + 2.0 * torch.cat([cond, cond], -1)], -1)
        assert bool(((got.double() - want.double()).abs() <= tol).all()), f'weighted_ipe alpha={a}'
        if a == 0.0:
            assert float(want[..., 3:].abs().max()) == 0.0          # BARF: every frequency weight is 0 at alpha = 0
    close(O.pos_enc(r.viewdirs, 0, 4, True), g['pos_enc'], rtol=0, atol=3e-7, what='pos_enc')


def test_mip_volumetric_rendering_and_resampling():
    g = load('mip')
    sc = C._scene(B=24, K=1, seed=121, far=200.0)
    r = O.Rays(*[T(a) for a in sc['rays']])
    rgb = torch.sigmoid(T(g['raw_rgb'])); den = torch.nn.functional.softplus(T(g['raw_den']) - 1.0)
    t = T(g['t_rnd'])
    for tag, white, rand in (('grey', False, False), ('white', True, False), ('randbg', False, True)):
        out = O.volumetric_rendering(rgb, den, t, r.directions, white, rand)
        close(out[0], g[f'vr_{tag}_comp_rgb'], what=f'{tag} comp_rgb')
        if tag == 'grey':
            for i, nm in enumerate(('comp_rgb', 'depth', 'acc', 'weights', 't_vals', 't_mids', 't_dists')):
                close(out[i], g[f'vr_grey_{nm}'], rtol=3e-6, atol=3e-6, what=nm)
    w = T(g['vr_grey_weights'])
    det = O.resample_t_vals(t, w, False, 0.01)
    rnd = O.resample_t_vals(t, w, True, 0.01, u_rand=T(sc['u_rand']))
    wp = torch.cat([w[:, :1], w, w[:, -1:]], -1)
    wmax = torch.maximum(wp[:, :-1], wp[:, 1:])
    wblur = 0.5 * (wmax[:, :-1] + wmax[:, 1:]) + 0.01
    for got, want, tag in ((det, g['resample_det'], 'det'), (rnd, g['resample_rnd'], 'rand')):
        frac = H.assert_samples_close(got, T(want), t, wblur, what=f'resample {tag}', pos_rtol=2e-6, cdf_ulps=8)
        assert frac > 0.99


# ------------------------------------------------------------------------------------------------ obbpose_model.py
def oracle_model_case(name, dtype=torch.float32, params=None):
    skw, mover, akw = C.MODEL_CASES[name]
    sc = C._scene(**skw)
    fields = {k: v for k, v in mover.items() if k in O.ModelConfig._fields}
    cfg = O.ModelConfig(**fields)
    params = params or H.oracle_params(sc, dtype)
    noise = [T(sc['noise'][0]).to(dtype), T(sc['noise'][1]).to(dtype)]
    ret = O.model_forward(params, H.oracle_rays(sc, dtype), T(sc['ext']).to(dtype), akw['ts'], akw['randomized'],
                          akw.get('rand_bkgd', False), akw.get('white_bkgd', False), akw['alpha'], cfg=cfg,
                          t_rand=T(sc['t_rand']).to(dtype), u_rand=T(sc['u_rand']).to(dtype), density_noise=noise)
    return sc, cfg, ret


@pytest.mark.parametrize("name", list(C.MODEL_CASES))
def test_model_forward_against_reference(name):
    """MipNerfModel.__call__ (obbpose_model.py:69-261) executed by the reference vs oracle.model_forward, both levels."""
    g = load('model')
    sc, cfg, ret = oracle_model_case(name)
    for lvl, r in enumerate(ret):
        # level 1 sits behind the inverse CDF of level-0 weights (ill-conditioned in position where a bin holds little mass)
        rt, at = (3e-6, 3e-6) if lvl == 0 else (2e-4, 2e-5)
        if name == 'c7_gain3':
            rt, at = rt * 4, at * 4            # gain-3 heads amplify the fp32 summation-order differences of the GEMMs
        for nm, v in zip(('comp_rgb', 'distance', 'acc', 'weights', 't_vals'), r[:5]):
            close(v, g[f'{name}/L{lvl}/{nm}'], rtol=rt, atol=at * max(1.0, float(np.nanmax(np.abs(g[f"{name}/L{lvl}/{nm}"]))) if nm in ('distance', 't_vals') else 1.0),
                  what=f'{name} L{lvl} {nm}')
        close(r.dyn_mask, g[f'{name}/L{lvl}/dyn_mask'], rtol=0, atol=0, what='dyn_mask')
        close(r.zo, g[f'{name}/L{lvl}/zo'], what='zo')
    if name == 'c4_k8_overlap':
        assert float(g[f'{name}/L0/dyn_mask'].max()) >= 2.0, "the overlap case must contain rays that cross two boxes"


def test_param_tree_matches_reference_init():
    """Names and shapes of the parameter tree the reference's construct_mipnerf creates == the checkpoint layout we write."""
    from durf_b200 import synthetic as S
    want = sorted(str(s) for s in load('model')['param_tree'])
    got = []
    for net, (fin, width) in (('MLP_0', (60, 256)), ('BoxMLP_0', (63, 128)), ('BoxMLP_1', (63, 128))):
        for i, (a, b) in enumerate(S.layer_shapes(fin, width)):
            got += [f'{net}/Dense_{i}/kernel:{a}x{b}', f'{net}/Dense_{i}/bias:{b}']
    got.append('box_centers:5x2x6')
    assert sorted(got) == want


# ------------------------------------------------------------------------------------------------ train_boxpose.py
def oracle_train_case(name):
    skw, mover, cover, st = C.TRAIN_CASES[name]
    sc = C._scene(**skw)
    cfg = O.ModelConfig(**{k: v for k, v in mover.items() if k in O.ModelConfig._fields})
    lcfg = O.LossConfig(**{k: float(v) if isinstance(v, int) and not isinstance(v, bool) else v for k, v in cover.items()})
    params = H.oracle_params(sc)
    named = {}
    for i, (k, b) in enumerate(params['mlp']):
        named[f'MLP_0/Dense_{i}/kernel'], named[f'MLP_0/Dense_{i}/bias'] = k, b
    for j, m in enumerate(params['box_mlps']):
        for i, (k, b) in enumerate(m):
            named[f'BoxMLP_{j}/Dense_{i}/kernel'], named[f'BoxMLP_{j}/Dense_{i}/bias'] = k, b
    named['box_centers'] = params['box_centers']
    for t in named.values():
        t.requires_grad_(True)
    ts = st['ts']
    noise = [T(sc['noise'][0]), T(sc['noise'][1])]
    ret = O.model_forward(params, H.oracle_rays(sc), T(sc['ext']), ts, True, False, False, st['alpha'], cfg=cfg,
                          t_rand=T(sc['t_rand']), u_rand=T(sc['u_rand']), density_noise=noise)
    tg = {k: T(v) for k, v in sc['targets'].items()}
    prev = T(sc['centers'][ts + 1 if ts == 0 else ts - 1])[None]
    loss, stats = O.loss_fn(ret, H.oracle_rays(sc), tg['pixels'], tg['depth'], tg['sky'], eps=st['eps'], cfg=lcfg, prev=prev,
                            param_tensors=list(named.values()))
    names = C.param_names(sc['K'])
    grads = torch.autograd.grad(loss, [named[n] for n in names], allow_unused=True)
    grads = [torch.zeros_like(named[n]) if gr is None else gr for n, gr in zip(names, grads)]
    return sc, lcfg, st, names, named, grads, loss, stats


# |g_oracle - g_reference| / |g_reference| per tensor.  Two float32 evaluations of the same step differ where the fine
# level's inverse CDF and ReLU masks amplify rounding: measured against a float64 evaluation, BOTH float32 results of
# t_noclip_single (eps = 0.5: a very peaked line-of-sight loss) sit 1.4e-3 .. 3.7e-3 away from it while agreeing with each
# other to 3.6e-4, so that case gets 1e-3; the others agree to 5e-6 .. 4e-5.
GRAD_TOL = {'t_default': 3e-5, 't_pose': 3e-5, 't_extras': 1e-4, 't_noclip_single': 1e-3}


@pytest.mark.parametrize("name", list(C.TRAIN_CASES))
def test_train_step_against_reference(name):
    """train_boxpose.train_step executed by the reference (loss block :94-220, value_and_grad :251, nan_to_num / clip /
    global-norm clip :262-286, flax Adam :288) vs the oracle's loss_fn + autograd + postprocess_grads + adam_step."""
    g = load('train')
    sc, lcfg, st, names, named, grads, loss, stats = oracle_train_case(name)
    pre = name + '/'
    close(loss, g[pre + 'stats/loss'], rtol=3e-6, atol=0, what='loss')
    for f in ('losses', 'd_losses', 'n_losses', 'e_losses', 's_losses', 'distr_losses', 'tv_losses', 'weight_l2'):
        close(stats[f], g[pre + 'stats/' + f], rtol=2e-5, atol=1e-9, what=f)
    want_obj = g[pre + 'stats/obj_losses']
    close(stats['obj_losses'], want_obj, rtol=2e-5, atol=1e-9, what='obj_losses')
    # raw gradients: norm and 32 random projections per tensor
    for i, n in enumerate(names):
        gr = grads[i].double().reshape(-1).numpy()
        want_norm = float(g[pre + 'grad_norms'][i])
        if want_norm == 0.0:
            assert np.linalg.norm(gr) == 0.0, f"{n}: expected an exactly zero gradient"
            continue
        R = C.projections(n, gr.size).astype(np.float64)
        err = np.sqrt(np.mean((R @ gr - g[pre + 'grad_projs'][i]) ** 2))          # estimates |g - g_ref|
        assert err <= GRAD_TOL[name] * want_norm, f"{n}: |dg| ~ {err:.3e} vs |g| {want_norm:.3e}"
        assert abs(np.linalg.norm(gr) / want_norm - 1.0) <= GRAD_TOL[name], n
    close(grads[-1], g[pre + 'grad/box_centers'], rtol=1e-4, atol=1e-6 * max(1e-30, float(np.abs(g[pre + 'grad/box_centers']).max())), what='d box_centers')
    # post-processing + Adam
    gs, norm = O.postprocess_grads(grads, lcfg)
    close(norm, g[pre + 'stats/grad_norm'], rtol=1e-5, atol=0, what='grad_norm')
    p_list = [named[n].detach() for n in names]
    new_p, _, _ = O.adam_step(p_list, gs, [torch.zeros_like(p) for p in p_list], [torch.zeros_like(p) for p in p_list], step=0, lr=st['lr'])
    for i, n in enumerate(names):
        d = (new_p[i].double() - p_list[i].double()).reshape(-1).numpy()
        want_norm = float(g[pre + 'update_norms'][i])
        if want_norm == 0.0:
            assert np.linalg.norm(d) == 0.0, n
            continue
        # the first Adam step is lr * g / (|g| + 1e-8): ill-conditioned where |g| ~ 1e-8, so compare in norm (2e-3) and on
        # the projections (which weigh every entry alike)
        R = C.projections(n, d.size).astype(np.float64)
        err = np.sqrt(np.mean((R @ d - g[pre + 'update_projs'][i]) ** 2))
        assert err <= 2e-2 * want_norm, f"{n}: update differs by ~{err:.3e} of {want_norm:.3e}"
    close(new_p[-1], g[pre + 'new/box_centers'], rtol=1e-6, atol=1e-6, what='box_centers after Adam')
