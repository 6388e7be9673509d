"""Golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py from the CPU oracle).

CPU leg: the oracle still reproduces its committed vectors (a silent change of the restatement fails here).
GPU leg: the CUDA path, called through the C ABI, reproduces the same vectors without running the oracle.
Tolerances: fp32 stages 1e-5 * max(|x|, 1) (IPE features with the conditioning allowance of sin(2^l x)); the whole
model through fp32 GEMMs 2e-4."""
import os

import numpy as np
import pytest
import torch

from oracle import durf_oracle as O
import durf_test_helpers as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return {k: v for k, v in np.load(os.path.join(GOLD, name)).items()}


def _t(a):
    return torch.from_numpy(np.asarray(a))


# ---- CPU: the oracle is pinned by its own committed outputs -------------------------------------------------------
def test_oracle_reproduces_raymarch_golden():
    g = _load("raymarch.npz")
    t_vals = O.sample_t_vals(_t(g['near']), _t(g['far']), 128, True, t_rand=_t(g['t_rand']))
    mean, cov = O.cast_rays(t_vals, _t(g['origins']), _t(g['directions']), _t(g['radii']), 'cone')
    cm, cc = O.new_space((mean, cov))
    H.assert_close(t_vals, _t(g['t_vals']), rtol=1e-6, what="t_vals")
    H.assert_close(mean, _t(g['means']), rtol=1e-6, what="means")
    H.assert_close(cm, _t(g['contracted_means']), rtol=1e-6, what="contracted means")
    H.assert_close(O.integrated_pos_enc((cm, cc), 0, 10), _t(g['ipe']), rtol=1e-6, atol_scale=1.0, what="ipe")
    H.assert_close(O.weighted_ipe((mean, cov), 0, 10, float(g['alpha'])), _t(g['weighted_ipe']), rtol=1e-6, atol_scale=1.0, what="wipe")
    # fp32 vs fp64 of the same oracle: the Gaussians agree to fp32 rounding
    H.assert_close(mean, _t(g['means_f64']), rtol=2e-6, atol_scale=1.0, what="means vs fp64")
    H.assert_close(torch.diagonal(cov, dim1=-2, dim2=-1), _t(g["cov_diag_f64"]), rtol=1e-4, atol_scale=1e-3, what="cov vs fp64")  # t_var has a cancellation (mip.py:120-122)


def test_oracle_reproduces_composite_and_resample_golden():
    g = _load("composite_resample.npz")
    rgb = torch.sigmoid(_t(g['raw_rgb']))
    den = torch.nn.functional.softplus(_t(g['raw_density']) - 1.0)
    comp = O.volumetric_rendering(rgb, den[..., None], _t(g['t_vals']), _t(g['dirs']), False, False)
    for got, name in zip((comp[0], comp[1], comp[2], comp[3]), ('comp_rgb', 'distance', 'acc', 'weights')):
        H.assert_close(got, _t(g[name]), rtol=1e-6, atol_scale=1.0, what=name)
    new_t = O.resample_t_vals(_t(g['t_vals']), _t(g['weights']), True, 0.01, u_rand=_t(g['u_rand']))
    H.assert_close(new_t, _t(g['resampled_randomized']), rtol=1e-6, what="resampled")


def test_oracle_reproduces_model_golden():
    g = _load("model_train.npz")
    sc = H.scene(B=int(g['B']), K=int(g['K']), seed=int(g['seed']))
    cfg = O.ModelConfig(no_pose_opt=False, no_yaw_opt=False)
    params = H.oracle_params(sc)
    params['box_centers'].requires_grad_(True)
    ret = O.model_forward(params, H.oracle_rays(sc), _t(sc['ext']), int(g['ts']), True, False, False, float(g['alpha']), cfg=cfg,
                          t_rand=_t(sc['t_rand']), u_rand=_t(sc['u_rand']))
    tg = {k: _t(v) for k, v in sc['targets'].items()}
    loss, _ = O.loss_fn(ret, H.oracle_rays(sc), tg['pixels'], tg['depth'], tg['sky'], eps=float(g['eps']))
    H.assert_close(ret[-1].comp_rgb, _t(g['l1_comp_rgb']), rtol=2e-5, atol_scale=1.0, what="fine rgb")
    assert abs(float(loss) - float(g['loss'])) <= 1e-5 * max(1.0, abs(float(g['loss'])))
    assert abs(float(g['loss']) - float(g['loss_f64'])) <= 1e-4 * max(1.0, abs(float(g['loss_f64']))), "fp32 vs fp64 oracle"
    gbox = torch.autograd.grad(loss, params['box_centers'])[0]
    scale = float(np.abs(g['d_box_centers_f64']).max())
    assert float((gbox - _t(g['d_box_centers'])).abs().max()) <= 1e-3 * scale


# ---- GPU: the CUDA path reproduces the golden vectors through the C ABI -------------------------------------------
@pytest.mark.gpu
def test_gpu_raymarch_matches_golden():
    from durf_b200 import ops
    g = _load("raymarch.npz")
    c = lambda k: _t(g[k]).cuda()
    out = ops.raymarch(c('origins'), c('directions'), c('radii'), 128, near=c('near'), far=c('far'), t_rand=c('t_rand'),
                       contract=True, want_gaussians=True)
    H.assert_close(out['t_vals'], _t(g['t_vals']), what="t_vals")
    H.assert_close(out['means'], _t(g['contracted_means']), what="contracted means")
    H.assert_close(out['cov_diag'], _t(g['contracted_cov_diag']), rtol=2e-5, atol_scale=1e-3, what="contracted cov")
    # sin(2^l x) is conditioned like 2^l * ulp(x): allow 2^9 ulp of the largest |mean| on top of 1e-5
    allow = 1e-5 + 512 * 1.2e-7 * float(np.abs(g['contracted_means']).max())
    assert float((out['features'].cpu() - _t(g['ipe'])).abs().max()) <= allow
    plain = ops.raymarch(c('origins'), c('directions'), c('radii'), 128, t_vals=c('t_vals'), weighted=True, alpha=float(g['alpha']))
    x = float(np.abs(g['means']).max())
    assert float((plain['features'].cpu() - _t(g['weighted_ipe'])).abs().max()) <= 1e-5 * max(x, 1.0) + 512 * 1.2e-7 * x


@pytest.mark.gpu
def test_gpu_composite_resample_obb_match_golden():
    from durf_b200 import ops
    g = _load("composite_resample.npz")
    c = lambda k: _t(g[k]).cuda()
    comp = ops.composite(c('raw_rgb'), c('raw_density'), c('t_vals'), c('dirs'))
    for name, key in (('comp_rgb', 'comp_rgb'), ('depth', 'distance'), ('acc', 'acc'), ('weights', 'weights'), ('t_mids', 't_mids'),
                      ('t_dists', 't_dists')):
        H.assert_close(comp[name], _t(g[key]), rtol=2e-5, atol_scale=1.0, what=name)
    wp = np.concatenate([g['weights'][:, :1], g['weights'], g['weights'][:, -1:]], -1)
    wmax = np.maximum(wp[:, :-1], wp[:, 1:])
    wblur = 0.5 * (wmax[:, :-1] + wmax[:, 1:]) + 0.01
    for key, u in (('resampled', None), ('resampled_randomized', c('u_rand'))):
        got = ops.resample(c('t_vals'), c('weights'), u_rand=u).cpu()
        H.assert_samples_close(got, _t(g[key]), _t(g['t_vals']), _t(wblur), what=key)
    o = _load("obb.npz")
    fe = ops.obb_frontend(_t(o['origins']).cuda(), _t(o['directions']).cuda(), _t(o['box']).cuda(), _t(o['ext']).cuda(),
                          want_object_rays=True)
    assert torch.equal(fe['hit'].cpu(), _t(o['hit']))
    H.assert_close(fe['origins_o'], _t(o['origins_o']), rtol=1e-5, atol_scale=1.0, what="origins_o")
    H.assert_close(fe['dirs_o'], _t(o['dirs_o']), rtol=1e-5, atol_scale=1.0, what="dirs_o")
    hit = _t(o['hit']).bool()
    H.assert_close(fe['zo'].cpu()[hit], _t(o['zo'])[hit], rtol=2e-5, atol_scale=1.0, what="zo")


@pytest.mark.gpu
def test_gpu_model_matches_golden():
    from durf_b200.obbpose_model import MipNerfModel
    g = _load("model_train.npz")
    sc = H.scene(B=int(g['B']), K=int(g['K']), seed=int(g['seed']))
    model = MipNerfModel(precision='fp32', no_pose_opt=False, no_yaw_opt=False)
    v = H.cuda_variables(sc, model)
    rng = dict(t_rand=_t(sc['t_rand']).cuda(), u_rand=_t(sc['u_rand']).cuda())
    ret = model.apply(v, rng, H.cuda_rays(sc), None, _t(sc['ext']).cuda(), torch.tensor([int(g['ts'])]), True, False, False,
                      float(g['alpha']))
    for lvl in range(2):
        H.assert_close(ret[lvl][0], _t(g[f'l{lvl}_comp_rgb']), rtol=2e-4, atol_scale=1.0, what=f"level {lvl} rgb")
        H.assert_close(ret[lvl][2], _t(g[f'l{lvl}_acc']), rtol=2e-4, atol_scale=1.0, what=f"level {lvl} acc")
