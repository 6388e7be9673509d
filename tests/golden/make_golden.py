#!/usr/bin/env python
"""Generates tests/golden/*.npz: seeded inputs and outputs of the CPU oracle (oracle/durf_oracle.py) for every stage of
the hot path.  The reference itself (JAX) cannot be run in this environment, so these vectors pin the ORACLE (a change of
the restatement is caught by tests/test_golden.py) and give the GPU parity tests fixed data that needs no oracle run.

    python tests/golden/make_golden.py          # rewrites the fixtures (commit the result)

Every array is float32 unless stated; fp64 companions (suffix _f64) come from the same oracle run in double precision.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p_ in (ROOT, os.path.join(ROOT, "tests")):
    if p_ not in sys.path:
        sys.path.insert(0, p_)

from oracle import durf_oracle as O            # noqa: E402
import durf_test_helpers as H                  # noqa: E402

SEED = 20200823


def np32(t):
    return t.detach().to(torch.float32).cpu().numpy()


def raymarch_golden():
    """sample_along_rays -> cast_rays -> new_space -> integrated_pos_enc / weighted_ipe on 16 rays x 128 samples."""
    sc = H.scene(B=16, K=2, seed=SEED % 1000)
    out = {}
    for dt, suf in ((torch.float32, ""), (torch.float64, "_f64")):
        rays = H.oracle_rays(sc, dt)
        t_rand = torch.from_numpy(sc['t_rand']).to(dt)
        t_vals = O.sample_t_vals(rays.near, rays.far, 128, True, t_rand=t_rand)
        mean, cov = O.cast_rays(t_vals, rays.origins, rays.directions, rays.radii, 'cone')
        cm, cc = O.new_space((mean, cov))
        ipe = O.integrated_pos_enc((cm, cc), 0, 10)
        wipe = O.weighted_ipe((mean, cov), 0, 10, 4.5)
        out.update({f"t_vals{suf}": t_vals.numpy(), f"means{suf}": mean.numpy(),
                    f"cov_diag{suf}": torch.diagonal(cov, dim1=-2, dim2=-1).numpy(),
                    f"contracted_means{suf}": cm.numpy(), f"contracted_cov_diag{suf}": torch.diagonal(cc, dim1=-2, dim2=-1).numpy(),
                    })
        if dt == torch.float32:      # the encodings only in fp32 (size); their fp64 conditioning is covered by test_oracle_analytic
            out.update(ipe=ipe.numpy(), weighted_ipe=wipe.numpy())
    r = sc['rays']
    out.update(origins=r.origins, directions=r.directions, radii=r.radii, near=r.near, far=r.far, t_rand=sc['t_rand'],
               alpha=np.float32(4.5))
    np.savez_compressed(os.path.join(HERE, "raymarch.npz"), **out)


def composite_resample_golden():
    g = torch.Generator().manual_seed(SEED)
    B, N = 96, 128
    t = torch.sort(torch.rand(B, N + 1, generator=g) * 40, dim=-1).values
    raw_rgb = torch.randn(B, N, 3, generator=g)
    raw_den = torch.randn(B, N, generator=g) * 2
    dirs = torch.randn(B, 3, generator=g) * 1.5
    rgb = torch.sigmoid(raw_rgb)
    den = torch.nn.functional.softplus(raw_den - 1.0)
    comp = O.volumetric_rendering(rgb, den[..., None], t, dirs, False, False)
    w = comp[3]
    u = torch.rand(B, N + 1, generator=g)
    new_det = O.resample_t_vals(t, w, False, 0.01)
    new_rand = O.resample_t_vals(t, w, True, 0.01, u_rand=u)
    np.savez_compressed(os.path.join(HERE, "composite_resample.npz"), t_vals=np32(t), raw_rgb=np32(raw_rgb), raw_density=np32(raw_den),
                        dirs=np32(dirs), comp_rgb=np32(comp[0]), distance=np32(comp[1]), acc=np32(comp[2]), weights=np32(w),
                        t_mids=np32(comp[5]), t_dists=np32(comp[6]), u_rand=np32(u), resampled=np32(new_det),
                        resampled_randomized=np32(new_rand))


def obb_golden():
    sc = H.scene(B=128, K=3, seed=11)
    rays = H.oracle_rays(sc)
    box = torch.from_numpy(sc['centers'])[1]
    ext = torch.from_numpy(sc['ext'])
    B, K = 128, 3
    rot = O.aa2matrix(box[:, 3:])
    o_o, d_o = O.world2object_rpy(rays.origins, rays.directions, box[:, :3].expand(B, K, 3), rot.expand(B, K, 3, 3))
    zi, zo, hit = O.ray_box_intersection(o_o, d_o, -ext.expand(B, K, 3), ext.expand(B, K, 3))
    np.savez_compressed(os.path.join(HERE, "obb.npz"), origins=np32(rays.origins), directions=np32(rays.directions), box=np32(box),
                        ext=np32(ext), rotation=np32(rot), origins_o=np32(o_o), dirs_o=np32(d_o), zi=np32(zi), zo=np32(zo),
                        hit=hit.to(torch.int32).numpy())


def model_golden():
    """Whole dynamic-scene forward (2 levels, background + 2 object MLPs), loss and box-pose gradient on 48 rays."""
    sc = H.scene(B=48, K=2, seed=23)
    cfg = O.ModelConfig(no_pose_opt=False, no_yaw_opt=False)
    out = {}
    for dt, suf in ((torch.float32, ""), (torch.float64, "_f64")):
        params = H.oracle_params(sc, dt)
        params['box_centers'].requires_grad_(True)
        ret = O.model_forward(params, H.oracle_rays(sc, dt), torch.from_numpy(sc['ext']).to(dt), 2, True, False, False, 4.5, cfg=cfg,
                              t_rand=torch.from_numpy(sc['t_rand']).to(dt), u_rand=torch.from_numpy(sc['u_rand']).to(dt))
        tg = {k: torch.from_numpy(v).to(dt) for k, v in sc['targets'].items()}
        loss, stats = O.loss_fn(ret, H.oracle_rays(sc, dt), tg['pixels'], tg['depth'], tg['sky'], eps=3.0)
        gbox = torch.autograd.grad(loss, params['box_centers'])[0]
        for lvl, r in enumerate(ret):
            out.update({f"l{lvl}_comp_rgb{suf}": r.comp_rgb.detach().numpy(), f"l{lvl}_distance{suf}": r.distance.detach().numpy(),
                        f"l{lvl}_acc{suf}": r.acc.detach().numpy(), f"l{lvl}_weights{suf}": r.weights.detach().numpy(),
                        f"l{lvl}_t_vals{suf}": r.t_vals.detach().numpy()})
        out.update({f"loss{suf}": np.asarray(float(loss)), f"d_box_centers{suf}": gbox.numpy()})
        for k in ('losses', 'd_losses', 'n_losses', 'e_losses', 's_losses', 'distr_losses'):
            out[f"{k}{suf}"] = stats[k].detach().numpy()
    np.savez_compressed(os.path.join(HERE, "model_train.npz"), seed=np.int64(23), B=np.int64(48), K=np.int64(2), ts=np.int64(2),
                        alpha=np.float32(4.5), eps=np.float32(3.0), **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    raymarch_golden()
    composite_resample_golden()
    obb_golden()
    model_golden()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
