"""Generate tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES (unmodified, from /root/reference) on
the seeded inputs of tests/ref_cases.py, through the test-only jax / flax / gin stand-in of tests/_refshim.

    python tests/golden/make_ref_golden.py            # rewrite every fixture
    python tests/golden/make_ref_golden.py --check    # regenerate in memory and compare with the committed files

What runs is reference code: internal/{math,mip,mip360,box_helpers,obbpose_model}.py and train_boxpose.train_step
(train_boxpose.py:50-321, loss block + gradient post-processing + Adam).  The primitives underneath are torch-CPU float32
(see tests/_refshim/README.md).  Random draws are logged by the shim's jax.random and injected from the case's own
t_rand / u_rand / noise buffers, so the oracle and the CUDA path can be fed the very same numbers.

The fixtures record sha256 of the reference files they were produced from.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
for p in (ROOT, TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_cases as C            # noqa: E402
import refshim_loader as L       # noqa: E402

REF_FILES = ['internal/math.py', 'internal/mip.py', 'internal/mip360.py', 'internal/box_helpers.py',
             'internal/obbpose_model.py', 'internal/utils.py', 'train_boxpose.py', 'configs/carla_dyn.gin']


def npa(x):
    return np.asarray(x.detach() if hasattr(x, 'detach') else x)


class Gen:
    def __init__(self):
        self.ref = L.load_reference()
        assert self.ref is not None, "needs the reference checkout at /root/reference"
        self.jax, self.jnp, self.flax, self.gin = self.ref.jax, self.ref.jnp, self.ref.flax, self.ref.gin
        self.src = {f: L.source_sha256(f) for f in REF_FILES}

    # -------------------------------------------------------------------------------------------- helpers
    def provider(self, table):
        """jax.random provider: hands out the case's buffers in call order per (kind, shape)."""
        queues = {k: list(v) for k, v in table.items()}

        def fn(kind, path, shape):
            q = queues.get((kind, tuple(shape)))
            assert q, f"the reference drew {kind}{tuple(shape)} at {path}: no buffer of this case was provided for it"
            return q.pop(0)
        return fn

    def meta(self, files):
        return {'ref_sha256__' + f.replace('/', '__').replace('.', '_'): np.array(self.src[f]) for f in files}

    def gin_reset(self, model_over=None, config_over=None):
        self.gin.clear_config()
        self.gin.parse_config_file(os.path.join(self.ref.root, 'configs', 'carla_dyn.gin'))
        for k, v in (model_over or {}).items():
            self.gin.bind_parameter('MipNerfModel.' + k, v)
        for k, v in (config_over or {}).items():
            self.gin.bind_parameter('Config.' + k, v)

    def variables(self, sc):
        jnp = self.jnp

        def tree(layers):
            return {f'Dense_{i}': {'kernel': jnp.array(k), 'bias': jnp.array(b)} for i, (k, b) in enumerate(layers)}
        p = {'MLP_0': tree(sc['mlp']), 'box_centers': jnp.array(sc['centers'])}
        for k, m in enumerate(sc['box_mlps']):
            p[f'BoxMLP_{k}'] = tree(m)
        return self.flax.core.freeze({'params': p})

    def rays(self, sc):
        return self.ref.utils.BoxRays(*[self.jnp.array(a) for a in sc['rays']])

    # -------------------------------------------------------------------------------------------- math.py
    def math(self):
        m, jnp, jax = self.ref.math, self.jnp, self.jax
        I = C.math_inputs()
        out = dict(inputs_sha256=np.array(C.digest(I.values())))
        out['safe_sin'] = npa(m.safe_sin(jnp.array(I['trig_x'])))
        out['safe_cos'] = npa(m.safe_cos(jnp.array(I['trig_x'])))
        bins, w = jnp.array(I['bins']), jnp.array(I['weights'])
        out['pdf_det'] = npa(m.sorted_piecewise_constant_pdf(None, bins, w, C.N + 1, False))
        jax.random.set_provider(self.provider({('uniform', I['u'].shape): [I['u']]}))
        out['pdf_rand'] = npa(m.sorted_piecewise_constant_pdf(jax.random.PRNGKey(0), bins, w, C.N + 1, True))
        jax.random.set_provider(None)
        sb = jnp.array(np.array([[0, 1, 3, 6, 10]], np.float32))
        for i in range(4):
            sw = np.zeros((1, 4), np.float32); sw[0, i] = 1.0
            out[f'pdf_single_bin_{i}'] = npa(m.sorted_piecewise_constant_pdf(None, sb, jnp.array(sw), 625, False))
        cfg = dict(lr_init=5e-4, lr_final=5e-6, max_steps=200000, lr_delay_steps=2500, lr_delay_mult=0.01)
        out['lr'] = np.array([float(m.learning_rate_decay(int(s), **cfg)) for s in I['steps']], np.float64)
        ecfg = dict(lr_init=3.0, lr_final=0.2, max_steps=200000, lr_delay_steps=0, lr_delay_mult=0.01)
        out['eps'] = np.array([float(m.learning_rate_decay(int(s), **ecfg)) for s in I['steps']], np.float64)
        out['alpha'] = np.array([float(m.freq_alpha_rate(int(s), 0.0, 10.0, 1000, 100000)) for s in I['steps']], np.float64)
        out['ssim'] = npa(m.compute_ssim(jnp.array(I['img0']), jnp.array(I['img1']), 1.0))
        out['ssim_map'] = npa(m.compute_ssim(jnp.array(I['img0']), jnp.array(I['img1']), 1.0, return_map=True))
        out['linear_to_srgb'] = npa(m.linear_to_srgb(jnp.array(I['lin'])))
        out['srgb_to_linear'] = npa(m.srgb_to_linear(jnp.array(I['lin'])))
        mse = np.array([1e-4, 0.07, 0.5], np.float32)
        out['mse_to_psnr'] = npa(m.mse_to_psnr(jnp.array(mse)))
        out['psnr_to_mse'] = npa(m.psnr_to_mse(jnp.array(out['mse_to_psnr'])))
        out['avg_error'] = npa(m.compute_avg_error(jnp.array(np.float32(27.5)), jnp.array(np.float32(0.83)), jnp.array(np.float32(0.21))))
        out.update(self.meta(['internal/math.py']))
        return out

    # -------------------------------------------------------------------------------------------- box_helpers.py
    def obb(self):
        bh, jnp = self.ref.box_helpers, self.jnp
        sc = C._scene(B=512, K=4, seed=111, overlap=True)
        out = dict(inputs_sha256=np.array(C.scene_digest(sc)))
        B, K = sc['B'], sc['K']
        box = sc['centers'][2].copy()
        box[3, 3:] = 0.0                                   # zero rotation: theta = 1e-6 + 1e-12 path of aa2matrix
        out['box'] = box
        R = bh.aa2matrix(jnp.array(box[:, 3:]))
        out['aa2matrix'] = npa(R)
        rays = self.rays(sc)
        pose = jnp.broadcast_to(jnp.array(box[:, :3]), [B, K, 3])
        oo, do = bh.world2object_rpy(rays.origins, rays.directions, pose, jnp.broadcast_to(R, [B, K, 3, 3]))
        out['origins_o'], out['dirs_o'] = npa(oo), npa(do)
        dims = jnp.broadcast_to(jnp.array(sc['ext']), [B, K, 3])
        zi, zo, hit = bh.ray_box_intersection(oo, do, -dims, dims)
        out['zi'], out['zo'], out['hit'] = npa(zi), npa(zo), npa(hit).astype(np.int32)
        assert (out['hit'].sum(-1) >= 2).sum() > 10, "overlap case must contain multi-hit rays"
        out.update(self.meta(['internal/box_helpers.py', 'internal/math.py']))
        return out

    # -------------------------------------------------------------------------------------------- mip.py / mip360.py
    def mip(self):
        mip, mip360, jnp, jax = self.ref.mip, self.ref.mip360, self.jnp, self.jax
        sc = C._scene(B=24, K=1, seed=121, far=200.0)
        out = dict(inputs_sha256=np.array(C.scene_digest(sc)))
        r = self.rays(sc)
        B = sc['B']
        t_det, (mean_d, cov_d) = mip.sample_along_rays(None, r.origins, r.directions, r.radii, C.N, r.near, r.far, False, False, 'cone')
        jax.random.set_provider(self.provider({('uniform', (B, C.N + 1)): [sc['t_rand']]}))
        t_rnd, (mean_r, cov_r) = mip.sample_along_rays(jax.random.PRNGKey(0), r.origins, r.directions, r.radii, C.N, r.near, r.far,
                                                       True, False, 'cone')
        jax.random.set_provider(None)
        out['t_det'], out['t_rnd'] = npa(t_det), npa(t_rnd)
        out['cone_mean'], out['cone_cov'] = npa(mean_r), npa(cov_r)
        mc, cc = mip.cast_rays(t_rnd, r.origins, r.directions, r.radii, 'cylinder')
        out['cyl_mean'], out['cyl_cov'] = npa(mc), npa(cc)
        cm, ccov = mip360.new_space((mean_r, cov_r))
        out['contract_mean'], out['contract_cov'] = npa(cm), npa(ccov)
        out['ipe_contracted'] = npa(mip.integrated_pos_enc((cm, ccov), 0, 10))
        out['ipe_plain'] = npa(mip.integrated_pos_enc((mean_d, cov_d), 0, 10))
        # object-frame-like samples (unit directions, metres-scale t): what weighted_ipe sees for a BoxMLP
        dn = r.viewdirs
        t_obj = jnp.array(np.sort(np.random.default_rng(5).uniform(0, 12, size=(B, C.N + 1)).astype(np.float32), -1))
        out['t_obj'] = npa(t_obj)
        mo, co = mip.cast_rays(t_obj, r.origins, dn, r.radii, 'cone')
        for a in (0.0, 3.7, 10.0):
            out[f'wipe_alpha_{a}'] = npa(mip.weighted_ipe((mo[:8], co[:8]), 0, 10, alpha=a))
        out['pos_enc'] = npa(mip.pos_enc(r.viewdirs, min_deg=0, max_deg=4, append_identity=True))
        # volumetric rendering on activated inputs incl. a saturating and an empty ray
        rng = np.random.default_rng(7)
        raw_rgb = rng.standard_normal((B, C.N, 3)).astype(np.float32)
        raw_den = (rng.standard_normal((B, C.N, 1)) * 2).astype(np.float32)
        raw_den[5] = 80.0; raw_den[6] = -60.0
        out['raw_rgb'], out['raw_den'] = raw_rgb, raw_den
        rgb = jax.nn.sigmoid(jnp.array(raw_rgb)); den = jax.nn.softplus(jnp.array(raw_den) - 1.0)
        for tag, white, rand in (('grey', False, False), ('white', True, False), ('randbg', False, True)):
            vr = mip.volumetric_rendering(rgb, den, t_rnd, r.directions, white, rand, jax.random.PRNGKey(3))
            for nm, v in zip(('comp_rgb', 'depth', 'acc', 'weights', 't_vals', 't_mids', 't_dists'), vr):
                if tag == 'grey' or nm in ('comp_rgb',):
                    out[f'vr_{tag}_{nm}'] = npa(v)
        w = jnp.array(out['vr_grey_weights'])
        rs_det, _ = mip.resample_along_rays(None, r.origins, r.directions, r.radii, t_rnd, w, False, 'cone', True, 0.01)
        jax.random.set_provider(self.provider({('uniform', (B, C.N + 1)): [sc['u_rand']]}))
        rs_rnd, _ = mip.resample_along_rays(jax.random.PRNGKey(1), r.origins, r.directions, r.radii, t_rnd, w, True, 'cone', True, 0.01)
        jax.random.set_provider(None)
        out['resample_det'], out['resample_rnd'] = npa(rs_det), npa(rs_rnd)
        out.update(self.meta(['internal/mip.py', 'internal/mip360.py', 'internal/math.py']))
        return out

    # -------------------------------------------------------------------------------------------- obbpose_model.py
    def run_model(self, sc, model_over, ts, randomized, alpha, white_bkgd=False, rand_bkgd=False):
        jax, jnp = self.jax, self.jnp
        self.gin_reset(model_over)
        model = self.ref.obbpose_model.MipNerfModel()
        B = sc['B']
        jax.random.set_provider(self.provider({('uniform', (B, C.N + 1)): [sc['t_rand'], sc['u_rand']],
                                               ('normal', (B, C.N, 1)): [sc['noise'][0], sc['noise'][1]]}))
        jax.random.clear_log()
        ret = model.apply(self.variables(sc), jax.random.PRNGKey(20200823), self.rays(sc), jnp.array(sc['centers']),
                          jnp.array(sc['ext']), jnp.array(np.array([ts], np.int32)), randomized=randomized, rand_bkgd=rand_bkgd,
                          white_bkgd=white_bkgd, alpha=alpha)
        jax.random.set_provider(None)
        return model, ret

    def model(self):
        out = {}
        for name, (skw, mover, akw) in C.MODEL_CASES.items():
            sc = C._scene(**skw)
            model, ret = self.run_model(sc, mover, **akw)
            out[f'{name}/inputs_sha256'] = np.array(C.scene_digest(sc))
            for lvl, r in enumerate(ret):
                for nm, v in zip(('comp_rgb', 'distance', 'acc', 'weights', 't_vals'), r[:5]):
                    out[f'{name}/L{lvl}/{nm}'] = npa(v)
                out[f'{name}/L{lvl}/dyn_mask'] = npa(r[8]).astype(np.float32)
                out[f'{name}/L{lvl}/zo'] = npa(r[9])
            out[f'{name}/off_pose'], out[f'{name}/off_rot'] = npa(ret[-1][7][0]), npa(ret[-1][7][1])
        # the parameter tree the reference itself creates (names / shapes pin the checkpoint layout)
        sc = C._scene(B=8, K=2, seed=1)
        self.gin_reset()
        u = self.ref.utils
        ex = dict(rays=u.namedtuple_map(lambda x: x[None], self.rays(sc)), init=self.jnp.array(sc['centers'])[None],
                  ext=self.jnp.array(sc['ext'])[None], ts=self.jnp.array(np.array([0], np.int32)))
        _, variables = self.ref.obbpose_model.construct_mipnerf(self.jax.random.PRNGKey(0), ex)
        names = []

        def walk(t, pre):
            for k in t:
                if isinstance(t[k], dict):
                    walk(t[k], pre + k + '/')
                else:
                    names.append(f"{pre}{k}:{'x'.join(str(int(s)) for s in t[k].shape)}")
        walk(variables['params'], '')
        out['param_tree'] = np.array(sorted(names))
        out.update(self.meta(['internal/obbpose_model.py', 'internal/mip.py', 'internal/mip360.py', 'internal/box_helpers.py',
                              'internal/math.py', 'configs/carla_dyn.gin']))
        return out

    # -------------------------------------------------------------------------------------------- train_boxpose.py
    def train(self):
        jax, jnp, flax = self.jax, self.jnp, self.flax
        out = {}
        for name, (skw, mover, cover, st) in C.TRAIN_CASES.items():
            sc = C._scene(**skw)
            self.gin_reset(mover, cover)
            model = self.ref.obbpose_model.MipNerfModel()
            config = self.ref.utils.Config()
            V = self.variables(sc)
            state = self.ref.utils.TrainState(optimizer=flax.optim.Adam(config.lr_init).create(V))
            tg, ts = sc['targets'], st['ts']
            batch = dict(rays=self.rays(sc), init=jnp.array(sc['centers']), ext=jnp.array(sc['ext']),
                         ts=jnp.array(np.array([ts], np.int32)), depth=jnp.array(tg['depth']), sky=jnp.array(tg['sky']),
                         pixels=jnp.array(tg['pixels']), target=jnp.array(sc['centers'][ts]))
            prev_ts = ts + 1 if ts == 0 else ts - 1                       # train_boxpose.py:453-456
            prev = jnp.array(sc['centers'][prev_ts])[None]
            B = sc['B']
            jax.random.set_provider(self.provider({('uniform', (B, C.N + 1)): [sc['t_rand'], sc['u_rand']],
                                                   ('normal', (B, C.N, 1)): [sc['noise'][0], sc['noise'][1]]}))
            jax.random.clear_log(); del jax.grad_tap[:]
            new_state, stats, _, pose = self.ref.train_boxpose.train_step(model, config, jax.random.PRNGKey(7), state, batch,
                                                                          st['lr'], st['eps'], st['alpha'], prev)
            jax.random.set_provider(None)
            G = jax.grad_tap[-1]['params']
            NP = new_state.optimizer.target['params']
            P0 = V['params']
            pre = name + '/'
            out[pre + 'inputs_sha256'] = np.array(C.scene_digest(sc))
            for f in ('loss', 'losses', 'obj_losses', 'd_losses', 'n_losses', 'e_losses', 's_losses', 'distr_losses', 'tv_losses',
                      'weight_l2', 'grad_norm', 'grad_abs_max', 'grad_norm_clipped', 'psnrs'):
                out[pre + 'stats/' + f] = npa(getattr(stats, f)).astype(np.float32)
            out[pre + 'pose'] = npa(pose)

            def leaf(tree, nm):
                node = tree
                for part in nm.split('/'):
                    node = node[part]
                return npa(node)
            norms, projs, upd_norms, upd_projs = [], [], [], []
            for nm in C.param_names(sc['K']):
                g = leaf(G, nm).astype(np.float64).reshape(-1)
                d = (leaf(NP, nm).astype(np.float64) - leaf(P0, nm).astype(np.float64)).reshape(-1)
                R = C.projections(nm, g.size).astype(np.float64)
                norms.append(np.linalg.norm(g)); projs.append(R @ g)
                upd_norms.append(np.linalg.norm(d)); upd_projs.append(R @ d)
            out[pre + 'grad_norms'] = np.array(norms); out[pre + 'grad_projs'] = np.stack(projs)
            out[pre + 'update_norms'] = np.array(upd_norms); out[pre + 'update_projs'] = np.stack(upd_projs)
            # small tensors in full: box poses, the two heads, first-layer bias of every network
            out[pre + 'grad/box_centers'] = leaf(G, 'box_centers')
            out[pre + 'new/box_centers'] = leaf(NP, 'box_centers')
            for net in ['MLP_0'] + [f'BoxMLP_{k}' for k in range(sc['K'])]:
                for nm in (f'{net}/Dense_8/kernel', f'{net}/Dense_11/kernel', f'{net}/Dense_11/bias', f'{net}/Dense_0/bias'):
                    out[pre + 'grad/' + nm] = leaf(G, nm)
        out.update(self.meta(['train_boxpose.py', 'internal/obbpose_model.py', 'internal/utils.py', 'internal/mip.py',
                              'internal/math.py', 'configs/carla_dyn.gin']))
        return out


FIXTURES = ('math', 'obb', 'mip', 'model', 'train')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--check', action='store_true')
    ap.add_argument('--only', default=None)
    args = ap.parse_args()
    g = Gen()
    bad = 0
    for name in FIXTURES:
        if args.only and name != args.only:
            continue
        data = getattr(g, name)()
        path = os.path.join(HERE, f'ref_{name}.npz')
        if args.check:
            old = np.load(path)
            for k in data:
                a, b = np.asarray(data[k]), old[k]
                same = a.shape == b.shape and (a.dtype.kind in 'US' and str(a) == str(b) or np.array_equal(a, b, equal_nan=a.dtype.kind == 'f'))
                if not same:
                    bad += 1
                    print(f"MISMATCH {name}:{k}")
            print(f"checked ref_{name}.npz: {len(data)} arrays")
        else:
            np.savez_compressed(path, **data)
            print(f"wrote {path}: {len(data)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")
    sys.exit(1 if bad else 0)


if __name__ == '__main__':
    main()
