"""Seeded inputs of the reference-executed golden fixtures (tests/golden/ref_*.npz).

Shared by the generator (tests/golden/make_ref_golden.py, which runs the reference's own code on these inputs) and by
the tests that replay them through the oracle and the CUDA path.  numpy only; every case is a pure function of its seed.
The fixtures store a sha256 of the inputs, so a drift of this file or of durf_b200.synthetic is caught, not absorbed.
"""
import hashlib

import numpy as np

from durf_b200 import synthetic as S

N = 128


def digest(arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode()); h.update(str(a.shape).encode()); h.update(a.tobytes())
    return h.hexdigest()


# ------------------------------------------------------------------------------------------------ math.py
def math_inputs():
    x = (10.0 ** np.linspace(-30, 10, 1500))
    x = np.concatenate([-x[::-1], np.array([0.0]), x]).astype(np.float32)
    t = np.float32(100.0 * np.pi)
    k = np.arange(-6, 7, dtype=np.float32)
    edge = np.concatenate([k * t, np.nextafter(k * t, np.float32(np.inf)), np.nextafter(k * t, np.float32(-np.inf)),
                           k * t + np.float32(1.25), k * t - np.float32(0.75)]).astype(np.float32)
    rng = np.random.default_rng(101)
    wide = (rng.standard_normal(1500) * 2.0e4).astype(np.float32)           # |2^9 x| at object-frame magnitudes
    trig_x = np.concatenate([x, edge, wide])
    bins = np.sort(rng.uniform(0.0, 40.0, size=(48, N + 1)).astype(np.float32), axis=-1)
    w = (rng.uniform(size=(48, N)) ** 4).astype(np.float32)
    w[0] = 0.0                                  # the eps-padding path (math.py:237-241)
    w[1] = 0.0; w[1, 77] = 1.0                  # delta
    w[2] = 1.0                                  # flat
    w[3, :64] = 0.0                             # empty head
    u = rng.integers(0, 1 << 23, size=(48, N + 1)).astype(np.float32) * np.float32(2.0 ** -23)
    img0 = rng.uniform(size=(2, 24, 20, 3)).astype(np.float32)
    img1 = np.clip(img0 + rng.standard_normal(img0.shape).astype(np.float32) * 0.1, 0, 1).astype(np.float32)
    lin = np.concatenate([np.linspace(0, 1, 257), [0.0031308, 0.0031309, 0.04045, 0.04046]]).astype(np.float32)
    steps = np.array([0, 1, 10, 1250, 2499, 2500, 2501, 50000, 100000, 199999, 200000, 250000], np.int64)
    return dict(trig_x=trig_x, bins=bins, weights=w, u=u, img0=img0, img1=img1, lin=lin, steps=steps)


# ------------------------------------------------------------------------------------------------ scenes
def _scene(B, K, seed, behind=False, far=40.0, bias_scale=0.05, overlap=False, weight_gain=1.0):
    rng = np.random.default_rng(seed)
    rays, c2w = S.random_rays(rng, B, far=far)
    centers, ext = S.boxes_in_view(rng, c2w, K, behind=behind)
    if overlap and K >= 2:
        # put box 1 on box 0's line of sight (farther away): rays then cross two boxes, which the reference SUMS
        # (obbpose_model.py:118-122 "assumes that objects do not occlude each other")
        cam = c2w[:3, 3]
        centers[:, 1, :3] = cam + 1.35 * (centers[:, 0, :3] - cam)
        ext[1] = ext[0] * 1.2
        # aim a third of the rays at box 0 so the overlap region is well populated
        n = B // 3
        tgt = centers[2, 0, :3] + rng.uniform(-1, 1, size=(n, 3)).astype(np.float32) * ext[0] * 0.9
        d = tgt - cam
        d = d / -(d @ c2w[:3, 2])[:, None]                       # camera-frame z = -1 like _generate_rays_multi
        rays.directions[:n] = d.astype(np.float32)
        rays.viewdirs[:n] = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    mlp = S.glorot_mlp(rng, 60, 256, bias_scale)
    box_mlps = [S.glorot_mlp(rng, 63, 128, bias_scale) for _ in range(K)]
    if weight_gain != 1.0:
        # a network whose activations saturate (sigmoid / softplus away from their linear regime, dense ReLU flips):
        # random-init glorot weights give rgb ~ 0.5 and tiny densities, which makes bf16 tolerances nearly vacuous
        mlp = [(k * np.float32(weight_gain if i in (0, 8, 11) else 1.0), b) for i, (k, b) in enumerate(mlp)]
        box_mlps = [[(k * np.float32(weight_gain if i in (0, 8, 11) else 1.0), b) for i, (k, b) in enumerate(m)] for m in box_mlps]
    tg = S.targets(rng, B)
    t_rand = rng.integers(0, 1 << 23, size=(B, N + 1)).astype(np.float32) * np.float32(2.0 ** -23)
    u_rand = rng.integers(0, 1 << 23, size=(B, N + 1)).astype(np.float32) * np.float32(2.0 ** -23)
    noise = rng.standard_normal((2, B, N, 1)).astype(np.float32)
    return dict(rays=rays, c2w=c2w, centers=centers, ext=ext, mlp=mlp, box_mlps=box_mlps, targets=tg, t_rand=t_rand,
                u_rand=u_rand, noise=noise, B=B, K=K)


def scene_digest(sc) -> str:
    arrs = list(sc['rays']) + [sc['centers'], sc['ext'], sc['t_rand'], sc['u_rand'], sc['noise']]
    arrs += [a for kb in sc['mlp'] for a in kb] + [a for m in sc['box_mlps'] for kb in m for a in kb]
    arrs += [sc['targets'][k] for k in ('pixels', 'depth', 'sky')]
    return digest(arrs)


# name -> (scene kwargs, MipNerfModel field overrides on top of configs/carla_dyn.gin, apply kwargs)
MODEL_CASES = {
    # BASELINE configs[0]: static background, no contraction, deterministic sampling -> the pure mip.py path
    'c1_static': (dict(B=128, K=1, seed=201, behind=True),
                  dict(dynamics=False, contraction=False), dict(ts=0, randomized=False, alpha=10.0)),
    # configs[1]: mip360 contraction + hierarchical resampling, static
    'c2_contract': (dict(B=128, K=1, seed=202, behind=True, far=200.0),
                    dict(dynamics=False, contraction=True), dict(ts=0, randomized=False, alpha=10.0)),
    # configs[2]: dynamic scene graph, 2 objects, randomized sampling
    'c3_dynamic': (dict(B=192, K=2, seed=203), dict(), dict(ts=2, randomized=True, alpha=10.0)),
    # configs[3]: 8 objects, rays crossing two boxes included
    'c4_k8_overlap': (dict(B=256, K=8, seed=204, overlap=True), dict(num_objects=8), dict(ts=1, randomized=False, alpha=10.0)),
    # configs[4]: BARF coarse-to-fine weights mid-schedule
    'c5_barf': (dict(B=128, K=2, seed=205), dict(), dict(ts=3, randomized=True, alpha=2.5)),
    # density noise + cylinder rays + white background
    'c6_noise_cyl': (dict(B=96, K=2, seed=206), dict(density_noise=0.3, ray_shape='cylinder'),
                     dict(ts=4, randomized=True, alpha=6.25, white_bkgd=True)),
    # a saturating network (gain x3 on the first layer and the heads): the meaningful bf16 case
    'c7_gain3': (dict(B=128, K=2, seed=207, weight_gain=3.0), dict(), dict(ts=0, randomized=False, alpha=10.0)),
}

# name -> (scene kwargs, model overrides, Config overrides, step kwargs)
TRAIN_CASES = {
    't_default': (dict(B=160, K=2, seed=301), dict(), dict(), dict(ts=2, lr=3e-4, eps=2.5, alpha=10.0)),
    't_pose': (dict(B=160, K=2, seed=302), dict(no_pose_opt=False, no_yaw_opt=False), dict(),
               dict(ts=1, lr=1e-3, eps=3.0, alpha=2.5)),
    't_extras': (dict(B=128, K=2, seed=303), dict(no_pose_opt=False, no_yaw_opt=False, density_noise=0.2),
                 dict(box_loss_mult=2, tv_loss_mult=0.05, weight_decay_mult=0.01, coarse_loss_mult=0.3),
                 dict(ts=3, lr=5e-4, eps=1.0, alpha=7.0)),
    't_noclip_single': (dict(B=96, K=1, seed=304), dict(num_objects=1),
                        dict(grad_max_val=0.0, grad_max_norm=0.0, disable_multiscale_loss=True),
                        dict(ts=0, lr=5e-4, eps=0.5, alpha=10.0)),
}

N_PROJ = 32


def projections(name: str, n: int) -> np.ndarray:
    """[N_PROJ, n] Rademacher vectors, a pure function of the tensor's name and size: gradients of whole networks are
    pinned by their L2 norm plus N_PROJ random projections instead of megabytes of values."""
    seed = int.from_bytes(hashlib.sha256(name.encode()).digest()[:8], 'little')
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 2, size=(N_PROJ, n), dtype=np.int8) * 2 - 1).astype(np.float32)


def param_names(K: int):
    """Leaf order used for gradient summaries: MLP_0, BoxMLP_k (Dense_0..11 kernel then bias), box_centers."""
    out = []
    for net in ['MLP_0'] + [f'BoxMLP_{k}' for k in range(K)]:
        for i in range(12):
            out += [f'{net}/Dense_{i}/kernel', f'{net}/Dense_{i}/bias']
    return out + ['box_centers']
