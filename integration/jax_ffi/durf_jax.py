"""JAX side of the binding: registers the XLA-FFI handlers of libdurf_jax_ffi.so (integration/jax_ffi/durf_ffi.cc) and
exposes differentiable functions with the argument meaning of the reference's call sites in
`internal/obbpose_model.py` / `internal/mip.py`, so `MipNerfModel.__call__` keeps its structure and
`jax.value_and_grad(loss_fn)` (train_boxpose.py:251) flows through the CUDA kernels:

    forward handler                 backward handler        jax.custom_vjp below
    DurfObbFrontendFwd              DurfObbFrontendBwd      obb_frontend        (d box_centers[ts])
    DurfRaymarchFwd (fp32 / bf16)   DurfRaymarchBwd         encode_object_rays  (d origins_s, d dirs_s: the box-pose path)
    DurfMlpFwd                      DurfMlpBwd              mlp                 (d params, d features)
    DurfMlpFwdFused (N1)            DurfMlpBwd              background_mlp_fused (d params; the rays carry no gradient)
    DurfCompositeFwd                DurfCompositeBwd        volumetric_rendering_raw (d raw_rgb, d raw_density, d dirs_s)
    DurfResampleFwd                 - (stop_gradient, mip.py:413-414)           resample_along_rays_t
    DurfViewdirEnc, DurfCompactHits, DurfCompactHitsAll, DurfMlpPack            - (no differentiable inputs)
    DurfMlpMergeRaw                 - (linear: the cotangent of the compact rows is a gather)   merge_raw

Needs a JAX with `jax.ffi` (>= 0.4.38).  The build image of this repository has no jax, so nothing here is executed by its
tests; `tests/test_ffi_shim.py` checks what can be checked without it: every handler named here is defined in durf_ffi.cc
with the same operand / attribute / result counts, and durf_ffi.cc compiles against include/durf_b200.h.
jax is imported lazily so that the module itself can be imported (and inspected) anywhere.
"""
import ctypes
import os

_LIB = os.environ.get("DURF_JAX_FFI_LIB", os.path.join(os.path.dirname(__file__), "libdurf_jax_ffi.so"))
_ABI = os.environ.get("DURF_ABI_LIB", os.path.join(os.path.dirname(__file__), "..", "..", "durf_b200", "libdurf_b200.so"))
RM_SAMPLE, RM_RANDOMIZED, RM_CONTRACT, RM_WEIGHTED, RM_CYLINDER, RM_NO_INTEGRATE, RM_OUT_BF16_TILE = 1, 2, 4, 8, 16, 32, 64
PREC_FP32, PREC_BF16 = 0, 1

# handler -> (operands, attributes, results): the contract tests/test_ffi_shim.py checks against durf_ffi.cc
HANDLERS = {
    "DurfObbFrontendFwd": (4, 0, 7),
    "DurfObbFrontendBwd": (7, 2, 1),
    "DurfCompactHits": (1, 1, 2),
    "DurfCompactHitsAll": (1, 0, 2),
    "DurfMlpMergeRaw": (6, 0, 2),
    "DurfRaymarchFwd": (10, 5, 2),
    "DurfRaymarchBwd": (9, 5, 2),
    "DurfViewdirEnc": (1, 1, 1),
    "DurfMlpPack": (1, 6, 1),
    "DurfMlpFwd": (8, 10, 4),
    "DurfMlpFwdFused": (15, 12, 5),
    "DurfMlpBwd": (10, 9, 3),
    "DurfCompositeFwd": (4, 3, 6),
    "DurfCompositeBwd": (8, 3, 3),
    "DurfResampleFwd": (3, 2, 1),
}


def register():
    import jax
    lib = ctypes.CDLL(_LIB)
    for name in HANDLERS:
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name)), platform="CUDA")


class _Abi:
    """Host-side size queries of the C ABI (pure arithmetic, callable at trace time)."""
    _lib = None

    class Topology(ctypes.Structure):
        _fields_ = [(n, ctypes.c_int32) for n in ("in_dim", "width", "depth", "skip", "cond_dim", "cond_width")]

    @classmethod
    def lib(cls):
        if cls._lib is None:
            cls._lib = ctypes.CDLL(_ABI)
            for f in ("durf_mlp_workspace_bytes", "durf_mlp_saved_bytes"):
                getattr(cls._lib, f).restype = ctypes.c_size_t
            cls._lib.durf_mlp_packed_bytes.restype = ctypes.c_int64
        return cls._lib


def _sizes(topo, precision, M, N, training):
    t = _Abi.Topology(*topo)
    lib = _Abi.lib()
    ws = int(lib.durf_mlp_workspace_bytes(ctypes.byref(t), precision, M, N, int(training)))
    saved = int(lib.durf_mlp_saved_bytes(ctypes.byref(t), precision, M, N)) if training else 0
    return max(ws, 16), saved


def _topo_attrs(topo):
    return dict(zip(("in_dim", "width", "depth", "skip", "cond_dim", "cond_width"), (int(x) for x in topo)))


# ---- K0 ------------------------------------------------------------------------------------------------------------
def obb_frontend(origins, directions, box, ext, pose_grad=True, rot_grad=True):
    """world2object_rpy + ray_box_intersection + the scene-graph merge (obbpose_model.py:99-131) for box = box_centers[ts]
    [K,6]: returns (origins_s, dirs_s, hit[B,K] int32, zi, zo, zo_ret, nhit); differentiable w.r.t. `box` through origins_s / dirs_s."""
    import jax
    import jax.numpy as jnp
    B, K = origins.shape[0], box.shape[0]
    f = jnp.float32
    outs = (jax.ShapeDtypeStruct((B, 3), f), jax.ShapeDtypeStruct((B, 3), f), jax.ShapeDtypeStruct((B, K), jnp.int32),
            jax.ShapeDtypeStruct((B, K), f), jax.ShapeDtypeStruct((B, K), f), jax.ShapeDtypeStruct((B,), f),
            jax.ShapeDtypeStruct((B,), f))

    @jax.custom_vjp
    def fe(box_):
        return jax.ffi.ffi_call("DurfObbFrontendFwd", outs)(origins, directions, box_, ext)

    def fe_fwd(box_):
        res = fe(box_)
        return res, (box_, res[2])

    def fe_bwd(saved, cts):
        box_, hit = saved
        d_os, d_ds = cts[0], cts[1]
        d_box = jax.ffi.ffi_call("DurfObbFrontendBwd", jax.ShapeDtypeStruct((K, 6), f), input_output_aliases={6: 0})(
            origins, directions, box_, hit, d_os, d_ds, jnp.zeros((K, 6), f), pose_grad=int(pose_grad), rot_grad=int(rot_grad))
        return (d_box,)

    fe.defvjp(fe_fwd, fe_bwd)
    return fe(box)


def compact_hits(hit, k):
    """Indices of the rays that hit object k (unordered) and their count [1] (obbpose_model.py:174-201 evaluates every ray
    and masks; evaluating the hit rays only is result-identical)."""
    import jax
    import jax.numpy as jnp
    B = hit.shape[0]
    return jax.ffi.ffi_call("DurfCompactHits", (jax.ShapeDtypeStruct((B,), jnp.int32), jax.ShapeDtypeStruct((1,), jnp.int32)))(hit, k=int(k))


def compact_hits_all(hit):
    """`compact_hits` for every object in one custom call: (ray_index [K,B], count [K])."""
    import jax
    import jax.numpy as jnp
    B, K = hit.shape
    return jax.ffi.ffi_call("DurfCompactHitsAll", (jax.ShapeDtypeStruct((K, B), jnp.int32), jax.ShapeDtypeStruct((K,), jnp.int32)))(hit)


def merge_raw(raw_rgb, raw_density, src_rgb, src_density, ray_index, count):
    """raw[ray_index[m]] += src[m] (m < count): adds the compact outputs of an object network evaluated with accumulate = 2
    into the per-ray buffers -- `raw += mask_k * BoxMLP_k(...)` (obbpose_model.py:203-204, 233-234) without making the
    network's call depend on the buffers.  Linear in (raw, src); call it outside differentiated code or wrap it with the
    gather as its cotangent."""
    import jax
    outs = (jax.ShapeDtypeStruct(raw_rgb.shape, raw_rgb.dtype), jax.ShapeDtypeStruct(raw_density.shape, raw_density.dtype))
    return jax.ffi.ffi_call("DurfMlpMergeRaw", outs, input_output_aliases={4: 0, 5: 1})(src_rgb, src_density, ray_index, count,
                                                                                      raw_rgb, raw_density)


# ---- K1 ------------------------------------------------------------------------------------------------------------
def _raymarch(origins, dirs, radii, near, far, t_rand, ray_mult, t_vals, ray_index, count, rows, num_samples, min_deg, max_deg, flags,
              alpha):
    import jax
    import jax.numpy as jnp
    F = 6 * (max_deg - min_deg) + (3 if flags & RM_WEIGHTED else 0)
    feat = (jax.ShapeDtypeStruct((rows, 128 * 64), jnp.bfloat16) if flags & RM_OUT_BF16_TILE
            else jax.ShapeDtypeStruct((rows, num_samples, F), jnp.float32))
    outs = (jax.ShapeDtypeStruct((origins.shape[0], num_samples + 1), jnp.float32), feat)
    return jax.ffi.ffi_call("DurfRaymarchFwd", outs, input_output_aliases={7: 0})(
        origins, dirs, radii.reshape(-1), near, far, t_rand, ray_mult, t_vals, ray_index, count, num_samples=int(num_samples),
        min_deg=int(min_deg), max_deg=int(max_deg), flags=int(flags), alpha=float(alpha))


def sample_and_encode(t_rand, origins, directions, radii, num_samples, near, far, randomized, contract, min_deg, max_deg,
                      ray_mult=None, bf16_tiles=False):
    """mip.sample_along_rays -> [mip360.new_space] -> mip.integrated_pos_enc in one custom call (level 0 of the background):
    returns (t_vals[B,N+1], features).  `t_rand` replaces the PRNG key (explicit U[0,1) draws).  No differentiable input."""
    import jax.numpy as jnp
    B = origins.shape[0]
    e = jnp.zeros((0,), jnp.float32)
    ei = jnp.zeros((0,), jnp.int32)
    flags = RM_SAMPLE | (RM_RANDOMIZED if randomized else 0) | (RM_CONTRACT if contract else 0) | (RM_OUT_BF16_TILE if bf16_tiles else 0)
    return _raymarch(origins, directions, radii, near.reshape(-1), far.reshape(-1), t_rand if randomized else e,
                     e if ray_mult is None else ray_mult, jnp.zeros((B, num_samples + 1), jnp.float32), ei, ei, B, num_samples,
                     min_deg, max_deg, flags, 0.0)


def encode(t_vals, origins, directions, radii, contract, min_deg, max_deg, ray_mult=None, bf16_tiles=False):
    """cast_rays -> [new_space] -> integrated_pos_enc on given (resampled, stop_gradient'ed) fenceposts."""
    import jax.numpy as jnp
    B, S = t_vals.shape
    e = jnp.zeros((0,), jnp.float32)
    ei = jnp.zeros((0,), jnp.int32)
    flags = (RM_CONTRACT if contract else 0) | (RM_OUT_BF16_TILE if bf16_tiles else 0)
    return _raymarch(origins, directions, radii, e, e, e, e if ray_mult is None else ray_mult, t_vals, ei, ei, B, S - 1, min_deg,
                     max_deg, flags, 0.0)[1]


def encode_object_rays(t_vals, origins_s, dirs_s, radii, ray_index, count, rows, alpha, min_deg, max_deg):
    """mip.weighted_ipe (BARF coarse-to-fine weights, mip.py:182-223) of the rays in `ray_index`, fp32 [rows, N, 63];
    differentiable w.r.t. origins_s / dirs_s (the gradient path into the SE(3) box parameters)."""
    import jax
    import jax.numpy as jnp
    B, S = t_vals.shape
    N = S - 1
    e = jnp.zeros((0,), jnp.float32)
    attrs = dict(num_samples=N, min_deg=int(min_deg), max_deg=int(max_deg), flags=RM_WEIGHTED, alpha=float(alpha))

    @jax.custom_vjp
    def enc(o, d):
        return _raymarch(o, d, radii, e, e, e, e, t_vals, ray_index, count, rows, N, min_deg, max_deg, RM_WEIGHTED, alpha)[1]

    def enc_fwd(o, d):
        return enc(o, d), (o, d)

    def enc_bwd(saved, d_feat):
        o, d = saved
        f = jnp.float32
        outs = (jax.ShapeDtypeStruct((B, 3), f), jax.ShapeDtypeStruct((B, 3), f))
        return jax.ffi.ffi_call("DurfRaymarchBwd", outs, input_output_aliases={7: 0, 8: 1})(
            o, d, radii.reshape(-1), t_vals, ray_index, count, d_feat, jnp.zeros((B, 3), f), jnp.zeros((B, 3), f), **attrs)

    enc.defvjp(enc_fwd, enc_bwd)
    return enc(origins_s, dirs_s)


def pos_enc_viewdirs(viewdirs, deg):
    """mip.pos_enc(viewdirs, 0, deg, append_identity=True) (mip.py:36-45)."""
    import jax
    import jax.numpy as jnp
    return jax.ffi.ffi_call("DurfViewdirEnc", jax.ShapeDtypeStruct((viewdirs.shape[0], 3 + 6 * deg), jnp.float32))(viewdirs, deg=int(deg))


# ---- K2 ------------------------------------------------------------------------------------------------------------
def mlp_pack(params, topo):
    """fp32 parameter blob (flax creation order Dense_0.., kernel then bias) -> tensor-core weight image; after every update."""
    import jax
    import jax.numpy as jnp
    t = _Abi.Topology(*topo)
    n = int(_Abi.lib().durf_mlp_packed_bytes(ctypes.byref(t)))
    return jax.ffi.ffi_call("DurfMlpPack", jax.ShapeDtypeStruct((n,), jnp.uint8))(params, **_topo_attrs(topo))


def mlp(params, features, cond, topo, num_rays, num_samples=128, precision=PREC_BF16, packed=None, ray_index=None, count=None,
        into=None, want_d_features=False):
    """MLP.__call__ / BoxMLP.__call__ (obbpose_model.py:294-354, 358-418): returns (raw_rgb[B,N,3], raw_density[B,N]).
    `into` = (raw_rgb, raw_density) of the background: an object network ADDS its hit rows (obbpose_model.py:203-204).
    Differentiable w.r.t. `params` (flat blob) and, with want_d_features (fp32 features; the box-pose path), `features`."""
    import jax
    import jax.numpy as jnp
    f = jnp.float32
    B = cond.shape[0]
    ei = jnp.zeros((0,), jnp.int32)
    ray_index = ei if ray_index is None else ray_index
    count = ei if count is None else count
    packed_ = jnp.zeros((0,), jnp.uint8) if packed is None else packed
    acc = into is not None
    rgb0, den0 = into if acc else (jnp.zeros((B, num_samples, 3), f), jnp.zeros((B, num_samples), f))
    attrs = dict(_topo_attrs(topo), precision=int(precision), num_rays=int(num_rays), num_samples=int(num_samples))
    ws_fwd, _ = _sizes(topo, precision, num_rays, num_samples, False)          # the tensor-core forward needs none (16 B placeholder)
    ws_bwd, saved_bytes = _sizes(topo, precision, num_rays, num_samples, True)    # backward: the dZ records; saved: activations + masks
    n_params = params.shape[0]

    def call_fwd(p, x, training):
        outs = (jax.ShapeDtypeStruct(rgb0.shape, f), jax.ShapeDtypeStruct(den0.shape, f),
                jax.ShapeDtypeStruct((saved_bytes if training else 0,), jnp.uint8), jax.ShapeDtypeStruct((ws_fwd,), jnp.uint8))
        return jax.ffi.ffi_call("DurfMlpFwd", outs, input_output_aliases={6: 0, 7: 1})(
            x, cond, p, packed_, ray_index, count, rgb0, den0, accumulate=int(acc), **attrs)

    @jax.custom_vjp
    def run(p, x):
        r = call_fwd(p, x, False)
        return r[0], r[1]

    def run_fwd(p, x):
        r = call_fwd(p, x, True)
        return (r[0], r[1]), (p, x, r[2])

    def run_bwd(res, cts):
        p, x, saved = res
        d_rgb, d_den = cts
        n_dx = num_rays * num_samples if want_d_features else 0
        outs = (jax.ShapeDtypeStruct((n_params,), f), jax.ShapeDtypeStruct((n_dx, topo[0]) if n_dx else (0,), f),
                jax.ShapeDtypeStruct((ws_bwd,), jnp.uint8))
        d_p, d_x, _ = jax.ffi.ffi_call("DurfMlpBwd", outs, input_output_aliases={9: 0})(
            x, cond, p, packed_, ray_index, count, saved, d_rgb, d_den, jnp.zeros((n_params,), f), **attrs)
        return d_p, (d_x.reshape(x.shape) if want_d_features else jnp.zeros_like(x))

    run.defvjp(run_fwd, run_bwd)
    return run(params, features)


def background_mlp_fused(params, packed, rays_o, rays_d, radii, viewdirs_enc, topo, *, t_vals=None, near=None, far=None, t_rand=None,
                         ray_mult=None, contract=True, min_deg=0, max_deg=10):
    """SURVEY N1 from JAX: sample_along_rays / cast_rays / new_space / integrated_pos_enc AND the background MLP in ONE custom
    call (the tcgen05 kernel generates its input tiles; no feature tensor exists in HBM at inference).  Level 0: pass near / far
    (and t_rand for stratified sampling), t_vals is produced; resampled levels: pass t_vals.  Returns (raw_rgb, raw_density,
    t_vals); differentiable w.r.t. `params` (the background's samples depend on no parameter)."""
    import jax
    import jax.numpy as jnp
    f = jnp.float32
    B, N = rays_o.shape[0], 128
    e, ei = jnp.zeros((0,), f), jnp.zeros((0,), jnp.int32)
    sampling = t_vals is None
    flags = (RM_SAMPLE if sampling else 0) | (RM_RANDOMIZED if (sampling and t_rand is not None) else 0) | (RM_CONTRACT if contract else 0)
    tv0 = jnp.zeros((B, N + 1), f) if sampling else t_vals
    attrs = dict(_topo_attrs(topo), num_rays=B, min_deg=int(min_deg), max_deg=int(max_deg), flags=int(flags), alpha=0.0, accumulate=0)
    ws_bwd, saved_bytes = _sizes(topo, PREC_BF16, B, N, True)
    n_params = params.shape[0]

    def call(p, training):
        outs = (jax.ShapeDtypeStruct((B, N, 3), f), jax.ShapeDtypeStruct((B, N), f), jax.ShapeDtypeStruct((B, N + 1), f),
                jax.ShapeDtypeStruct((B if training else 0, 128 * 64), jnp.bfloat16),
                jax.ShapeDtypeStruct((saved_bytes if training else 0,), jnp.uint8))
        return jax.ffi.ffi_call("DurfMlpFwdFused", outs, input_output_aliases={7: 2, 13: 0, 14: 1})(
            rays_o, rays_d, radii.reshape(-1), e if near is None else near.reshape(-1), e if far is None else far.reshape(-1),
            e if t_rand is None else t_rand, e if ray_mult is None else ray_mult, tv0, viewdirs_enc, p, packed, ei, ei,
            jnp.zeros((B, N, 3), f), jnp.zeros((B, N), f), **attrs)

    @jax.custom_vjp
    def run(p):
        r = call(p, False)
        return r[0], r[1], r[2]

    def run_fwd(p):
        r = call(p, True)
        return (r[0], r[1], r[2]), (p, r[3], r[4])

    def run_bwd(res, cts):
        p, tiles, saved = res
        outs = (jax.ShapeDtypeStruct((n_params,), f), jax.ShapeDtypeStruct((0,), f), jax.ShapeDtypeStruct((ws_bwd,), jnp.uint8))
        d_p, _, _ = jax.ffi.ffi_call("DurfMlpBwd", outs, input_output_aliases={9: 0})(
            tiles, viewdirs_enc, p, packed, ei, ei, saved, cts[0], cts[1], jnp.zeros((n_params,), f), **_topo_attrs(topo),
            precision=PREC_BF16, num_rays=B, num_samples=N)
        return (d_p,)

    run.defvjp(run_fwd, run_bwd)
    return run(params)


# ---- K3 / K4 ---------------------------------------------------------------------------------------------------------
def volumetric_rendering_raw(raw_rgb, raw_density, t_vals, dirs, white_bkgd, rand_bkgd, density_bias=-1.0):
    """obbpose_model.py:243-245 + mip.volumetric_rendering (mip.py:285-327) -> the reference's 7-tuple; differentiable w.r.t.
    raw_rgb, raw_density and dirs (t_vals carry no gradient: mip.py:413-414)."""
    import jax
    import jax.numpy as jnp
    B, N = raw_density.shape[:2]
    f = jnp.float32
    attrs = dict(white_bkgd=int(white_bkgd), rand_bkgd=int(rand_bkgd), density_bias=float(density_bias))
    outs = (jax.ShapeDtypeStruct((B, 3), f), jax.ShapeDtypeStruct((B,), f), jax.ShapeDtypeStruct((B,), f),
            jax.ShapeDtypeStruct((B, N), f), jax.ShapeDtypeStruct((B, N), f), jax.ShapeDtypeStruct((B, N), f))

    @jax.custom_vjp
    def comp(rgb, den, d):
        return jax.ffi.ffi_call("DurfCompositeFwd", outs)(rgb, den.reshape(B, N), t_vals, d, **attrs)

    def comp_fwd(rgb, den, d):
        return comp(rgb, den, d), (rgb, den, d)

    def comp_bwd(saved, cts):
        rgb, den, d = saved
        g_rgb, g_depth, g_acc, g_w = cts[0], cts[1], cts[2], cts[3]      # t_mids / t_dists depend on t_vals only
        bouts = (jax.ShapeDtypeStruct((B, N, 3), f), jax.ShapeDtypeStruct((B, N), f), jax.ShapeDtypeStruct((B, 3), f))
        d_rgb, d_den, d_dirs = jax.ffi.ffi_call("DurfCompositeBwd", bouts)(rgb, den.reshape(B, N), t_vals, d, g_rgb, g_depth, g_acc,
                                                                            g_w, **attrs)
        return d_rgb, d_den.reshape(den.shape), d_dirs

    comp.defvjp(comp_fwd, comp_bwd)
    comp_rgb, depth, acc, weights, t_mids, t_dists = comp(raw_rgb, raw_density, dirs)
    return comp_rgb, depth, acc, weights, t_vals, t_mids, t_dists


def resample_along_rays_t(u_rand, t_vals, weights, randomized, resample_padding):
    """The fencepost part of mip.resample_along_rays (mip.py:393-412); the result is stop_gradient'ed like mip.py:413-414."""
    import jax
    import jax.numpy as jnp
    u = u_rand if randomized else jnp.zeros((0,), jnp.float32)
    new_t = jax.ffi.ffi_call("DurfResampleFwd", jax.ShapeDtypeStruct(t_vals.shape, jnp.float32))(
        t_vals, jax.lax.stop_gradient(weights), u, resample_padding=float(resample_padding), blurpool=1)
    return jax.lax.stop_gradient(new_t)
