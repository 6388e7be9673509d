"""JAX side of the binding: registers the XLA-FFI handlers of libdurf_jax_ffi.so and exposes functions with the
reference's `internal/mip.py` signatures, so `internal/obbpose_model.py` keeps its call sites.

Needs a JAX new enough to have `jax.ffi` (>= 0.4.38) -- not installable in this repository's build image, so this module
is documentation-grade source: it is exercised by nothing here and imports jax lazily.
"""
import ctypes
import os

_LIB = os.environ.get("DURF_JAX_FFI_LIB", os.path.join(os.path.dirname(__file__), "libdurf_jax_ffi.so"))
RM_SAMPLE, RM_RANDOMIZED, RM_CONTRACT, RM_WEIGHTED = 1, 2, 4, 8


def register():
    import jax
    lib = ctypes.CDLL(_LIB)
    for name in ("DurfRaymarchFwd", "DurfCompositeFwd", "DurfResampleFwd", "DurfMlpFwd"):
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name)), platform="CUDA")


def sample_and_encode(t_rand, origins, directions, radii, num_samples, near, far, randomized, contract, min_deg, max_deg):
    """mip.sample_along_rays -> [mip360.new_space] -> mip.integrated_pos_enc in one custom call:
    returns (t_vals[B,N+1], features[B,N,6*(max_deg-min_deg)]).  `t_rand` replaces the PRNG key (explicit U[0,1) draws)."""
    import jax
    import jax.numpy as jnp
    B = origins.shape[0]
    flags = RM_SAMPLE | (RM_RANDOMIZED if randomized else 0) | (RM_CONTRACT if contract else 0)
    out = (jax.ShapeDtypeStruct((B, num_samples + 1), jnp.float32),
           jax.ShapeDtypeStruct((B, num_samples, 6 * (max_deg - min_deg)), jnp.float32))
    empty = jnp.zeros((0,), jnp.float32)
    return jax.ffi.ffi_call("DurfRaymarchFwd", out)(origins, directions, radii.reshape(-1), near.reshape(-1), far.reshape(-1),
                                                    t_rand, empty, num_samples=num_samples, min_deg=min_deg, max_deg=max_deg,
                                                    flags=flags, alpha=0.0)


def volumetric_rendering_raw(raw_rgb, raw_density, t_vals, dirs, white_bkgd, rand_bkgd, density_bias=-1.0):
    """obbpose_model.py:243-245 + mip.volumetric_rendering (mip.py:285-327) -> the reference's 7-tuple."""
    import jax
    import jax.numpy as jnp
    B, N = raw_density.shape[:2]
    f = jnp.float32
    out = (jax.ShapeDtypeStruct((B, 3), f), jax.ShapeDtypeStruct((B,), f), jax.ShapeDtypeStruct((B,), f),
           jax.ShapeDtypeStruct((B, N), f), jax.ShapeDtypeStruct((B, N), f), jax.ShapeDtypeStruct((B, N), f))
    comp_rgb, depth, acc, weights, t_mids, t_dists = jax.ffi.ffi_call("DurfCompositeFwd", out)(
        raw_rgb, raw_density.reshape(B, N), t_vals, dirs, white_bkgd=int(white_bkgd), rand_bkgd=int(rand_bkgd),
        density_bias=float(density_bias))
    return comp_rgb, depth, acc, weights, t_vals, t_mids, t_dists


def resample_along_rays_t(u_rand, t_vals, weights, randomized, resample_padding):
    """The fencepost part of mip.resample_along_rays (mip.py:393-412); the result is stop_gradient'ed like mip.py:413-414."""
    import jax
    import jax.numpy as jnp
    u = u_rand if randomized else jnp.zeros((0,), jnp.float32)
    new_t = jax.ffi.ffi_call("DurfResampleFwd", jax.ShapeDtypeStruct(t_vals.shape, jnp.float32))(
        t_vals, weights, u, resample_padding=float(resample_padding), blurpool=1)
    return jax.lax.stop_gradient(new_t)

# Gradients: wrap each forward in jax.custom_vjp whose backward rule calls the matching *_bwd entry point
# (durf_composite_bwd, durf_mlp_bwd, durf_raymarch_bwd, durf_obb_frontend_bwd) through the same ffi_call mechanism.
