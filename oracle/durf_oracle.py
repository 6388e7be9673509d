"""CPU oracle for the DURF per-ray Mip-NeRF hot path.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, in PyTorch-CPU, of the arithmetic of the reference
(FelTris/durf, pure Python/JAX).  It exists so that the CUDA kernels in
``durf_b200/csrc`` can be checked for parity.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it; nothing under ``durf_b200/`` does.

PARITY STATUS: **parity unpinned** except for the inverse-CDF sampler and the safe
trigonometry.  The reference cannot be imported here (jax/jaxlib/flax/gin are not
installed and there is no network) and its only tests (internal/math_test.py) are
property tests of internal/math.py; those are ported in tests/test_oracle_reference_props.py
and pin `sorted_piecewise_constant_pdf`, `safe_sin/safe_cos`, `learning_rate_decay`.
Every other function below is pinned only by (a) line-by-line restatement with the
reference file:line cited in each docstring, (b) analytic / Monte-Carlo / finite-difference
checks in tests/test_oracle_analytic.py and (c) fp32-vs-fp64 self-consistency.

All functions are dtype-generic: they compute in the dtype of their tensor inputs
(float32 = what the reference computes on CPU; float64 for conditioning studies).
Random draws of the reference (threefry streams that cannot be reproduced without JAX)
enter as explicit tensors: `t_rand`, `u_rand`, `density_noise`.

JAX semantics that matter and are reproduced:
  * `jnp.nan_to_num(x, <positional 2nd arg>)` sets `copy`, not the fill value -> NaN -> 0,
    +-inf -> +-finfo.max  (mip.py:313, mip.py:320, math.py:282).
  * `x % t` on floats is floored remainder built on an exact fmod (math.py:36).
  * Python float constants are rounded to the array dtype when they meet an array
    (0.5*pi, 100*pi, 1/3, 4/15 ...).
  * flax `nn.Dense`: y = x @ kernel[in, out] + bias.
"""
from __future__ import annotations

import math as _pymath
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor

F32_EPS = float(np.finfo(np.float32).eps)


# --------------------------------------------------------------------------------------
# internal/math.py
# --------------------------------------------------------------------------------------

def safe_norm(x: Tensor) -> Tensor:
    """reference internal/math.py:27-32  sqrt(where(|x|^2 < 1e-12, 1e-12, |x|^2)), keepdims."""
    sq = torch.sum(x * x, dim=-1, keepdim=True)
    sq = torch.where(sq < 1e-12, torch.full_like(sq, 1e-12), sq)
    return torch.sqrt(sq)


def _floored_remainder(x: Tensor, t: float) -> Tensor:
    """jnp `%`: exact fmod followed by a sign fix (floored remainder)."""
    tt = torch.full_like(x, t)
    r = torch.fmod(x, tt)
    fix = (r != 0) & ((r < 0) != (tt < 0))
    return torch.where(fix, r + tt, r)


def _safe_trig(x: Tensor, fn) -> Tensor:
    """reference internal/math.py:35-36  fn(where(|x| < 100*pi, x, x % (100*pi)))."""
    t = 100.0 * _pymath.pi
    tt = torch.full_like(x, t)  # rounds 100*pi to the array dtype, like the weak-typed jnp scalar
    return fn(torch.where(torch.abs(x) < tt, x, _floored_remainder(x, t)))


def safe_sin(x: Tensor) -> Tensor:
    """reference internal/math.py:44-46."""
    return _safe_trig(x, torch.sin)


def safe_cos(x: Tensor) -> Tensor:
    """reference internal/math.py:39-41."""
    return _safe_trig(x, torch.cos)


def mse_to_psnr(mse):
    """reference internal/math.py:49-51."""
    return -10.0 / _pymath.log(10.0) * torch.log(torch.as_tensor(mse))


def psnr_to_mse(psnr):
    """reference internal/math.py:54-56."""
    return torch.exp(-0.1 * _pymath.log(10.0) * torch.as_tensor(psnr))


def learning_rate_decay(step, lr_init, lr_final, max_steps, lr_delay_steps=0, lr_delay_mult=1.0):
    """reference internal/math.py:156-190 (host scalar; also drives the eps schedule,
    train_boxpose.py:355-361)."""
    if lr_delay_steps > 0:
        delay_rate = lr_delay_mult + (1 - lr_delay_mult) * _pymath.sin(
            0.5 * _pymath.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
    else:
        delay_rate = 1.0
    t = min(max(step / max_steps, 0.0), 1.0)
    log_lerp = _pymath.exp(_pymath.log(lr_init) * (1 - t) + _pymath.log(lr_final) * t)
    return delay_rate * log_lerp


def freq_alpha_rate(step, alpha_init, alpha_final, alpha_delay_steps, alpha_max_steps):
    """reference internal/math.py:193-219 (BARF alpha schedule)."""
    if step < alpha_delay_steps:
        return alpha_init
    if step < alpha_max_steps:
        return (step - alpha_delay_steps) / (alpha_max_steps - alpha_delay_steps) * alpha_final
    return alpha_final


def sorted_piecewise_constant_pdf(bins: Tensor, weights: Tensor, num_samples: int,
                                  randomized: bool, u_rand: Optional[Tensor] = None,
                                  chunk: int = 0) -> Tensor:
    """reference internal/math.py:222-284.

    `u_rand` ([..., num_samples], U[0,1)) replaces the threefry draw at math.py:257-260:
    jax.random.uniform(maxval=m) returns u01 * m, so jitter = u_rand * (s - eps32).
    The interval search is the reference's dense mask/max/min formulation (math.py:270-280).
    """
    dt = weights.dtype
    eps = 1e-5
    weight_sum = torch.sum(weights, dim=-1, keepdim=True)
    padding = torch.clamp(eps - weight_sum, min=0.0)
    weights = weights + padding / weights.shape[-1]
    weight_sum = weight_sum + padding

    pdf = weights / weight_sum
    cdf = torch.clamp(torch.cumsum(pdf[..., :-1], dim=-1), max=1.0)
    lead = list(cdf.shape[:-1])
    cdf = torch.cat([torch.zeros(lead + [1], dtype=dt), cdf, torch.ones(lead + [1], dtype=dt)], dim=-1)

    if randomized:
        assert u_rand is not None, "randomized sampling needs the explicit u_rand buffer"
        s = 1.0 / num_samples
        u = torch.arange(num_samples, dtype=dt) * torch.tensor(s, dtype=dt)
        u = u + u_rand.to(dt) * torch.tensor(s - F32_EPS, dtype=dt)
        u = torch.clamp(u, max=float(np.float32(1.0 - F32_EPS)) if dt == torch.float32 else 1.0 - F32_EPS)
    else:
        stop = torch.tensor(1.0 - F32_EPS, dtype=dt)
        frac = torch.arange(num_samples, dtype=dt) / (num_samples - 1)
        u = stop * frac
        u[-1] = stop
        u = u.expand(lead + [num_samples])
    u = u.contiguous()

    def search(lo: int, hi: int):
        uu, cc, bb = u[lo:hi], cdf[lo:hi], bins[lo:hi]
        mask = uu[..., None, :] >= cc[..., :, None]

        def find(x):
            x0 = torch.max(torch.where(mask, x[..., None], x[..., :1, None]), dim=-2).values
            x1 = torch.min(torch.where(~mask, x[..., None], x[..., -1:, None]), dim=-2).values
            return x0, x1

        b0, b1 = find(bb)
        c0, c1 = find(cc)
        t = torch.clamp(torch.nan_to_num((uu - c0) / (c1 - c0)), 0.0, 1.0)
        return b0 + t * (b1 - b0)

    assert bins.dim() >= 2, "bins/weights carry at least one batch dimension"
    nb = bins.shape[0]
    if chunk <= 0 or nb <= chunk:
        return search(0, nb)
    return torch.cat([search(i, min(i + chunk, nb)) for i in range(0, nb, chunk)], 0)


# --------------------------------------------------------------------------------------
# internal/mip.py
# --------------------------------------------------------------------------------------

def pos_enc(x: Tensor, min_deg: int, max_deg: int, append_identity: bool = True) -> Tensor:
    """reference internal/mip.py:36-45; layout index = l*3+d, then the same shifted by pi/2."""
    scales = torch.tensor([2.0 ** i for i in range(min_deg, max_deg)], dtype=x.dtype)
    xb = (x[..., None, :] * scales[:, None]).reshape(list(x.shape[:-1]) + [-1])
    half_pi = torch.tensor(0.5 * _pymath.pi, dtype=x.dtype)
    four = torch.sin(torch.cat([xb, xb + half_pi], dim=-1))
    return torch.cat([x, four], dim=-1) if append_identity else four


def expected_sin(x: Tensor, x_var: Tensor) -> Tensor:
    """reference internal/mip.py:67-73 (first return value only: the callers take [0])."""
    return torch.exp(-0.5 * x_var) * safe_sin(x)


def lift_gaussian(d: Tensor, t_mean: Tensor, t_var: Tensor, r_var: Tensor) -> Tuple[Tensor, Tensor]:
    """reference internal/mip.py:76-96 with diag=False (the only mode the model uses)."""
    mean = d[..., None, :] * t_mean[..., None]
    d_mag_sq = torch.clamp(torch.sum(d * d, dim=-1, keepdim=True), min=1e-10)
    d_outer = d[..., :, None] * d[..., None, :]
    eye = torch.eye(d.shape[-1], dtype=d.dtype)
    null_outer = eye - d[..., :, None] * (d / d_mag_sq)[..., None, :]
    t_cov = t_var[..., None, None] * d_outer[..., None, :, :]
    xy_cov = r_var[..., None, None] * null_outer[..., None, :, :]
    return mean, t_cov + xy_cov


def conical_frustum_to_gaussian(d, t0, t1, base_radius):
    """reference internal/mip.py:99-130, stable=True, diag=False."""
    mu = (t0 + t1) / 2
    hw = (t1 - t0) / 2
    t_mean = mu + (2 * mu * hw ** 2) / (3 * mu ** 2 + hw ** 2)
    t_var = (hw ** 2) / 3 - (4 / 15) * ((hw ** 4 * (12 * mu ** 2 - hw ** 2)) /
                                        (3 * mu ** 2 + hw ** 2) ** 2)
    r_var = base_radius ** 2 * ((mu ** 2) / 4 + (5 / 12) * hw ** 2 - 4 / 15 *
                                (hw ** 4) / (3 * mu ** 2 + hw ** 2))
    return lift_gaussian(d, t_mean, t_var, r_var)


def cylinder_to_gaussian(d, t0, t1, radius):
    """reference internal/mip.py:133-152, diag=False."""
    t_mean = (t0 + t1) / 2
    r_var = radius ** 2 / 4
    t_var = (t1 - t0) ** 2 / 12
    return lift_gaussian(d, t_mean, t_var, r_var)


def cast_rays(t_vals, origins, directions, radii, ray_shape: str = 'cone'):
    """reference internal/mip.py:155-179."""
    t0 = t_vals[..., :-1]
    t1 = t_vals[..., 1:]
    if ray_shape == 'cone':
        fn = conical_frustum_to_gaussian
    elif ray_shape == 'cylinder':
        fn = cylinder_to_gaussian
    else:
        raise AssertionError("ray_shape must be 'cone' or 'cylinder'")  # mip.py:176 `assert False`
    means, covs = fn(directions, t0, t1, radii)
    return means + origins[..., None, :], covs


def _ipe_lift(x: Tensor, x_cov: Tensor, min_deg: int, max_deg: int):
    """reference internal/mip.py:273-278 (and 208-213): basis = [2^l * I_3]_l, so
    y[l*3+d] = 2^l x_d and y_var[l*3+d] = 4^l cov[d,d]."""
    dt = x.dtype
    basis = torch.cat([(2.0 ** i) * torch.eye(3, dtype=dt) for i in range(min_deg, max_deg)], dim=1)
    y = torch.matmul(x, basis)
    y_var = torch.sum(torch.matmul(x_cov, basis) * basis, dim=-2)
    return y, y_var


def integrated_pos_enc(x_coord, min_deg: int, max_deg: int) -> Tensor:
    """reference internal/mip.py:226-282, diag=False."""
    x, x_cov = x_coord
    y, y_var = _ipe_lift(x, x_cov, min_deg, max_deg)
    half_pi = torch.tensor(0.5 * _pymath.pi, dtype=x.dtype)
    return expected_sin(torch.cat([y, y + half_pi], dim=-1), torch.cat([y_var, y_var], dim=-1))


def barf_weights(alpha: float, max_deg: int, dtype=torch.float32) -> Tensor:
    """reference internal/mip.py:217-218  w_k = (1 - cos(clip(alpha - k, 0, 1) * pi)) / 2."""
    k = torch.arange(max_deg, dtype=dtype)
    a = torch.as_tensor(alpha, dtype=dtype)
    return (1 - torch.cos(torch.clamp(a - k, 0, 1) * torch.tensor(_pymath.pi, dtype=dtype))) / 2


def weighted_ipe(x_coord, min_deg: int, max_deg: int, alpha) -> Tensor:
    """reference internal/mip.py:182-223, diag=False.  Quirk kept: the weight is laid out
    [max_deg, 6] and flattened (mip.py:220), so feature i is scaled by w[i // 6]."""
    x, x_cov = x_coord
    y, y_var = _ipe_lift(x, x_cov, min_deg, max_deg)
    half_pi = torch.tensor(0.5 * _pymath.pi, dtype=x.dtype)
    enc = expected_sin(torch.cat([y, y + half_pi], dim=-1), torch.cat([y_var, y_var], dim=-1))
    w = barf_weights(alpha, max_deg, x.dtype)
    w = w[:, None].expand(max_deg, 6).reshape(-1)
    return torch.cat([x, w * enc], dim=-1)


def volumetric_rendering(rgb, density, t_vals, dirs, white_bkgd: bool, rand_bkgd: bool):
    """reference internal/mip.py:285-327.  Returns the 7-tuple
    (comp_rgb, depth, acc, weights, t_vals, t_mids, t_dists); `depth` is the un-normalised
    sum(w * t_mid) (mip.py:317,327).  rand_bkgd adds randint(.., 0, 1) == 0 (mip.py:324)."""
    t_mids = 0.5 * (t_vals[..., :-1] + t_vals[..., 1:])
    t_dists = t_vals[..., 1:] - t_vals[..., :-1]
    delta = t_dists * torch.linalg.norm(dirs[..., None, :], dim=-1)
    density_delta = density[..., 0] * delta
    alpha = 1 - torch.exp(-density_delta)
    trans = torch.exp(-torch.cat([
        torch.zeros_like(density_delta[..., :1]),
        torch.cumsum(density_delta[..., :-1], dim=-1)], dim=-1))
    weights = torch.nan_to_num(alpha * trans)
    comp_rgb = (weights[..., None] * rgb).sum(dim=-2)
    acc = weights.sum(dim=-1)
    depth = (weights * t_mids).sum(dim=-1)
    if white_bkgd:
        comp_rgb = comp_rgb + (1.0 - acc[..., None])
    if rand_bkgd:
        comp_rgb = comp_rgb + 0.0 * (1.0 - acc[..., None])
    elif not white_bkgd:
        comp_rgb = comp_rgb + 0.5 * (1.0 - acc[..., None])
    return comp_rgb, depth, acc, weights, t_vals, t_mids, t_dists


def sample_t_vals(near: Tensor, far: Tensor, num_samples: int, randomized: bool,
                  t_rand: Optional[Tensor] = None, lindisp: bool = False) -> Tensor:
    """reference internal/mip.py:351-368 (the t_vals part of sample_along_rays)."""
    dt = near.dtype
    s = torch.linspace(0.0, 1.0, num_samples + 1, dtype=dt)
    t_vals = near * (1.0 - s) + far * s
    if lindisp:
        t_vals = 1.0 / t_vals
    if randomized:
        assert t_rand is not None
        mids = 0.5 * (t_vals[..., 1:] + t_vals[..., :-1])
        upper = torch.cat([mids, t_vals[..., -1:]], -1)
        lower = torch.cat([t_vals[..., :1], mids], -1)
        t_vals = lower + (upper - lower) * t_rand.to(dt)
    else:
        t_vals = t_vals.expand(near.shape[0], num_samples + 1)
    return t_vals


def sample_along_rays(origins, directions, radii, num_samples, near, far, randomized,
                      lindisp=False, ray_shape='cone', t_rand=None):
    """reference internal/mip.py:330-370."""
    t_vals = sample_t_vals(near, far, num_samples, randomized, t_rand, lindisp)
    return t_vals, cast_rays(t_vals, origins, directions, radii, ray_shape)


def resample_t_vals(t_vals, weights, randomized, resample_padding, u_rand=None, chunk=1024):
    """reference internal/mip.py:393-412 (blur-pool + inverse-CDF; num_samples = t_vals.shape[-1])."""
    wp = torch.cat([weights[..., :1], weights, weights[..., -1:]], dim=-1)
    wmax = torch.maximum(wp[..., :-1], wp[..., 1:])
    wblur = 0.5 * (wmax[..., :-1] + wmax[..., 1:])
    w = wblur + resample_padding
    return sorted_piecewise_constant_pdf(t_vals, w, t_vals.shape[-1], randomized, u_rand, chunk=chunk)


def resample_along_rays(origins, directions, radii, t_vals, weights, randomized, ray_shape='cone',
                        stop_grad=True, resample_padding=0.01, u_rand=None):
    """reference internal/mip.py:373-416."""
    new_t = resample_t_vals(t_vals, weights, randomized, resample_padding, u_rand)
    if stop_grad:
        new_t = new_t.detach()
    return new_t, cast_rays(new_t, origins, directions, radii, ray_shape)


# --------------------------------------------------------------------------------------
# internal/mip360.py
# --------------------------------------------------------------------------------------

def contract(x: Tensor) -> Tensor:
    """reference internal/mip360.py:47-60.  Threshold 0.1 (not 1) with the |x|-1 formula: kept."""
    n = safe_norm(x)
    smaller = (n <= 0.1).to(x.dtype)
    larger = (n > 0.1).to(x.dtype)
    xc = (2.0 - torch.nan_to_num(1.0 / n)) * torch.nan_to_num(x / n)
    return smaller * x + larger * xc


def new_space(samples):
    """reference internal/mip360.py:63-79.  v = JVP of `contract` at `mean` along an all-ones
    tangent (jax.linearize, :72-73); cov' = cov @ diag(v)^2, i.e. cov'[i,j] = cov[i,j] v_j^2 (:77)."""
    mean, cov = samples
    meanc, v = torch.func.jvp(contract, (mean,), (torch.ones_like(mean),))
    eye = torch.eye(3, dtype=mean.dtype)
    dv = v[..., :, None] * eye
    covc = torch.matmul(dv, torch.matmul(cov, dv).transpose(-1, -2)).transpose(-1, -2)
    return meanc, covc


# --------------------------------------------------------------------------------------
# internal/box_helpers.py
# --------------------------------------------------------------------------------------

def aa2matrix(angles: Tensor) -> Tensor:
    """reference internal/box_helpers.py:148-167 (Rodrigues, theta = safe_norm + 1e-12)."""
    zero = torch.zeros_like(angles[:, 0:1])
    r0 = torch.cat([zero, -angles[:, 2:3], angles[:, 1:2]], dim=-1)
    r1 = torch.cat([angles[:, 2:3], zero, -angles[:, 0:1]], dim=-1)
    r2 = torch.cat([-angles[:, 1:2], angles[:, 0:1], zero], dim=-1)
    skew = torch.stack([r0, r1, r2], dim=-2)
    th = safe_norm(angles) + 1e-12
    eye = torch.eye(3, dtype=angles.dtype).expand(angles.shape[0], 3, 3)
    return eye + (torch.sin(th) / th)[..., None] * skew + \
        ((1 - torch.cos(th)) / th ** 2)[..., None] * torch.matmul(skew, skew)


def world2object_rpy(pts: Tensor, dirs: Tensor, pose: Tensor, rot: Tensor):
    """reference internal/box_helpers.py:286-341 (dim=None, inverse=False):
    o_o = R o + R(-p);  d_o = R d / |R d|.   pts,dirs [B,3]; pose [B,K,3]; rot [B,K,3,3]."""
    t_w_o = torch.matmul(rot, (-pose)[..., None])[..., 0]
    pts_o = torch.matmul(rot, pts[:, None, :, None])[..., 0] + t_w_o
    dirs_o = torch.matmul(rot, dirs[:, None, :, None])[..., 0]
    dirs_o = dirs_o / torch.linalg.norm(dirs_o, dim=-1, keepdim=True)
    return pts_o, dirs_o


def ray_box_intersection(ray_o: Tensor, ray_d: Tensor, aabb_min: Tensor, aabb_max: Tensor):
    """reference internal/box_helpers.py:59-106 (slab test; intersection is int32)."""
    inv_d = torch.reciprocal(ray_d)
    t_min = (aabb_min - ray_o) * inv_d
    t_max = (aabb_max - ray_o) * inv_d
    t0 = torch.minimum(t_min, t_max)
    t1 = torch.maximum(t_min, t_max)
    t_near = torch.maximum(torch.maximum(t0[..., 0], t0[..., 1]), t0[..., 2])
    t_far = torch.minimum(torch.minimum(t1[..., 0], t1[..., 1]), t1[..., 2])
    hit = (t_far > t_near).to(torch.int32)
    positive_far = ((t_far * hit) > 0).to(torch.int32)
    hit = hit * positive_far
    return t_near * hit, t_far * hit, hit


# --------------------------------------------------------------------------------------
# internal/obbpose_model.py : MLP / BoxMLP
# --------------------------------------------------------------------------------------

class MLPTopology(NamedTuple):
    """Shape descriptor of reference MLP (obbpose_model.py:294-354) / BoxMLP (:358-418)."""
    in_dim: int = 60
    width: int = 256
    depth: int = 8
    skip: int = 4
    cond_dim: int = 27
    cond_width: int = 128

    def layer_shapes(self) -> List[Tuple[int, int]]:
        """[in, out] of Dense_0 .. Dense_{depth+3}, in flax creation order
        (trunk, density, bottleneck, condition, rgb)."""
        shapes = []
        k = self.in_dim
        for i in range(self.depth):
            shapes.append((k, self.width))
            k = self.width
            if i % self.skip == 0 and i > 0:
                k = self.width + self.in_dim
        shapes.append((k, 1))
        shapes.append((k, self.width))
        shapes.append((self.width + self.cond_dim, self.cond_width))
        shapes.append((self.cond_width, 3))
        return shapes

    def num_params(self) -> int:
        return sum(i * o + o for i, o in self.layer_shapes())


BG_TOPOLOGY = MLPTopology(60, 256, 8, 4, 27, 128)     # configs/carla_dyn.gin:55-58
BOX_TOPOLOGY = MLPTopology(63, 128, 8, 4, 27, 128)    # BoxMLP defaults, obbpose_model.py:360-363


def init_mlp_params(topo: MLPTopology, rng: np.random.Generator, dtype=torch.float32,
                    bias_scale: float = 0.0) -> List[Tuple[Tensor, Tensor]]:
    """glorot-uniform kernels U(+-sqrt(6/(fan_in+fan_out))), zero biases
    (obbpose_model.py:326-327; flax Dense default bias init).  `bias_scale` > 0 draws small
    non-zero biases so that tests exercise the bias path."""
    out = []
    for fi, fo in topo.layer_shapes():
        lim = _pymath.sqrt(6.0 / (fi + fo))
        k = rng.uniform(-lim, lim, size=(fi, fo)).astype(np.float32)
        b = (rng.uniform(-1, 1, size=(fo,)) * bias_scale).astype(np.float32)
        out.append((torch.from_numpy(k).to(dtype), torch.from_numpy(b).to(dtype)))
    return out


def mlp_apply(params: Sequence[Tuple[Tensor, Tensor]], topo: MLPTopology, x: Tensor,
              condition: Optional[Tensor]):
    """reference internal/obbpose_model.py:305-354 (MLP) == :369-418 (BoxMLP).
    x [B,N,F], condition [B,C] -> raw_rgb [B,N,3], raw_density [B,N,1]."""
    B, N, F = x.shape
    x = x.reshape(-1, F)
    inputs = x
    li = 0
    for i in range(topo.depth):
        k, b = params[li]; li += 1
        x = torch.relu(x @ k + b)
        if i % topo.skip == 0 and i > 0:
            x = torch.cat([x, inputs], dim=-1)
    k, b = params[li]; li += 1
    raw_density = (x @ k + b).reshape(-1, N, 1)
    if condition is not None:
        k, b = params[li]; li += 1
        bottleneck = x @ k + b
        cond = condition[:, None, :].expand(B, N, condition.shape[-1]).reshape(-1, condition.shape[-1])
        x = torch.cat([bottleneck, cond], dim=-1)
        k, b = params[li]; li += 1
        x = torch.relu(x @ k + b)
    else:
        li += 2
    k, b = params[li]
    raw_rgb = (x @ k + b).reshape(-1, N, 3)
    return raw_rgb, raw_density


# --------------------------------------------------------------------------------------
# internal/obbpose_model.py : MipNerfModel.__call__
# --------------------------------------------------------------------------------------

class Rays(NamedTuple):
    """reference internal/utils.py:84-86 (BoxRays)."""
    origins: Tensor
    directions: Tensor
    viewdirs: Tensor
    radii: Tensor
    lossmult: Tensor
    near: Tensor
    far: Tensor


class ModelConfig(NamedTuple):
    """reference MipNerfModel fields (obbpose_model.py:45-66) with configs/carla_dyn.gin values."""
    num_samples: int = 128
    num_levels: int = 2
    resample_padding: float = 0.01
    stop_level_grad: bool = True
    use_viewdirs: bool = True
    lindisp: bool = False
    ray_shape: str = 'cone'
    min_deg_point: int = 0
    max_deg_point: int = 10
    deg_view: int = 4
    density_noise: float = 0.0
    density_bias: float = -1.0
    disable_integration: bool = False
    contraction: bool = True
    dynamics: bool = True
    no_pose_opt: bool = True
    no_yaw_opt: bool = True


class LevelOut(NamedTuple):
    """the 10-tuple appended per level at obbpose_model.py:258-260."""
    comp_rgb: Tensor
    distance: Tensor
    acc: Tensor
    weights: Tensor
    t_vals: Tensor
    t_mids: Tensor
    t_dists: Tensor
    off: Tuple[Tensor, Tensor]
    dyn_mask: Tensor
    zo: Tensor


def model_forward(params: Dict, rays: Rays, ext: Tensor, ts: int, randomized: bool, rand_bkgd: bool,
                  white_bkgd: bool, alpha: float, cfg: ModelConfig = ModelConfig(),
                  t_rand: Optional[Tensor] = None, u_rand: Optional[Tensor] = None,
                  density_noise: Optional[Sequence[Tensor]] = None,
                  bg_topo: MLPTopology = BG_TOPOLOGY, box_topo: MLPTopology = BOX_TOPOLOGY,
                  keep_raw: Optional[list] = None) -> List[LevelOut]:
    """reference internal/obbpose_model.py:69-261.

    params = {'mlp': [(k,b)...], 'box_mlps': [[(k,b)...] per object], 'box_centers': [T,K,6]}.
    """
    pose_offsets = params['box_centers']
    if pose_offsets.dim() < 3:
        pose_offsets = pose_offsets[:, None, :]          # init_boxes, obbpose_model.py:35-39
    K = pose_offsets.shape[1]
    origins, dirs = rays.origins, rays.directions
    B = origins.shape[0]

    box_pose = pose_offsets[ts, :, :3].expand(B, K, 3)
    if cfg.no_pose_opt:
        box_pose = box_pose.detach()
    box_rot = pose_offsets[ts, :, 3:]
    if cfg.no_yaw_opt:
        box_rot = box_rot.detach()
    box_mat = aa2matrix(box_rot).expand(B, K, 3, 3)
    box_dims = ext.expand(B, K, 3)

    origins_o, dirs_o = world2object_rpy(origins, dirs, box_pose, box_mat)
    zi, zo, hit = ray_box_intersection(origins_o, dirs_o, -box_dims, box_dims)
    hit = hit.detach()
    hitf = hit.to(origins.dtype)

    bkgd_mask = (hit.sum(dim=-1) == 0).to(origins.dtype)
    origins_s = (origins_o * hitf[..., None]).sum(dim=-2) + bkgd_mask[..., None] * origins
    dirs_s = (dirs_o * hitf[..., None]).sum(dim=-2) + bkgd_mask[..., None] * dirs
    zo_ret = (hitf * zo).sum(dim=-1)
    # near/far box clipping (obbpose_model.py:126-129) is computed then unused by the reference.

    viewdirs_enc = pos_enc(rays.viewdirs, 0, cfg.deg_view, True) if cfg.use_viewdirs else None

    ret = []
    t_vals = None
    weights = None
    for i_level in range(cfg.num_levels):
        if i_level == 0:
            t_vals, samples = sample_along_rays(origins_s, dirs_s, rays.radii, cfg.num_samples,
                                                rays.near, rays.far, randomized, cfg.lindisp,
                                                cfg.ray_shape, t_rand=t_rand)
        else:
            t_vals, samples = resample_along_rays(origins_s, dirs_s, rays.radii, t_vals, weights,
                                                  randomized, cfg.ray_shape, cfg.stop_level_grad,
                                                  cfg.resample_padding, u_rand=u_rand)
        if cfg.disable_integration:
            samples = (samples[0], torch.zeros_like(samples[1]))

        raw_rgbs = 0.0
        raw_densities = 0.0
        ret_masks = []
        if cfg.dynamics:
            Bs, N, _ = samples[0].shape
            masks_sum = 0.0
            for k in range(K):
                mask = hitf[:, k].reshape(-1, 1)
                ret_masks.append(mask)
                mask3 = mask[:, None, :].expand(Bs, N, 1)
                obj_samples = (mask3 * samples[0], mask3[..., None] * samples[1])
                enc = weighted_ipe(obj_samples, cfg.min_deg_point, cfg.max_deg_point, alpha)
                o_rgb, o_den = mlp_apply(params['box_mlps'][k], box_topo, enc, viewdirs_enc)
                raw_rgbs = raw_rgbs + mask3 * o_rgb
                raw_densities = raw_densities + mask3 * o_den
                masks_sum = masks_sum + mask3
            bm = (1 - masks_sum).detach()
            samples = (bm * samples[0], bm[..., None] * samples[1])

        if cfg.contraction:
            samples = new_space(samples)
        samples_enc = integrated_pos_enc(samples, cfg.min_deg_point, cfg.max_deg_point)
        raw_rgb, raw_density = mlp_apply(params['mlp'], bg_topo, samples_enc, viewdirs_enc)
        if cfg.dynamics:
            raw_rgb = raw_rgb + raw_rgbs
            raw_density = raw_density + raw_densities
        if randomized and cfg.density_noise > 0:
            raw_density = raw_density + cfg.density_noise * density_noise[i_level]
        if keep_raw is not None:
            keep_raw.append((raw_rgb, raw_density, samples_enc))

        rgb = torch.sigmoid(raw_rgb)
        density = torch.nn.functional.softplus(raw_density + cfg.density_bias)
        comp_rgb, distance, acc, weights, t_vals, t_mids, t_dists = volumetric_rendering(
            rgb, density, t_vals, dirs_s, white_bkgd=white_bkgd, rand_bkgd=rand_bkgd)
        if cfg.dynamics:
            dyn = torch.stack(ret_masks, 0).sum(dim=0)
        else:
            dyn = hitf.sum(dim=-1)[..., None]
        ret.append(LevelOut(comp_rgb, distance, acc, weights, t_vals, t_mids, t_dists,
                            (box_pose[0], box_rot[0]), dyn, zo_ret))       # box_rot is [K,3]: the reference returns object 0's rotation
    return ret


def render_image(render_fn, rays: Rays, chunk: int = 8192):
    """reference internal/obbpose_model.py:421-479: python loop over `chunk`-ray slices of the
    flattened frame, keep the fine level, reshape to (H, W).  `render_fn(chunk_rays)` returns the
    list of LevelOut."""
    height, width = rays.origins.shape[:2]
    num_rays = height * width
    flat = Rays(*[r.reshape(num_rays, -1) for r in rays])
    rgbs, dists, accs = [], [], []
    for i in range(0, num_rays, chunk):
        out = render_fn(Rays(*[r[i:i + chunk] for r in flat]))[-1]
        rgbs.append(out.comp_rgb); dists.append(out.distance); accs.append(out.acc)
    return (torch.cat(rgbs, 0).reshape(height, width, -1), torch.cat(dists, 0).reshape(height, width),
            torch.cat(accs, 0).reshape(height, width))


# --------------------------------------------------------------------------------------
# train_boxpose.py : loss block, gradient post-processing, Adam
# --------------------------------------------------------------------------------------

class LossConfig(NamedTuple):
    """reference internal/utils.py Config fields used by the loss, configs/carla_dyn.gin values."""
    coarse_loss_mult: float = 0.1
    box_loss_mult: float = 0.0
    tv_loss_mult: float = 0.0
    depth_loss_mult: float = 0.0001
    near_loss_mult: float = 0.01
    empty_loss_mult: float = 1.0
    sky_loss_mult: float = 1.0
    weight_decay_mult: float = 0.0
    disable_multiscale_loss: bool = False
    grad_max_norm: float = 1.0
    grad_max_val: float = 0.1


def loss_fn(ret: List[LevelOut], rays: Rays, pixels: Tensor, depth_gt: Tensor, sky: Tensor,
            eps: float, cfg: LossConfig = LossConfig(), prev: Optional[Tensor] = None,
            param_tensors: Optional[Sequence[Tensor]] = None):
    """reference train_boxpose.py:94-220.  depth_gt, sky: [B,1]; pixels [B,3].
    Returns (loss, stats dict)."""
    dt = pixels.dtype
    mask = rays.lossmult
    if cfg.disable_multiscale_loss:
        mask = torch.ones_like(mask)
    if param_tensors is not None and cfg.weight_decay_mult != 0.0:
        weight_l2 = cfg.weight_decay_mult * (sum((p ** 2).sum() for p in param_tensors) /
                                             sum(p.numel() for p in param_tensors))
    else:
        weight_l2 = torch.zeros((), dtype=dt)

    z = depth_gt.squeeze(-1)
    depth_mask = (z > 0.0).to(dt)
    sky_mask = (sky.squeeze(-1) > 0.0).to(dt)
    sky_mask = sky_mask - depth_mask * sky_mask

    losses, d_losses, distr_losses, tv_losses, s_losses, e_losses, n_losses, obj_losses = ([] for _ in range(8))
    for lv in ret:
        pose = lv.off[0]
        if prev is not None:
            tv_losses.append(((pose - prev[:, :, :3]) ** 2).sum())
        else:
            tv_losses.append(torch.zeros((), dtype=dt))
        box_mask = (z < lv.zo).to(dt)
        depth_mask = depth_mask + cfg.box_loss_mult * lv.dyn_mask.squeeze(-1) * box_mask  # accumulates (:140)

        tvals = lv.t_vals[:, :-1]
        w = lv.weights
        s = lv.t_mids
        Sij = torch.abs(s[:, :, None] - s[:, None, :])                                 # (:146-150)
        term1 = (w[:, :, None] * w[:, None, :] * Sij).sum()
        term2 = (1 / 3) * (w ** 2 * lv.t_dists).sum()
        distr_losses.append(term1 + term2)

        depth_t = depth_gt.expand_as(tvals)
        sigma = (eps / 3.0) ** 2
        mask_near = ((tvals > (depth_t - eps)) & (tvals < (depth_t + eps))).to(dt)
        mask_near = mask_near * depth_mask.reshape(tvals.shape[0], -1)
        mask_empty = (tvals > (depth_t + eps)).to(dt)
        mask_empty = mask_empty * depth_mask.reshape(tvals.shape[0], -1)
        dist = mask_near * (tvals - depth_t)
        distr = 1.0 / (sigma * _pymath.sqrt(2 * _pymath.pi)) * torch.exp(-(dist ** 2 / (2 * sigma ** 2)))
        distr = distr / distr.max()
        distr = distr * mask_near
        norm = torch.clamp(depth_mask.sum(), min=1.0)
        n_losses.append(((mask_near * w - distr) ** 2).sum() / norm)
        e_losses.append(((mask_empty * w) ** 2).sum() / norm)
        d_losses.append((depth_mask * (lv.distance - z) ** 2).sum() / norm)

        sky_depth = sky_mask * (1.0 - (1.0 / torch.clamp(sky_mask * lv.distance, min=1.0)))
        s_losses.append((sky_mask * (sky_depth - sky.squeeze(-1)) ** 2).sum() / torch.clamp(sky_mask.sum(), min=1.0))

        rgb_w = mask + cfg.box_loss_mult * lv.dyn_mask * box_mask[..., None]
        sq = (lv.comp_rgb - pixels[..., :3]) ** 2
        losses.append((rgb_w * sq).sum() / mask.sum())
        obj_losses.append((lv.dyn_mask * sq).sum() / lv.dyn_mask.sum())

    st = lambda xs: torch.stack(xs)
    losses, d_losses, distr_losses, tv_losses = st(losses), st(d_losses), st(distr_losses), st(tv_losses)
    n_losses, e_losses, s_losses, obj_losses = st(n_losses), st(e_losses), st(s_losses), st(obj_losses)

    loss = cfg.coarse_loss_mult * losses[:-1].sum() + losses[-1] + weight_l2
    loss = loss + cfg.sky_loss_mult * s_losses[:-1].sum() + 10.0 * cfg.sky_loss_mult * s_losses[-1]
    loss = loss + cfg.depth_loss_mult * d_losses[-1] + 0.1 * cfg.depth_loss_mult * d_losses[:-1].sum()
    loss = loss + cfg.near_loss_mult * n_losses[-1] + 0.1 * cfg.near_loss_mult * n_losses[:-1].sum()
    loss = loss + cfg.empty_loss_mult * e_losses[-1] + 0.1 * cfg.empty_loss_mult * e_losses[:-1].sum()
    loss = loss + cfg.tv_loss_mult * tv_losses[-1] + 0.1 * cfg.tv_loss_mult * tv_losses[:-1].sum()
    loss = loss + 0.000001 * distr_losses[-1] + 0.000001 * distr_losses[:-1].sum()
    stats = dict(loss=loss, losses=losses, d_losses=d_losses, n_losses=n_losses, e_losses=e_losses,
                 s_losses=s_losses, distr_losses=distr_losses, tv_losses=tv_losses, obj_losses=obj_losses,
                 weight_l2=weight_l2)
    return loss, stats


def postprocess_grads(grads: Sequence[Tensor], cfg: LossConfig = LossConfig()):
    """reference train_boxpose.py:262-286: nan_to_num(posinf=0) -> clip value -> global-norm clip."""
    gs = [torch.nan_to_num(g, nan=0.0, posinf=0.0) for g in grads]
    if cfg.grad_max_val > 0:
        gs = [torch.clamp(g, -cfg.grad_max_val, cfg.grad_max_val) for g in gs]
    norm = torch.sqrt(sum((g ** 2).sum() for g in gs))
    if cfg.grad_max_norm > 0:
        mult = torch.clamp(cfg.grad_max_norm / (1e-7 + norm), max=1.0)
        gs = [mult * g for g in gs]
    return gs, norm


def adam_step(params: Sequence[Tensor], grads: Sequence[Tensor], m: Sequence[Tensor], v: Sequence[Tensor],
              step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8):
    """flax.optim.Adam.apply_param_gradient (flax<=0.3, pinned by requirements_jax.txt:4; call site
    train_boxpose.py:288,343), weight_decay = 0:
      m' = (1-b1) g + b1 m ; v' = (1-b2) g^2 + b2 v ; t = step+1
      p' = p - lr * (m'/(1-b1^t)) / (sqrt(v'/(1-b2^t)) + eps)."""
    t = step + 1.0
    outp, outm, outv = [], [], []
    for p, g, mi, vi in zip(params, grads, m, v):
        m2 = (1.0 - beta1) * g + beta1 * mi
        v2 = (1.0 - beta2) * (g * g) + beta2 * vi       # flax: (1. - beta2) * lax.square(grad)
        mhat = m2 / (1.0 - beta1 ** t)
        denom = torch.sqrt(v2 / (1.0 - beta2 ** t)) + eps
        outp.append(p - lr * mhat / denom); outm.append(m2); outv.append(v2)
    return outp, outm, outv


# --------------------------------------------------------------------------------------
# internal/obbpose_dataset.py : pinhole ray generation (numpy on the host in the reference)
# --------------------------------------------------------------------------------------

def generate_rays(c2w: np.ndarray, w: int, h: int, focal: float, near: float, far: float, principal_point=None):
    """reference internal/obbpose_dataset.py:613-661 (`_generate_rays_multi`) for one camera, in float32 like the
    reference's arrays (numpy 1.x value-based casting keeps `v * 2 / np.sqrt(12)` in float32).
    `principal_point` = (cx, cy): the Waymo loader's variant (:1868-1917), which differs only in the pixel offset.
    Returns Rays of numpy arrays shaped [h, w, 3] / [h, w, 1]."""
    f32 = np.float32
    c2w = np.asarray(c2w, f32)
    x, y = np.meshgrid(np.arange(w, dtype=f32), np.arange(h, dtype=f32), indexing='xy')
    pcx, pcy = (f32(w) * f32(0.5), f32(h) * f32(0.5)) if principal_point is None else (f32(principal_point[0]), f32(principal_point[1]))
    cam_dirs = np.stack([(x - pcx) / f32(focal), -(y - pcy) / f32(focal), -np.ones_like(x)], axis=-1)
    directions = (cam_dirs[..., None, :] * c2w[:3, :3]).sum(axis=-1).astype(f32)
    origins = np.broadcast_to(c2w[:3, -1], directions.shape).astype(f32)
    viewdirs = (directions / np.linalg.norm(directions, axis=-1, keepdims=True)).astype(f32)
    dx = np.sqrt(np.sum((directions[:-1, :, :] - directions[1:, :, :]) ** 2, -1))
    dx = np.concatenate([dx, dx[-2:-1, :]], 0)
    radii = ((dx[..., None] * f32(2)) / f32(np.sqrt(12))).astype(f32)
    ones = np.ones_like(origins[..., :1])
    return Rays(origins, directions, viewdirs, radii, ones, f32(near) * ones, f32(far) * ones)
