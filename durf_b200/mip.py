"""Mirror of the reference's internal/mip.py call surface on top of the CUDA kernels.

`sample_along_rays`, `resample_along_rays` and `cast_rays` return a `Samples` handle instead of materialised
(means, covs) arrays; `mip360.new_space(samples)` and a multiplication by a per-ray mask only tag the handle, and
`integrated_pos_enc` / `weighted_ipe` launch ONE fused kernel (frustum -> Gaussian -> mask -> contraction -> encoding)
so the 48 B/sample Gaussians never travel through HBM.  `samples.gaussians()` materialises (means, cov_diag) when a
caller really wants them (only the covariance diagonal reaches the encoding: mip.py:273-278).

`key` arguments take the explicit random buffer (a tensor of U[0,1) draws) or None: the reference's threefry stream
cannot be reproduced without JAX.
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Optional

import torch

from . import ops


@dataclass
class Samples:
    t_vals: torch.Tensor
    origins: torch.Tensor
    directions: torch.Tensor
    radii: torch.Tensor
    ray_shape: str = 'cone'
    ray_mult: Optional[torch.Tensor] = None     # `mask * samples` (obbpose_model.py:179-180, 207-208)
    contracted: bool = False                    # mip360.new_space applied
    integrate: bool = True                      # False: disable_integration (covariances zeroed)

    def __getitem__(self, i):                   # samples[0] / samples[1] like the reference's tuple
        g = self.gaussians()
        return g[0] if i == 0 else g[1]

    def masked(self, mult: torch.Tensor) -> "Samples":
        m = ops.f32(mult).reshape(-1)
        return replace(self, ray_mult=m if self.ray_mult is None else self.ray_mult * m)

    def _encode(self, **kw):
        N = self.t_vals.shape[-1] - 1
        return ops.raymarch(self.origins, self.directions, self.radii, N, t_vals=self.t_vals, contract=self.contracted,
                            ray_shape=self.ray_shape, integrate=self.integrate, ray_mult=self.ray_mult, **kw)

    def gaussians(self):
        out = self._encode(want_gaussians=True)
        return out['means'], out['cov_diag']


def pos_enc(x, min_deg, max_deg, append_identity=True):
    """mip.py:36-45 (the model calls it with min_deg=0 and append_identity=True on the view directions)."""
    if min_deg != 0 or not append_identity:
        raise NotImplementedError("pos_enc kernel covers min_deg=0, append_identity=True (the model's only use)")
    return ops.viewdir_enc(x, max_deg)


def cast_rays(t_vals, origins, directions, radii, ray_shape, diag=False) -> Samples:
    """mip.py:155-179."""
    if ray_shape not in ('cone', 'cylinder'):
        raise AssertionError("ray_shape must be 'cone' or 'cylinder'")
    return Samples(ops.f32(t_vals), origins, directions, radii, ray_shape)


def sample_along_rays(key, origins, directions, radii, num_samples, near, far, randomized, lindisp, ray_shape):
    """mip.py:330-370 -> (t_vals, samples).  `key`: t_rand [B,N+1] or None (drawn on device when randomized)."""
    if lindisp:
        raise NotImplementedError("lindisp=True is not built (both gin files set lindisp=False; the reference's formula "
                                  "at mip.py:354-356 is also wrong)")
    if ray_shape not in ('cone', 'cylinder'):
        raise AssertionError("ray_shape must be 'cone' or 'cylinder'")
    B = origins.shape[0]
    t_rand = None
    if randomized:
        t_rand = key if torch.is_tensor(key) else torch.rand(B, num_samples + 1, device=origins.device)
    # the fenceposts come out of the same fused kernel; features are not needed here -> cheapest variant
    out = ops.raymarch(origins, directions, radii, num_samples, near=near, far=far, t_rand=t_rand, ray_shape=ray_shape,
                       min_deg=0, max_deg=1)
    return out['t_vals'], Samples(out['t_vals'], origins, directions, radii, ray_shape)


def resample_along_rays(key, origins, directions, radii, t_vals, weights, randomized, ray_shape, stop_grad,
                        resample_padding):
    """mip.py:373-416 -> (new_t_vals, samples).  The result carries no gradient (stop_grad is always honoured)."""
    u_rand = None
    if randomized:
        u_rand = key if torch.is_tensor(key) else torch.rand(t_vals.shape[0], t_vals.shape[-1], device=t_vals.device)
    new_t = ops.resample(t_vals, weights, u_rand=u_rand, padding=resample_padding)
    return new_t, Samples(new_t, origins, directions, radii, ray_shape)


def integrated_pos_enc(x_coord: Samples, min_deg, max_deg, diag=False):
    """mip.py:226-282 -> [B,N,6*(max_deg-min_deg)]."""
    return x_coord._encode(min_deg=min_deg, max_deg=max_deg)['features']


def weighted_ipe(x_coord: Samples, min_deg, max_deg, alpha, diag=False):
    """mip.py:182-223 -> [B,N,3+6*(max_deg-min_deg)] with the reference's i//6 weight layout."""
    return x_coord._encode(min_deg=min_deg, max_deg=max_deg, weighted=True, alpha=float(alpha))['features']


def volumetric_rendering(rgb, density, t_vals, dirs, white_bkgd, rand_bkgd, key=None):
    """mip.py:285-327 (inputs already activated) -> (comp_rgb, depth, acc, weights, t_vals, t_mids, t_dists)."""
    o = ops.composite(rgb, density, t_vals, dirs, white_bkgd=white_bkgd, rand_bkgd=rand_bkgd, activated=True)
    return o['comp_rgb'], o['depth'], o['acc'], o['weights'], t_vals, o['t_mids'], o['t_dists']
