"""Host-side mirror of the reference's train_step (train_boxpose.py:49-321): forward through the model, the loss
block (94-220) and its gradient, gradient mean over ranks (pmean, :253), nan_to_num / clip / global-norm clip
(262-286) and flax.optim.Adam (288).  Every array operation is a kernel of libdurf_b200.so; torch.distributed (NCCL)
provides the one gradient all-reduce."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops
from . import parallel
from .obbpose_model import MipNerfModel, Variables
from .utils import Config


@dataclass
class TrainState:
    """utils.TrainState(optimizer) of the reference: parameters + Adam moments + step counter."""
    variables: Variables
    m: torch.Tensor
    v: torch.Tensor
    step: int = 0

    @staticmethod
    def create(variables: Variables) -> "TrainState":
        return TrainState(variables, torch.zeros_like(variables.flat), torch.zeros_like(variables.flat), 0)


def loss_and_grads(model: MipNerfModel, config: Config, ret, batch: Dict[str, torch.Tensor], eps: float):
    """train_boxpose.py:94-220 for every level: returns (stats dict of device scalars, per-level gradient dicts).
    stats['loss'] is the total loss; the per-term arrays keep the reference's Stats names."""
    dev = ret[0][0].device
    B, N = ret[0][3].shape
    nl = len(ret)
    partials = torch.zeros(nl * L.LP_STRIDE, device=dev)
    norms = torch.zeros(nl, 4, device=dev)
    depth_mask = torch.empty(B, device=dev)
    lb = dict(pixels=ops.f32(batch['pixels'])[..., :3].contiguous(), depth=ops.f32(batch['depth']).reshape(-1),
              sky=ops.f32(batch['sky']).reshape(-1), lossmult=ops.f32(batch['rays'].lossmult).reshape(-1) if not
              config.disable_multiscale_loss else torch.ones(B, device=dev),
              dyn_mask=ret[0][8].reshape(-1).contiguous(), zo=ret[0][9].contiguous())
    grads = []
    for i, lv in enumerate(ret):
        lvd = dict(comp_rgb=lv[0], depth=lv[1], weights=lv[3], t_vals=lv[4])
        g = dict(comp_rgb=torch.empty(B, 3, device=dev), depth=torch.empty(B, device=dev), weights=torch.empty(B, N, device=dev))
        a = ops.loss_args(i, nl, eps, config, lvd, lb, depth_mask, partials, g)
        ops.losses_prepare(a, norms[i])
        ops.losses_fwd_bwd(a, norms[i])
        grads.append(g)
    # normalise the partial sums into the reference's per-level loss arrays (tiny device-side scalar math)
    p = partials.view(nl, L.LP_STRIDE)
    nd = torch.clamp(norms[:, 1], min=1.0)
    losses = p[:, 0] / norms[:, 0]
    d_losses, n_losses, e_losses = p[:, 1] / nd, p[:, 2] / nd, p[:, 3] / nd
    s_losses = p[:, 4] / torch.clamp(norms[:, 2], min=1.0)
    distr = p[:, 5]
    c = config
    loss = (c.coarse_loss_mult * losses[:-1].sum() + losses[-1]
            + c.sky_loss_mult * s_losses[:-1].sum() + 10.0 * c.sky_loss_mult * s_losses[-1]
            + c.depth_loss_mult * d_losses[-1] + 0.1 * c.depth_loss_mult * d_losses[:-1].sum()
            + c.near_loss_mult * n_losses[-1] + 0.1 * c.near_loss_mult * n_losses[:-1].sum()
            + c.empty_loss_mult * e_losses[-1] + 0.1 * c.empty_loss_mult * e_losses[:-1].sum()
            + 0.000001 * distr[-1] + 0.000001 * distr[:-1].sum())
    stats = dict(loss=loss, losses=losses, d_losses=d_losses, n_losses=n_losses, e_losses=e_losses, s_losses=s_losses,
                 distr_losses=distr)
    return stats, grads


def train_step(model: MipNerfModel, config: Config, rng, state: TrainState, batch: Dict, lr: float, eps: float, alpha: float,
               prev: Optional[torch.Tensor] = None, world_size: int = 1):
    """train_boxpose.py:49-321.  batch = dict(rays, init, ext, ts, pixels, depth, sky).  Returns (state, stats)."""
    v = state.variables
    ctx: dict = {}
    ret = model.apply(v, rng, batch['rays'], batch.get('init'), batch['ext'], batch['ts'], randomized=config.randomized,
                      rand_bkgd=config.rand_bkgd, white_bkgd=config.white_bkgd, alpha=alpha, ctx=ctx)
    stats, grads = loss_and_grads(model, config, ret, batch, eps)
    d_flat = torch.zeros_like(v.flat)
    model.backward(v, ctx, grads, d_flat)
    if config.tv_loss_mult != 0.0 and prev is not None:
        # tv_losses (train_boxpose.py:136, 219): (pose - prev)^2 per level, weights 1 (fine) + 0.1 (coarse)
        ts = ctx['ts']
        w = config.tv_loss_mult * (1.0 + 0.1 * (len(ret) - 1))
        if not model.no_pose_opt:
            v.view_of(d_flat, 'box_centers')[ts, :, :3] += 2.0 * w * (v.box_centers[ts, :, :3] - prev.reshape(-1, 3)[: v.K])
    scale = parallel.allreduce_gradients(d_flat) if world_size > 1 else 1.0   # jax.lax.pmean(grad, 'batch'), :253
    sumsq = torch.zeros(1, device=d_flat.device)
    ops.grad_sanitize(d_flat, config.grad_max_val, scale, sumsq)
    ops.adam_step(v.flat, d_flat, state.m, state.v, sumsq, max_norm=config.grad_max_norm, lr=lr, step=state.step)
    v.mark_dirty()
    state.step += 1
    stats['grad_norm'] = torch.sqrt(sumsq[0])
    stats['grad'] = d_flat
    return state, stats
