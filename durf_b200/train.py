"""Host-side mirror of the reference's train_step (train_boxpose.py:49-321): forward through the model, the loss
block (94-220) and its gradient, gradient mean over ranks (pmean, :253), nan_to_num / clip / global-norm clip
(262-286) and flax.optim.Adam (288).  Every array operation is a kernel of libdurf_b200.so; torch.distributed (NCCL)
provides the gradient all-reduce.

The step makes no host synchronisation and reads its per-step scalars (learning rate, eps, alpha, timestep, step
counter) either from Python numbers or from device memory (`StepScalars`), so a whole step can be captured once in a
CUDA graph and replayed (`GraphedTrainStep`): at the reference's shipped batch of 512 rays the eager step is bound by
~100 kernel launches, not by the GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops
from . import parallel
from .obbpose_model import MipNerfModel, Variables
from .utils import Config, Rays


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


@dataclass
class StepScalars:
    """Per-step scalars in device memory: what changes from one replay of a captured step to the next."""
    f: torch.Tensor          # float32[4]: lr, eps, alpha, and the bits of `step`
    step: torch.Tensor       # int32[1] view of f[3]: 0-based optimizer step
    ts: torch.Tensor         # int64[1]: timestep row of box_centers
    # True: the Adam launch increments `step` (one more one-thread kernel per step); False: the host uploads it with the
    # other scalars before every replay (GraphedTrainStep does, in the same 16-byte copy)
    advance_on_device: bool = True

    @staticmethod
    def create(device, step: int = 0, advance_on_device: bool = True) -> "StepScalars":
        f = torch.zeros(4, device=device)
        s = f[3:4].view(torch.int32)
        s.fill_(step)
        return StepScalars(f, s, torch.zeros(1, device=device, dtype=torch.int64), advance_on_device)

    lr = property(lambda self: self.f[0:1])
    eps = property(lambda self: self.f[1:2])
    alpha = property(lambda self: self.f[2:3])


_STAT_NAMES = ('losses', 'd_losses', 'n_losses', 'e_losses', 's_losses', 'distr_losses', 'obj_losses', 'tv_losses')


def loss_and_grads(model: MipNerfModel, config: Config, ret, batch: Dict[str, torch.Tensor], eps, tv=None, weight_l2=None,
                   workspace: Optional[dict] = None, norms: Optional[torch.Tensor] = None):
    """train_boxpose.py:94-220 for every level: returns (stats dict of device scalars, per-level gradient dicts).
    stats['loss'] is the total loss (tv and weight_l2 terms included); the per-term arrays keep the reference's names."""
    dev = ret[0][0].device
    B, N = ret[0][3].shape
    nl = len(ret)
    ws = workspace if workspace is not None else {}
    if 'reduce_ws' not in ws or ws.get('reduce_B') != B:
        ws['reduce_ws'], ws['reduce_B'] = ops.losses_reduce_ws(B, dev), B
    partials = torch.empty(nl * L.LP_STRIDE, device=dev)
    if norms is None:                   # [nl,4] zeros (train_step hands in a slice of the buffer it zeroes once per step)
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
        a = ops.loss_args(i, nl, eps, config, lvd, lb, depth_mask, partials, g, ws['reduce_ws'])
        ops.losses_prepare(a, norms[i])
        ops.losses_fwd_bwd(a, norms[i])
        grads.append(g)
    out = torch.empty(nl * L.LS_STRIDE + 2, device=dev)
    fa = L.LossFinalizeArgs(num_levels=nl, coarse_loss_mult=config.coarse_loss_mult, depth_loss_mult=config.depth_loss_mult,
                            near_loss_mult=config.near_loss_mult, empty_loss_mult=config.empty_loss_mult,
                            sky_loss_mult=config.sky_loss_mult, tv_loss_mult=config.tv_loss_mult, distortion_mult=1e-6,
                            partials=L.ptr(partials), norms=L.ptr(norms), tv=L.ptr(tv), weight_l2=L.ptr(weight_l2), stats=L.ptr(out))
    L.check(L.load().durf_losses_finalize(L.stream_ptr(), C.byref(fa)), "durf_losses_finalize")
    per = out[:nl * L.LS_STRIDE].view(nl, L.LS_STRIDE)
    stats = Stats({name: per[:, i] for i, name in enumerate(_STAT_NAMES)})
    stats['loss'] = out[nl * L.LS_STRIDE]
    stats['weight_l2'] = out[nl * L.LS_STRIDE + 1]
    return stats, grads


def train_step(model: MipNerfModel, config: Config, rng, state: TrainState, batch: Dict, lr, eps, alpha,
               prev: Optional[torch.Tensor] = None, world_size: int = 1, scalars: Optional[StepScalars] = None,
               workspace: Optional[dict] = None):
    """train_boxpose.py:49-321.  batch = dict(rays, init, ext, ts, pixels, depth, sky).  Returns (state, stats).

    With `scalars` the learning rate, eps, alpha, timestep and step counter are read from device memory (the Python
    values are ignored) and nothing in the step depends on a host value that changes between steps."""
    v = state.variables
    ctx: dict = {}
    dev_mode = scalars is not None
    ts = scalars.ts if dev_mode else batch['ts']
    alpha_ = scalars.alpha if dev_mode else alpha
    eps_ = scalars.eps if dev_mode else eps
    ret = model.apply(v, rng, batch['rays'], batch.get('init'), batch['ext'], ts, randomized=config.randomized,
                      rand_bkgd=config.rand_bkgd, white_bkgd=config.white_bkgd, alpha=alpha_, ctx=ctx)
    # one zero-fill per step: the flat gradient, the squared gradient norm and the loss kernels' normalisers
    n_flat, nl = v.flat.numel(), len(ret)
    n_pad = (n_flat + 3) // 4 * 4
    zeros = torch.zeros(n_pad + 4 + 4 * nl, device=v.flat.device)
    d_flat, sumsq, norms = zeros[:n_flat], zeros[n_pad:n_pad + 1], zeros[n_pad + 4:].view(nl, 4)
    tv = wl2 = None
    pose = ret[0][7][0]                                                       # box_pose[0] of this timestep, [K,3]
    if prev is not None:
        # tv_losses (train_boxpose.py:136, 219): sum (pose - prev[:, :, :3])^2, the same value for every level (reported
        # even when tv_loss_mult = 0); weights: fine x1, coarse x0.1
        diff = pose - ops.f32(prev).reshape(-1, prev.shape[-1])[: v.K, :3]
        tv = (diff * diff).sum().expand(len(ret)).contiguous()
        if config.tv_loss_mult != 0.0 and not model.no_pose_opt:
            w = config.tv_loss_mult * (1.0 + 0.1 * (len(ret) - 1))
            _add_box_grad(v, d_flat, ctx, 2.0 * w * diff, cols=slice(0, 3))
    if config.weight_decay_mult != 0.0:
        # weight_l2 = mult * sum(p^2) / count over the whole parameter tree (train_boxpose.py:67-74)
        n = v.flat.numel()
        wl2 = (config.weight_decay_mult / n) * (v.flat * v.flat).sum().reshape(1)
        d_flat.add_(v.flat, alpha=2.0 * config.weight_decay_mult / n)
    stats, grads = loss_and_grads(model, config, ret, batch, eps_, tv=tv, weight_l2=wl2, workspace=workspace, norms=norms)
    # jax.lax.pmean(grad, 'batch') (:253): per-network buckets all-reduced on a side stream while the backward continues
    # (DURF_ALLREDUCE=single: one all-reduce of the whole flat gradient after the backward, the round-1 behaviour)
    bucketed = world_size > 1 and os.environ.get('DURF_ALLREDUCE', 'buckets') == 'buckets'
    side = parallel.GradientBuckets(v, d_flat, world_size) if bucketed else None
    model.backward(v, ctx, grads, d_flat, on_network_done=None if side is None else side.network_done)
    if side is not None:
        scale = side.finish()
    else:
        if world_size > 1 and os.environ.get('DURF_ALLREDUCE') == 'none':     # diagnosis only: replicas drift apart
            scale = 1.0 / world_size
        else:
            scale = parallel.allreduce_gradients(d_flat) if world_size > 1 else 1.0
    ops.grad_sanitize(d_flat, config.grad_max_val, scale, sumsq)
    if dev_mode:
        ops.adam_step(v.flat, d_flat, state.m, state.v, sumsq, max_norm=config.grad_max_norm, lr=scalars.lr, step=scalars.step,
                      advance_step=scalars.advance_on_device)
    else:
        ops.adam_step(v.flat, d_flat, state.m, state.v, sumsq, max_norm=config.grad_max_norm, lr=lr, step=state.step)
    v.mark_dirty()
    state.step += 1
    stats['grad_norm_sq'] = sumsq[0]           # |g|^2 after nan_to_num / clip; stats['grad_norm'] takes the root when it is read
    stats['grad'] = d_flat
    stats['pose'] = pose                       # the forward pass's pose (stats.pose, :255): what `prevs` is updated with
    return state, stats


class Stats(dict):
    """The step's statistics.  stats['grad_norm'] (train_boxpose.py:283) is formed from 'grad_norm_sq' when it is read, not as
    one more kernel of every step."""

    def __missing__(self, key):
        if key == 'grad_norm':
            return torch.sqrt(self['grad_norm_sq'])
        raise KeyError(key)


def _add_box_grad(v: Variables, d_flat: torch.Tensor, ctx: dict, g: torch.Tensor, cols: slice) -> None:
    """d box_centers[ts, :, cols] += g, with ts a Python int or a device index."""
    bc = v.view_of(d_flat, 'box_centers')
    ts = ctx['ts']
    if torch.is_tensor(ts):
        full = torch.zeros(1, v.K, 6, device=d_flat.device)
        full[0, :, cols] = g
        bc.index_add_(0, ts.reshape(1), full)
    else:
        bc[ts, :, cols] += g


class GraphedTrainStep:
    """One train step captured in a CUDA graph and replayed: inputs are copied into static device buffers, per-step
    scalars into `StepScalars`, then a single cudaGraphLaunch runs the ~100 kernels of the step.

        step = GraphedTrainStep(model, config, state, B, K, world_size=1)
        stats = step(batch, lr, eps, alpha, rng=dict(t_rand=..., u_rand=...))     # device tensors, no sync

    The graph is captured lazily on the first call (after `warmup` eager steps on a side stream, so that every buffer of
    the caching allocator and every lazily-built weight image exists)."""

    def __init__(self, model: MipNerfModel, config: Config, state: TrainState, B: int, K: int, world_size: int = 1,
                 device='cuda', use_prev: bool = False):
        self.model, self.config, self.state, self.B, self.K, self.world_size = model, config, state, B, K, world_size
        dev = torch.device(device)
        N = model.num_samples
        z3 = lambda: torch.zeros(B, 3, device=dev)
        z1 = lambda: torch.zeros(B, 1, device=dev)
        self.rays = Rays(z3(), z3(), z3(), z1(), z1(), z1(), z1())
        self.buf = dict(pixels=z3(), depth=z1(), sky=z1(), ext=torch.zeros(K, 3, device=dev))
        self.rand = dict(t_rand=torch.zeros(B, N + 1, device=dev), u_rand=torch.zeros(B, N + 1, device=dev))
        if model.density_noise > 0:
            self.rand['density_noise'] = [torch.zeros(B, N, device=dev) for _ in range(model.num_levels)]
        self.prev = torch.zeros(1, K, 6, device=dev) if use_prev else None
        self.scalars = StepScalars.create(dev, step=state.step, advance_on_device=False)
        self.workspace: dict = {}
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.kernels_per_replay = 0
        self.stats: Optional[dict] = None

    def _load(self, batch, lr, eps, alpha, rng, prev):
        for dst, src in zip(self.rays, batch['rays']):
            dst.copy_(src.reshape(dst.shape), non_blocking=True)
        for k in ('pixels', 'depth', 'sky'):
            self.buf[k].copy_(batch[k].reshape(self.buf[k].shape), non_blocking=True)
        self.buf['ext'].copy_(torch.as_tensor(batch['ext']).reshape(self.K, 3), non_blocking=True)
        if rng is not None:
            for k, dst in self.rand.items():
                if k == 'density_noise':
                    for d, s in zip(dst, rng[k]):
                        d.copy_(s.reshape(d.shape), non_blocking=True)
                else:
                    dst.copy_(rng[k], non_blocking=True)
        else:
            self.rand['t_rand'].uniform_(); self.rand['u_rand'].uniform_()
            for d in self.rand.get('density_noise', []):
                d.normal_()
        if self.prev is not None and prev is not None:
            self.prev.copy_(torch.as_tensor(prev).reshape(self.prev.shape), non_blocking=True)
        ts = int(torch.as_tensor(batch['ts']).reshape(-1)[0])
        host = torch.tensor([float(lr), float(eps), float(alpha), 0.0], dtype=torch.float32)
        host.view(torch.int32)[3] = self.state.step                 # the 0-based step this replay's Adam uses
        self.scalars.f.copy_(host.pin_memory(), non_blocking=True)
        self.scalars.ts.copy_(torch.tensor([ts], dtype=torch.int64).pin_memory(), non_blocking=True)

    def _eager(self):
        batch = dict(rays=self.rays, ext=self.buf['ext'], ts=self.scalars.ts, pixels=self.buf['pixels'], depth=self.buf['depth'],
                     sky=self.buf['sky'])
        _, stats = train_step(self.model, self.config, self.rand, self.state, batch, None, None, None, prev=self.prev,
                              world_size=self.world_size, scalars=self.scalars, workspace=self.workspace)
        return stats

    def __call__(self, batch, lr, eps, alpha, rng=None, prev=None, warmup: int = 2):
        self._load(batch, lr, eps, alpha, rng, prev)
        if self.graph is None:
            # warm-up on a side stream (allocator pools, packed weight images, NCCL communicators), then capture.  The
            # warm-up steps are real steps: parameters, moments and counters are restored afterwards.
            st = self.state
            snap = (st.variables.flat.clone(), st.m.clone(), st.v.clone(), self.scalars.step.clone(), st.step)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(warmup):
                    self._eager()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            before = ops.launch_count()
            with torch.cuda.graph(self.graph):
                self.stats = self._eager()          # recorded, not executed
            self.kernels_per_replay = ops.launch_count() - before      # library kernels inside one replay
            st.variables.flat.copy_(snap[0]); st.m.copy_(snap[1]); st.v.copy_(snap[2]); self.scalars.step.copy_(snap[3])
            st.step = snap[4]
            st.variables.mark_dirty()
        self.graph.replay()
        self.state.step += 1
        self.state.variables.mark_dirty()           # the replayed Adam changed the parameters behind the host flag's back
        return self.stats
