"""Host-side mirror of the reference's internal/obbpose_model.py, driving the CUDA kernels.

`MipNerfModel.apply(variables, rng, rays, init, ext, ts, randomized, rand_bkgd, white_bkgd, alpha)` keeps the
argument order and the per-level 10-tuple of MipNerfModel.__call__ (obbpose_model.py:69-261); `render_image`
keeps obbpose_model.py:421-479.  Tensors are torch CUDA tensors instead of jnp arrays.

Differences that are part of the boundary (documented in INTEGRATION.md):
  * `rng` carries the explicit random buffers (the reference's threefry streams cannot be reproduced without JAX):
    a dict with optional 't_rand' [B,N+1], 'u_rand' [B,N+1], 'density_noise' [L][B,N]; or a torch.Generator / None,
    in which case the buffers are drawn on the device.
  * parameters live in one flat fp32 tensor (`Variables.flat`) with named views, so the optimizer and the gradient
    all-reduce are single launches; `Variables.to_flax_dict()` gives the reference's params/MLP_0/Dense_i/{kernel,bias}.
  * object MLPs run only on the rays that hit their box (result-identical to the reference's evaluate-everything-and-
    mask, obbpose_model.py:174-201, because masked rows are multiplied by 0).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import ops
from .utils import Rays


@dataclass
class MipNerfModel:
    """Fields and defaults of the reference's gin-configurable MipNerfModel (obbpose_model.py:45-66); the values of
    configs/carla_dyn.gin are the defaults here where the gin file overrides the class default."""
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
    num_objects: int = 2
    density_noise: float = 0.0
    density_bias: float = -1.0
    rgb_padding: float = 0.001
    disable_integration: bool = False
    contraction: bool = True
    dynamics: bool = True
    timesteps: int = 5
    no_pose_opt: bool = True
    no_yaw_opt: bool = True
    # MLP / BoxMLP shapes (obbpose_model.py:294-303, 358-367 + gin)
    net_width: int = 256
    box_net_width: int = 128
    net_depth: int = 8
    skip_layer: int = 4
    net_width_condition: int = 128
    # which MLP kernels to use: 'bf16' = tcgen05 chain (throughput), 'fp32' = CUDA-core parity mode
    precision: str = 'bf16'
    # Rows reserved per object for the compacted hit-ray buffers of the tensor-core path (None = the whole batch).  The
    # hit COUNT stays on the device (no host sync); lower this to save memory when few rays hit a box -- rays beyond the
    # cap would be dropped, so `apply` records the overflow in ctx['obj_overflow'] (a device flag) for the caller to check.
    max_obj_rays: Optional[int] = None
    # on the tensor-core path the MLP kernel generates its own input tiles (frustum -> Gaussian -> contraction ->
    # IPE in two warps of the kernel, written straight into the shared-memory A operand): no ray-march launch and no
    # 16 KB/ray-level feature image in HBM.  False = separate durf_raymarch_fwd launches (the round-1 path).
    fuse_raymarch: bool = field(default_factory=lambda: os.environ.get('DURF_FUSE_RAYMARCH', '1') != '0')
    # Opt-in (DURF_BWD_OVERLAP=1): batches of at least `overlap_min_rays` rays run the background network's two backward kernels
    # CONCURRENTLY on disjoint SMs, the weight-gradient kernel consuming each dZ block out of L2 as soon as the chain has
    # published it (ops.OverlappedBackward, durf_mlp_bwd_data / durf_mlp_bwd_weights).  Result-identical and tested, but on a
    # B200 it is not faster than running them back to back (both kernels are bound per SM, not by HBM, so
    # splitting the SMs between them only moves the time around): off by default.
    overlap_backward: bool = field(default_factory=lambda: os.environ.get('DURF_BWD_OVERLAP', '0') == '1')
    # Batches of up to `shared_level_max_rays` rays: the levels of a training forward share one set of activation / mask / tile
    # buffers, and the background network's backward is ONE data-gradient and ONE weight-gradient launch over all of them (the
    # tail wave and the accumulator flush are paid once: -6.5 % at 512 rays, -1.7 % at 2,048; at 16,384 rays a pair of launches
    # per level is 0.7 % faster - the weight-gradient kernel still finds part of its level's dZ in L2 - and is kept)
    shared_level_backward: bool = field(default_factory=lambda: os.environ.get('DURF_SHARED_LEVELS', '1') != '0')
    # Batches of up to `concurrent_objects_max_rays` rays on the fused tensor-core path: every object network's forward runs on
    # its own side stream NEXT TO the background network's instead of after it.  It writes compact rows (DurfMlpArgs.accumulate
    # == 2; at level 0 it re-forms the fenceposts itself, DURF_RM_NO_TVALS_OUT) that durf_mlp_merge_raw adds into the per-ray
    # outputs in object order afterwards - the sums of the serial path bit for bit.  At the reference's 512-ray batch the
    # background network is 3.5 waves of tiles: the objects' tiles run on the SMs its last wave leaves idle.
    # None (default, DURF_OBJ_CONCURRENT unset): only while the step is being captured into a CUDA graph - launched from Python
    # a 512-ray step is bound by the host, and the extra stream switches cost it 0.15 ms.
    concurrent_objects: Optional[bool] = field(default_factory=lambda: {'0': False, '1': True}.get(os.environ.get('DURF_OBJ_CONCURRENT')))
    concurrent_objects_max_rays: int = 4096
    shared_level_max_rays: int = 4096
    overlap_min_rays: int = 2048

    # -- topology helpers ------------------------------------------------------------------------------
    def bg_topology(self):
        return (6 * (self.max_deg_point - self.min_deg_point), self.net_width, self.net_depth, self.skip_layer,
                3 + 6 * self.deg_view, self.net_width_condition)

    def box_topology(self):
        return (3 + 6 * (self.max_deg_point - self.min_deg_point), self.box_net_width, self.net_depth, self.skip_layer,
                3 + 6 * self.deg_view, self.net_width_condition)

    # -- parameters --------------------------------------------------------------------------------------
    def init(self, rng: np.random.Generator, box_centers_init, device='cuda', bias_scale: float = 0.0) -> "Variables":
        """construct_mipnerf / model.init (obbpose_model.py:264-291): glorot-uniform kernels, zero biases,
        box_centers <- init (init_boxes, :35-39)."""
        from .synthetic import glorot_mlp
        init = np.asarray(box_centers_init, np.float32)
        if init.ndim < 3:
            init = init[:, None, :]
        K = init.shape[1]
        v = Variables.allocate(self, K, init.shape[0], device)
        bt, ot = self.bg_topology(), self.box_topology()
        v.load_mlp('MLP_0', glorot_mlp(rng, bt[0], bt[1], bias_scale, depth=bt[2], skip=bt[3], cond_dim=bt[4], cond_width=bt[5]))
        for k in range(K):
            v.load_mlp(f'BoxMLP_{k}', glorot_mlp(rng, ot[0], ot[1], bias_scale, depth=ot[2], skip=ot[3], cond_dim=ot[4],
                                                 cond_width=ot[5]))
        v.box_centers.copy_(torch.from_numpy(init).to(device))
        v.mark_dirty()
        return v

    # -- forward -----------------------------------------------------------------------------------------
    def apply(self, variables: "Variables", rng, rays: Rays, init, ext, ts, randomized: bool, rand_bkgd: bool,
              white_bkgd: bool, alpha: float, ctx: Optional[dict] = None):
        """MipNerfModel.__call__ (obbpose_model.py:69-261).  `init` is accepted for signature parity (the box
        parameters live in `variables`, as `self.param('box_centers', ...)` does after initialisation).
        When `ctx` (a dict) is given the forward keeps what the backward pass needs in it."""
        if self.lindisp:
            # the reference's lindisp branch (mip.py:354-356) is never enabled by the shipped gin files; refuse rather
            # than silently sample linearly in depth
            raise NotImplementedError("lindisp=True is not implemented (configs/*.gin set MipNerfModel.lindisp = False)")
        if not self.stop_level_grad and ctx is not None:
            raise NotImplementedError("stop_level_grad=False: the backward pass treats resampled t_vals as constants "
                                      "(mip.py:413-414 with the reference's default stop_level_grad=True)")
        prec = L.PREC_BF16 if self.precision == 'bf16' else L.PREC_FP32
        # Joint box-pose optimisation needs the gradient w.r.t. the object MLPs' input features.  The tensor-core backward
        # produces it for width-128 networks (the BoxMLP default); other widths fall to the fp32 kernels for the objects.
        pose_train = ctx is not None and self.dynamics and not (self.no_pose_opt and self.no_yaw_opt)
        obj_prec = L.PREC_FP32 if (pose_train and self.box_net_width != 128) else prec
        N = self.num_samples
        origins, dirs = ops.f32(rays.origins), ops.f32(rays.directions)
        B = origins.shape[0]
        dev = origins.device
        if torch.is_tensor(ts) and ts.is_cuda:
            # device-resident timestep (captured CUDA graphs): gather the row on the device, no host read
            ts_i = ts.reshape(-1)[:1].long()
            box = variables.box_centers.index_select(0, ts_i)[0].contiguous()
        else:
            ts_h = ts.reshape(-1) if torch.is_tensor(ts) else torch.as_tensor(np.asarray(ts)).reshape(-1)
            # the reference indexes pose_offsets[ts.squeeze()] (obbpose_model.py:99): one timestep per batch
            # ('timestep' batching); a batch mixing timesteps would silently use the wrong boxes
            assert bool((ts_h == ts_h[0]).all()), "all rays of a batch must share one timestep (Config.batching = 'timestep')"
            ts_i = int(ts_h[0])
            box = variables.box_centers[ts_i].contiguous()                   # [K,6]
        K = box.shape[0]
        ext = ops.f32(torch.as_tensor(ext, device=dev)).reshape(K, 3)
        rb = _rand_buffers(rng, randomized, B, N, self.num_levels, self.density_noise, dev)

        fe = ops.obb_frontend(origins, dirs, box, ext)
        origins_s, dirs_s, hit = fe['origins_s'], fe['dirs_s'], fe['hit']
        viewenc = ops.viewdir_enc(rays.viewdirs, self.deg_view)
        radii = ops.f32(rays.radii).reshape(-1)

        obj_lists = []
        bg_mult = None
        if self.dynamics:
            bg_mult = fe['nhit']               # the kernels form 1 - sum_k mask_k (obbpose_model.py:205) from it: ray_mult_is_nhit
            cap = B if self.max_obj_rays is None else min(B, self.max_obj_rays)
            overflow = torch.zeros((), device=dev, dtype=torch.bool) if cap < B else None
            idx_all, cnt_all = ops.compact_hits_all(hit)
            for k in range(K):
                idx, cnt = idx_all[k], cnt_all[k:k + 1]
                if overflow is not None:
                    overflow = overflow | (cnt[0] > cap)
                m_host = None
                if obj_prec == L.PREC_FP32:
                    m_host = int(cnt.item())                                 # the fp32 parity GEMMs are sized on the host
                obj_lists.append((idx, cnt, m_host))
        bt, ot = self.bg_topology(), self.box_topology()
        bf16 = prec == L.PREC_BF16
        if bf16:
            variables.ensure_packed(self)

        ret = []
        t_vals = weights = None
        if ctx is not None:
            ctx.update(dict(fe=fe, viewenc=viewenc, obj_lists=obj_lists, levels=[], B=B, K=K, ts=ts_i,
                            obj_overflow=overflow if self.dynamics else None, alpha=alpha, obj_prec=obj_prec,
                            white_bkgd=white_bkgd, rand_bkgd=rand_bkgd, rays=rays, box=box, radii=radii))
        for i_level in range(self.num_levels):
            common = dict(min_deg=self.min_deg_point, max_deg=self.max_deg_point, ray_shape=self.ray_shape,
                          integrate=not self.disable_integration, bf16_tiles=bf16)
            fuse = bf16 and self.fuse_raymarch and (self.max_deg_point - self.min_deg_point) == 10 and N == 128
            if i_level > 0:
                t_vals = ops.resample(t_vals, weights, u_rand=rb['u_rand'], padding=self.resample_padding)
            rm_kw = dict(contract=self.contraction, ray_mult=bg_mult, ray_mult_is_nhit=True, min_deg=self.min_deg_point, max_deg=self.max_deg_point,
                         ray_shape=self.ray_shape, integrate=not self.disable_integration)
            if i_level == 0:
                rm_kw.update(near=rays.near, far=rays.far, t_rand=rb['t_rand'])
            else:
                rm_kw.update(t_vals=t_vals)
            # object networks next to the background network (see `concurrent_objects`)
            want_conc = self.concurrent_objects if self.concurrent_objects is not None else torch.cuda.is_current_stream_capturing()
            conc = (fuse and self.dynamics and obj_prec == L.PREC_BF16 and want_conc and K > 0
                    and B <= self.concurrent_objects_max_rays)
            pending = []
            if conc:
                cur = torch.cuda.current_stream()
                side = _object_streams(dev, K)
                o_kw = dict(weighted=True, alpha=alpha, min_deg=self.min_deg_point, max_deg=self.max_deg_point,
                            ray_shape=self.ray_shape, integrate=not self.disable_integration)
                if i_level == 0:
                    o_kw.update(near=rays.near, far=rays.far, t_rand=rb['t_rand'], store_t_vals=False)
                else:
                    o_kw.update(t_vals=t_vals)
                for k, (idx, cnt, _m) in enumerate(obj_lists):
                    rows = B if self.max_obj_rays is None else min(B, self.max_obj_rays)
                    # every buffer is allocated on the calling stream (the side stream only runs the kernel)
                    feat_o = torch.empty(rows, 128 * 64, device=dev, dtype=torch.bfloat16) if ctx is not None else None
                    saved_o = ops.mlp_saved_buffer(ot, rows, N, dev) if ctx is not None else None
                    c_rgb, c_den = torch.empty(rows, N, 3, device=dev), torch.empty(rows, N, device=dev)
                    fzo, _, _keep_o = ops.fused_raymarch_args(origins_s, dirs_s, radii, N, **o_kw)
                    side[k].wait_stream(cur)
                    with torch.cuda.stream(side[k]):
                        ops.mlp_fwd(ot, feat_o, viewenc, variables.blob(f'BoxMLP_{k}'), M=rows, N=N, precision=obj_prec,
                                    packed=variables.packed.get(f'BoxMLP_{k}'), ray_index=idx, count=cnt, accumulate=2,
                                    raw_rgb=c_rgb, raw_density=c_den, save=ctx is not None, fused=fzo, saved_buf=saved_o,
                                    saved_total=rows)
                    pending.append(dict(feat=feat_o, saved=saved_o, rows=rows, count=cnt, c_rgb=c_rgb, c_den=c_den, idx=idx,
                                        keep=_keep_o))
            if fuse:
                # training keeps the tile image in HBM as well: the weight-gradient kernel reads it (job 0 and the skip layer).
                # All levels write their tiles / activations / masks into ONE set of buffers (level i = tiles i*B .. (i+1)*B - 1),
                # so that the background network's backward is one data-gradient and one weight-gradient launch for all levels
                # (at the reference's 512-ray batch a launch is 3.5 waves: the tail wave and the accumulator flush are paid once).
                shared = None
                if (ctx is not None and self.shared_level_backward and B <= self.shared_level_max_rays
                        and not (self.overlap_backward and B >= self.overlap_min_rays)):
                    if i_level == 0:
                        ctx['bg_shared'] = dict(saved=ops.mlp_saved_buffer(bt, self.num_levels * B, N, dev),
                                                feat=torch.empty(self.num_levels * B, 128 * 64, device=dev, dtype=torch.bfloat16))
                    shared = ctx['bg_shared']
                if shared is not None:
                    feat_bg = shared['feat'][i_level * B:(i_level + 1) * B]
                else:
                    feat_bg = torch.empty(B, 128 * 64, device=dev, dtype=torch.bfloat16) if ctx is not None else None
                fz, t_vals, _keep = ops.fused_raymarch_args(origins_s, dirs_s, radii, N, **rm_kw)
                raw_rgb, raw_density, saved_bg = ops.mlp_fwd(bt, feat_bg, viewenc, variables.blob('MLP_0'), M=B, N=N, precision=prec,
                                                             packed=variables.packed.get('MLP_0'), save=ctx is not None, fused=fz,
                                                             saved_buf=None if shared is None else shared['saved'],
                                                             saved_offset=i_level * B, saved_total=self.num_levels * B)
            else:
                rm = ops.raymarch(origins_s, dirs_s, radii, N, bf16_tiles=bf16, **rm_kw)
                t_vals, feat_bg = rm['t_vals'], rm['features']
                raw_rgb, raw_density, saved_bg = ops.mlp_fwd(bt, feat_bg, viewenc, variables.blob('MLP_0'), M=B, N=N,
                                                             precision=prec, packed=variables.packed.get('MLP_0'),
                                                             save=ctx is not None)
            lvl_ctx = dict(feat_bg=feat_bg, saved_bg=saved_bg, obj=[]) if ctx is not None else None
            if conc:
                for k, o in enumerate(pending):
                    cur.wait_stream(side[k])
                    ops.mlp_merge_raw(o['c_rgb'], o['c_den'], o['idx'], o['count'], raw_rgb, raw_density)
                    if lvl_ctx is not None:
                        lvl_ctx['obj'].append(dict(feat=o['feat'], saved=o['saved'], rows=o['rows'], count=o['count']))
            elif self.dynamics:
                for k, (idx, cnt, m_host) in enumerate(obj_lists):
                    if m_host == 0:
                        if lvl_ctx is not None:
                            lvl_ctx['obj'].append(None)
                        continue
                    rows = m_host if m_host is not None else (B if self.max_obj_rays is None else min(B, self.max_obj_rays))
                    dev_cnt = None if m_host is not None else cnt
                    if fuse and obj_prec == L.PREC_BF16:
                        feat_o = torch.empty(rows, 128 * 64, device=dev, dtype=torch.bfloat16) if ctx is not None else None
                        fzo, _, _keep_o = ops.fused_raymarch_args(origins_s, dirs_s, radii, N, t_vals=t_vals, weighted=True, alpha=alpha,
                                                                  min_deg=self.min_deg_point, max_deg=self.max_deg_point,
                                                                  ray_shape=self.ray_shape, integrate=not self.disable_integration)
                    else:
                        fzo = None
                        feat_o = ops.raymarch(origins_s, dirs_s, radii, N, t_vals=t_vals, weighted=True, alpha=alpha, ray_index=idx,
                                              count=dev_cnt, rows=rows, **dict(common, bf16_tiles=obj_prec == L.PREC_BF16))['features']
                    _, _, saved_o = ops.mlp_fwd(ot, feat_o, viewenc, variables.blob(f'BoxMLP_{k}'), M=rows, N=N,
                                                precision=obj_prec, packed=variables.packed.get(f'BoxMLP_{k}'), ray_index=idx,
                                                count=dev_cnt, accumulate=True, raw_rgb=raw_rgb, raw_density=raw_density,
                                                save=ctx is not None, fused=fzo)
                    if lvl_ctx is not None:
                        lvl_ctx['obj'].append(dict(feat=feat_o, saved=saved_o, rows=rows, count=dev_cnt))
            if randomized and self.density_noise > 0:
                raw_density = raw_density + self.density_noise * rb['density_noise'][i_level]   # obbpose_model.py:237-240
            comp = ops.composite(raw_rgb, raw_density, t_vals, dirs_s, white_bkgd=white_bkgd, rand_bkgd=rand_bkgd,
                                 density_bias=self.density_bias)
            weights = comp['weights']
            dyn_mask = fe['nhit'].reshape(B, 1)
            ret.append((comp['comp_rgb'], comp['depth'], comp['acc'], weights, t_vals, comp['t_mids'], comp['t_dists'],
                        [box[:, :3], box[0, 3:]], dyn_mask, fe['zo_ret']))   # [box_pose[0], box_rot[0]] (obbpose_model.py:258): [K,3], [3]
            if lvl_ctx is not None:
                lvl_ctx.update(raw_rgb=raw_rgb, raw_density=raw_density, t_vals=t_vals)
                ctx['levels'].append(lvl_ctx)
        return ret

    def step_is_capturable(self) -> bool:
        """True if a train step of this model makes no host read (train.GraphedTrainStep can capture it): the tensor-core
        path sizes the object networks' work from device counts; the fp32 parity kernels are sized on the host."""
        pose_train = self.dynamics and not (self.no_pose_opt and self.no_yaw_opt)
        return self.precision == 'bf16' and not (pose_train and self.box_net_width != 128)

    # -- backward ----------------------------------------------------------------------------------------
    def backward(self, variables: "Variables", ctx: dict, level_grads: Sequence[dict], d_flat: torch.Tensor,
                 on_network_done=None) -> None:
        """Reverse of `apply` (what jax.value_and_grad does through model.apply, train_boxpose.py:251): given per level
        dL/d(comp_rgb, depth, weights) accumulate dL/d parameters into `d_flat` (same layout as variables.flat).
        Levels are independent: t_vals are stop_gradient'ed (mip.py:413-414).

        Order: background MLP (both levels), then each object MLP, then the box parameters; `on_network_done(name)` is
        called as soon as a network's slice of d_flat is final, so its all-reduce overlaps the rest of the backward."""
        N = self.num_samples
        fe, viewenc = ctx['fe'], ctx['viewenc']
        B, K = ctx['B'], ctx['K']
        bt, ot = self.bg_topology(), self.box_topology()
        pose_opt = self.dynamics and not (self.no_pose_opt and self.no_yaw_opt)
        d_os = torch.zeros(B, 3, device=d_flat.device) if pose_opt else None
        d_ds = torch.zeros(B, 3, device=d_flat.device) if pose_opt else None
        done = on_network_done or (lambda name: None)
        prec = L.PREC_BF16 if self.precision == 'bf16' else L.PREC_FP32
        raw_grads = []
        # opt-in: the background network's dZ chain and its weight-gradient kernel run concurrently on disjoint SMs
        # (ops.OverlappedBackward); the weight kernel of level 0 keeps going under the chain of level 1.
        overlap = None
        if prec == L.PREC_BF16 and self.overlap_backward and B >= self.overlap_min_rays:
            overlap = ops.OverlappedBackward()
        shared = ctx.get('bg_shared') if overlap is None else None      # all levels' records in one buffer: one backward call
        nl = len(ctx['levels'])
        if shared is not None:
            g_rgb_all = torch.empty(nl, B, N, 3, device=d_flat.device)
            g_den_all = torch.empty(nl, B, N, device=d_flat.device)
        for i_lvl, (lvl, g) in enumerate(zip(ctx['levels'], level_grads)):
            g_rgb, g_den, g_dirs = ops.composite_bwd(lvl['raw_rgb'], lvl['raw_density'], lvl['t_vals'], fe['dirs_s'],
                                                     g['comp_rgb'], g['depth'], g['weights'], white_bkgd=ctx['white_bkgd'],
                                                     rand_bkgd=ctx['rand_bkgd'], density_bias=self.density_bias,
                                                     want_d_dirs=pose_opt, out_rgb=None if shared is None else g_rgb_all[i_lvl],
                                                     out_density=None if shared is None else g_den_all[i_lvl])
            raw_grads.append((g_rgb, g_den))
            if pose_opt:
                d_ds += g_dirs * (fe['nhit'] > 0).float()[:, None]            # only object rays carry pose-dependent dirs
        # The networks' backward passes are independent of each other (disjoint slices of d_flat, read-only upstream gradients):
        # without pose optimisation every object network runs on its own side stream next to the background network.  That
        # matters for small batches, where each of the ~12 backward kernels is a few waves long and latency-bound (the
        # reference's shipped batch is 512 rays); large batches fill the GPU with the background network alone.
        n_obj = K if self.dynamics else 0
        side = _object_streams(d_flat.device, n_obj) if (n_obj and not pose_opt and ctx['obj_prec'] == L.PREC_BF16 and
                                                        os.environ.get('DURF_OBJ_STREAMS', '1') != '0') else None
        cur = torch.cuda.current_stream()
        if side is not None:
            for sk in side:
                sk.wait_stream(cur)

        def object_backward(k):
            for lvl, (g_rgb, g_den) in zip(ctx['levels'], raw_grads):
                o = lvl['obj'][k]
                if o is None:
                    continue
                idx = ctx['obj_lists'][k][0]
                dfeat = ops.mlp_bwd(ot, o['feat'], viewenc, variables.blob(f'BoxMLP_{k}'), o['saved'], g_rgb, g_den,
                                    variables.blob_of(d_flat, f'BoxMLP_{k}'), M=o['rows'], N=N, ray_index=idx,
                                    want_d_features=pose_opt, precision=ctx['obj_prec'], packed=variables.packed.get(f'BoxMLP_{k}'),
                                    count=o.get('count'))
                if pose_opt:
                    go = torch.zeros(B, 3, device=d_flat.device)
                    gd = torch.zeros(B, 3, device=d_flat.device)
                    ops.raymarch_bwd(fe['origins_s'], fe['dirs_s'], ctx['radii'], lvl['t_vals'], dfeat, weighted=True,
                                     alpha=ctx['alpha'], min_deg=self.min_deg_point, max_deg=self.max_deg_point, ray_index=idx,
                                     rows=o['rows'], count=o.get('count'), d_origins=go, d_dirs=gd)
                    d_os.add_(go)
                    d_ds.add_(gd)

        if side is not None:
            for k, sk in enumerate(side):
                with torch.cuda.stream(sk):
                    object_backward(k)
        if shared is not None:
            ops.mlp_bwd(bt, shared['feat'], viewenc.repeat(nl, 1), variables.blob('MLP_0'), shared['saved'],
                        g_rgb_all.view(nl * B, N, 3), g_den_all.view(nl * B, N), variables.blob_of(d_flat, 'MLP_0'), M=nl * B, N=N,
                        precision=prec, packed=variables.packed.get('MLP_0'))
        for lvl, (g_rgb, g_den) in zip(ctx['levels'], raw_grads):
            if shared is not None:
                break
            if overlap is not None:
                overlap(bt, lvl['feat_bg'], viewenc, variables.blob('MLP_0'), lvl['saved_bg'], g_rgb, g_den,
                        variables.blob_of(d_flat, 'MLP_0'), M=B, N=N, packed=variables.packed.get('MLP_0'))
            else:
                ops.mlp_bwd(bt, lvl['feat_bg'], viewenc, variables.blob('MLP_0'), lvl['saved_bg'], g_rgb, g_den,
                            variables.blob_of(d_flat, 'MLP_0'), M=B, N=N, precision=prec, packed=variables.packed.get('MLP_0'))
        if overlap is not None:
            overlap.join()
        done('MLP_0')
        for k in range(n_obj):
            if side is not None:
                cur.wait_stream(side[k])
            else:
                object_backward(k)
            done(f'BoxMLP_{k}')
        for k in range(0 if self.dynamics else K):
            done(f'BoxMLP_{k}')                                              # static scene: object networks get no gradient
        if pose_opt:
            d_box = torch.zeros(K, 6, device=d_flat.device)
            ops.obb_frontend_bwd(ctx['rays'].origins, ctx['rays'].directions, ctx['box'], fe['hit'], d_os, d_ds,
                                 pose_grad=not self.no_pose_opt, rot_grad=not self.no_yaw_opt, d_box=d_box)
            bc = variables.view_of(d_flat, 'box_centers')
            if torch.is_tensor(ctx['ts']):
                bc.index_add_(0, ctx['ts'], d_box[None])
            else:
                bc[ctx['ts']] += d_box
        done('box_centers')


_OBJ_STREAMS: Dict[Any, List[torch.cuda.Stream]] = {}


def _object_streams(device, n: int) -> List[torch.cuda.Stream]:
    """One side stream per object network (created once per device)."""
    pool = _OBJ_STREAMS.setdefault(device, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


def _rand_buffers(rng, randomized: bool, B: int, N: int, levels: int, density_noise: float, dev) -> dict:
    out = dict(t_rand=None, u_rand=None, density_noise=None)
    if not randomized:
        return out
    if isinstance(rng, dict):
        out.update({k: (None if v is None else v) for k, v in rng.items() if k in out})
        gen = None
    else:
        gen = rng if isinstance(rng, torch.Generator) else None
    if out['t_rand'] is None:
        out['t_rand'] = torch.rand(B, N + 1, device=dev, generator=gen)
    if out['u_rand'] is None:
        out['u_rand'] = torch.rand(B, N + 1, device=dev, generator=gen)
    if density_noise > 0 and out['density_noise'] is None:
        out['density_noise'] = [torch.randn(B, N, device=dev, generator=gen) for _ in range(levels)]
    return out


class Variables:
    """All trainable parameters in ONE flat fp32 device tensor: [MLP_0 | BoxMLP_0 .. BoxMLP_{K-1} | box_centers].
    Names follow the reference's flax tree (params/MLP_0, params/BoxMLP_k, params/box_centers)."""

    def __init__(self, flat: torch.Tensor, slots: Dict[str, Tuple[int, int]], topo: Dict[str, tuple], T: int, K: int):
        self.flat = flat
        self.slots = slots
        self.topo = topo
        self.T, self.K = T, K
        self.packed: Dict[str, torch.Tensor] = {}
        self._dirty = True

    @staticmethod
    def allocate(model: MipNerfModel, K: int, T: int, device) -> "Variables":
        slots, topo, off = {}, {}, 0
        names = [('MLP_0', model.bg_topology())] + [(f'BoxMLP_{k}', model.box_topology()) for k in range(K)]
        for name, t in names:
            n = ops.mlp_param_count(t)
            slots[name] = (off, n)
            topo[name] = t
            off += n
        slots['box_centers'] = (off, T * K * 6)
        off += T * K * 6
        return Variables(torch.zeros(off, device=device, dtype=torch.float32), slots, topo, T, K)

    # views
    def blob_of(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        o, n = self.slots[name]
        return flat[o:o + n]

    def view_of(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        v = self.blob_of(flat, name)
        return v.view(self.T, self.K, 6) if name == 'box_centers' else v

    def blob(self, name: str) -> torch.Tensor:
        return self.blob_of(self.flat, name)

    @property
    def box_centers(self) -> torch.Tensor:
        return self.view_of(self.flat, 'box_centers')

    def layers(self, name: str, flat: Optional[torch.Tensor] = None):
        return ops.mlp_layer_views(self.topo[name], self.blob_of(self.flat if flat is None else flat, name))

    def load_mlp(self, name: str, layers) -> None:
        for (w, b), (kw, kb) in zip(self.layers(name), layers):
            w.copy_(torch.as_tensor(kw)); b.copy_(torch.as_tensor(kb))
        self._dirty = True

    def mark_dirty(self) -> None:
        self._dirty = True

    def ensure_packed(self, model: MipNerfModel) -> None:
        """(Re)build the tensor-core weight images after the parameters changed."""
        if not self._dirty and self.packed:
            return
        names = list(self.topo)
        images = ops.mlp_pack_multi([self.topo[n] for n in names], [self.blob(n) for n in names], [self.packed.get(n) for n in names])
        self.packed.update(zip(names, images))
        self._dirty = False

    def to_flax_dict(self) -> Dict[str, Any]:
        """The reference's parameter tree: params/{MLP_0,BoxMLP_k}/Dense_i/{kernel,bias}, params/box_centers."""
        tree: Dict[str, Any] = {}
        for name in self.topo:
            tree[name] = {f'Dense_{i}': {'kernel': w, 'bias': b} for i, (w, b) in enumerate(self.layers(name))}
        tree['box_centers'] = self.box_centers
        return {'params': tree}


def construct_mipnerf(rng: np.random.Generator, example_batch: dict, device='cuda', **model_kwargs):
    """obbpose_model.py:264-291: build the model and its variables from an example batch (uses 'init')."""
    model = MipNerfModel(**model_kwargs)
    init = np.asarray(example_batch['init'], np.float32).squeeze()
    return model, model.init(rng, init, device=device)


def render_image(render_fn, rays: Rays, init, ext, ts, rng, alpha, chunk: int = 8192):
    """obbpose_model.py:421-479: render all pixels of a frame in `chunk`-ray slices; returns the fine level's
    (rgb[H,W,3], distance[H,W], acc[H,W]).  `render_fn(rng, batch)` returns the model's per-level list, like the
    reference's pmapped render_eval_pfn.  Rays may live on the host (pinned) or on the device; every chunk is moved
    with non-blocking copies on the current stream, results stay on the device until the end (one sync per frame,
    not one per chunk)."""
    height, width = rays[0].shape[:2]
    num_rays = height * width
    flat = Rays(*[r.reshape(num_rays, -1) for r in rays])
    dev = torch.device('cuda')
    rgb = torch.empty(num_rays, 3, device=dev)
    dist = torch.empty(num_rays, device=dev)
    acc = torch.empty(num_rays, device=dev)
    for i in range(0, num_rays, chunk):
        chunk_rays = Rays(*[r[i:i + chunk].to(dev, non_blocking=True) for r in flat])
        batch = dict(rays=chunk_rays, init=init, ext=ext, ts=ts, alpha=alpha)
        out = render_fn(rng, batch)[-1]
        n = chunk_rays.origins.shape[0]
        rgb[i:i + n] = out[0]; dist[i:i + n] = out[1]; acc[i:i + n] = out[2]
    return rgb.reshape(height, width, 3), dist.reshape(height, width), acc.reshape(height, width)


def render_camera(render_fn, c2w, width: int, height: int, focal: float, near: float, far: float, init, ext, ts, rng, alpha,
                  chunk: int = 65536, principal_point=None):
    """`render_image` for a pinhole camera whose rays are generated ON THE DEVICE (durf_generate_rays, the restatement of
    obbpose_dataset.py:613-661; `principal_point` selects the Waymo variant, :1868-1917): no per-ray host->device traffic at
    all.  Chunks are whole image rows; returns the fine level's (rgb[H,W,3], distance[H,W], acc[H,W]) on the device."""
    dev = torch.device('cuda')
    rows_per_chunk = max(1, chunk // width)
    rgb = torch.empty(height * width, 3, device=dev)
    dist = torch.empty(height * width, device=dev)
    acc = torch.empty(height * width, device=dev)
    for r0 in range(0, height, rows_per_chunk):
        r1 = min(height, r0 + rows_per_chunk)
        rays = ops.generate_rays(c2w, width, height, focal, near, far, r0, r1, device=dev, principal_point=principal_point)
        out = render_fn(rng, dict(rays=rays, init=init, ext=ext, ts=ts, alpha=alpha))[-1]
        i, n = r0 * width, (r1 - r0) * width
        rgb[i:i + n] = out[0]; dist[i:i + n] = out[1]; acc[i:i + n] = out[2]
    return rgb.reshape(height, width, 3), dist.reshape(height, width), acc.reshape(height, width)
