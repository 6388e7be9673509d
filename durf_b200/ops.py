"""One Python function per C-ABI entry point: allocates outputs with torch, passes raw device pointers.

This is the lowest host layer; `mip.py`, `mip360.py`, `box_helpers.py`, `math.py`, `obbpose_model.py` and
`train.py` (the mirrors of the reference's modules) are written on top of it.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from . import _lib as L
from ._lib import MlpTopology, check, f32, ptr, stream_ptr

BG_TOPOLOGY = (60, 256, 8, 4, 27, 128)    # MLP, configs/carla_dyn.gin:55-58
BOX_TOPOLOGY = (63, 128, 8, 4, 27, 128)   # BoxMLP defaults, obbpose_model.py:360-363


# bench.py sets PROFILE = {'mlp': []} to collect a CUDA-event pair around every durf_mlp_fwd launch (same stream).
PROFILE: Optional[dict] = None


def topology(t) -> MlpTopology:
    return t if isinstance(t, MlpTopology) else MlpTopology(*t)


def _dev(t: torch.Tensor):
    if not t.is_cuda:
        raise L.DurfError("durf_b200 needs CUDA tensors; there is no CPU path")
    return t.device


def launch_count() -> int:
    return int(L.load().durf_launch_count())


def reset_launch_count() -> None:
    L.load().durf_reset_launch_count()


# ---- N2: ray generation on the device ------------------------------------------------------------------
def generate_rays(c2w, width: int, height: int, focal: float, near: float, far: float, row0: int = 0, row1=None, device='cuda',
                  principal_point=None):
    """Rays of the pixel rows [row0,row1) of one pinhole camera as a `utils.Rays` tuple of CUDA tensors
    (obbpose_dataset.py:613-661; [n,3] x3 and [n,1] x4 like the reference's flattened BoxRays).  `principal_point` =
    (cx, cy) in pixels selects the Waymo loader's variant (:1868-1917); None = image centre (Carla)."""
    import numpy as np
    from .utils import Rays
    row1 = height if row1 is None else row1
    n = (row1 - row0) * width
    cam = L.Camera(width=width, height=height, focal=float(focal), near=float(near), far=float(far))
    if principal_point is not None:
        cam.use_principal_point, cam.cx, cam.cy = 1, float(principal_point[0]), float(principal_point[1])
    flat = np.asarray(c2w, np.float32)[:3, :4].reshape(-1)
    for i in range(12):
        cam.c2w[i] = float(flat[i])
    v3 = [torch.empty(n, 3, device=device) for _ in range(3)]
    v1 = [torch.empty(n, 1, device=device) for _ in range(4)]
    check(L.load().durf_generate_rays(stream_ptr(), C.byref(cam), row0, row1, *[ptr(t) if n else None for t in v3 + v1]),
          "durf_generate_rays")
    return Rays(v3[0], v3[1], v3[2], v1[0], v1[1], v1[2], v1[3])


# ---- K0 ---------------------------------------------------------------------------------------------
def aa2matrix(angles: torch.Tensor) -> torch.Tensor:
    angles = f32(angles)
    K = angles.shape[0]
    R = torch.empty(K, 3, 3, device=_dev(angles), dtype=torch.float32)
    check(L.load().durf_aa2matrix_fwd(stream_ptr(), K, ptr(angles), ptr(R)), "durf_aa2matrix_fwd")
    return R


def obb_frontend(origins, directions, box, ext, want_object_rays: bool = False):
    """-> dict(origins_s, dirs_s, hit[B,K] i32, zi, zo, zo_ret[B], nhit[B], [origins_o, dirs_o])."""
    origins, directions, box, ext = f32(origins), f32(directions), f32(box), f32(ext)
    B, K = origins.shape[0], box.shape[0]
    dev = _dev(origins)
    o = dict(origins_s=torch.empty(B, 3, device=dev), dirs_s=torch.empty(B, 3, device=dev),
             hit=torch.empty(B, K, device=dev, dtype=torch.int32), zi=torch.empty(B, K, device=dev),
             zo=torch.empty(B, K, device=dev), zo_ret=torch.empty(B, device=dev), nhit=torch.empty(B, device=dev))
    oo = torch.empty(B, K, 3, device=dev) if want_object_rays else None
    do = torch.empty(B, K, 3, device=dev) if want_object_rays else None
    check(L.load().durf_obb_frontend_fwd(stream_ptr(), B, K, ptr(origins), ptr(directions), ptr(box), ptr(ext),
                                         ptr(o['origins_s']), ptr(o['dirs_s']), ptr(o['hit']), ptr(o['zi']), ptr(o['zo']),
                                         ptr(o['zo_ret']), ptr(o['nhit']), ptr(oo), ptr(do)), "durf_obb_frontend_fwd")
    if want_object_rays:
        o['origins_o'], o['dirs_o'] = oo, do
    return o


def obb_frontend_bwd(origins, directions, box, hit, d_origins_s, d_dirs_s, pose_grad: bool, rot_grad: bool, d_box):
    B, K = origins.shape[0], box.shape[0]
    check(L.load().durf_obb_frontend_bwd(stream_ptr(), B, K, ptr(f32(origins)), ptr(f32(directions)), ptr(f32(box)), ptr(hit),
                                         ptr(d_origins_s), ptr(d_dirs_s), int(pose_grad), int(rot_grad), ptr(d_box)),
          "durf_obb_frontend_bwd")


def compact_hits(hit: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    B, K = hit.shape
    idx = torch.empty(B, device=hit.device, dtype=torch.int32)
    cnt = torch.empty(1, device=hit.device, dtype=torch.int32)
    check(L.load().durf_compact_hits(stream_ptr(), B, K, k, ptr(hit), ptr(idx), ptr(cnt)), "durf_compact_hits")
    return idx, cnt


def compact_hits_all(hit: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """`compact_hits` for every object in one launch: (ray_index int32[K,B], count int32[K])."""
    B, K = hit.shape
    idx = torch.empty(K, B, device=hit.device, dtype=torch.int32)
    cnt = torch.empty(K, device=hit.device, dtype=torch.int32)
    check(L.load().durf_compact_hits_all(stream_ptr(), B, K, ptr(hit), ptr(idx), ptr(cnt)), "durf_compact_hits_all")
    return idx, cnt


def mlp_merge_raw(src_rgb, src_density, ray_index, count, raw_rgb, raw_density) -> None:
    """raw[ray_index[m]] += src[m] for the rows of an object network evaluated with `mlp_fwd(..., accumulate=2)`."""
    M, N = src_density.shape
    check(L.load().durf_mlp_merge_raw(stream_ptr(), M, N, ptr(ray_index), ptr(count), ptr(src_rgb), ptr(src_density),
                                      ptr(raw_rgb), ptr(raw_density)), "durf_mlp_merge_raw")


# ---- K1 ---------------------------------------------------------------------------------------------
def raymarch(origins, dirs, radii, N: int, *, t_vals=None, near=None, far=None, t_rand=None, contract=False,
             weighted=False, alpha=0.0, min_deg=0, max_deg=10, ray_shape='cone', integrate=True, ray_mult=None,
             ray_index=None, count=None, rows=None, bf16_tiles=False, want_gaussians=False, ray_mult_is_nhit=False):
    """Fused sample/cast/contract/encode.  Returns dict(t_vals, features, [means, cov_diag])."""
    if ray_shape not in ('cone', 'cylinder'):
        raise AssertionError("ray_shape must be 'cone' or 'cylinder'")          # mip.py:176 `assert False`
    origins, dirs = f32(origins), f32(dirs)
    radii = f32(radii).reshape(-1)
    B = origins.shape[0]
    dev = _dev(origins)
    flags = 0
    if t_vals is None:
        flags |= L.RM_SAMPLE
        t_vals = torch.empty(B, N + 1, device=dev)
        near, far = f32(near).reshape(-1), f32(far).reshape(-1)
        if t_rand is not None:
            flags |= L.RM_RANDOMIZED
            t_rand = f32(t_rand)
    else:
        t_vals = f32(t_vals)
    if contract: flags |= L.RM_CONTRACT
    if weighted: flags |= L.RM_WEIGHTED
    if ray_shape == 'cylinder': flags |= L.RM_CYLINDER
    if not integrate: flags |= L.RM_NO_INTEGRATE
    if bf16_tiles: flags |= L.RM_OUT_BF16_TILE
    if ray_mult_is_nhit and ray_mult is not None: flags |= L.RM_MULT_IS_NHIT
    F = 6 * (max_deg - min_deg) + (3 if weighted else 0)
    M = B if rows is None else rows          # rows of the output buffers (compacted calls may pass fewer)
    if bf16_tiles:
        tiles = (M * N + 127) // 128
        feats = torch.empty(tiles, 128 * 64, device=dev, dtype=torch.bfloat16)
    else:
        feats = torch.empty(M, N, F, device=dev)
    means = torch.empty(M, N, 3, device=dev) if want_gaussians else None
    covd = torch.empty(M, N, 3, device=dev) if want_gaussians else None
    alpha_dev = alpha if torch.is_tensor(alpha) else None          # device scalar: a captured graph follows alpha_rate_fn
    a = L.RaymarchArgs(B=M, N=N, min_deg=min_deg, max_deg=max_deg, flags=flags, alpha=0.0 if alpha_dev is not None else float(alpha),
                       origins=ptr(origins), dirs=ptr(dirs), radii=ptr(radii), near=ptr(near), far=ptr(far),
                       t_rand=ptr(t_rand), t_vals=ptr(t_vals), ray_mult=ptr(ray_mult), ray_index=ptr(ray_index),
                       count=ptr(count), features=ptr(feats), means=ptr(means), cov_diag=ptr(covd), alpha_dev=ptr(alpha_dev))
    check(L.load().durf_raymarch_fwd(stream_ptr(), C.byref(a)), "durf_raymarch_fwd")
    out = dict(t_vals=t_vals, features=feats)
    if want_gaussians:
        out['means'], out['cov_diag'] = means, covd
    return out


def raymarch_bwd(origins, dirs, radii, t_vals, d_features, *, weighted, alpha, min_deg=0, max_deg=10, ray_mult=None,
                 ray_index=None, rows=None, count=None, d_origins=None, d_dirs=None):
    """Gradient of the object-frame encoding w.r.t. origins_s / dirs_s (written for the rays in ray_index)."""
    B = origins.shape[0]
    N = t_vals.shape[1] - 1
    M = B if rows is None else rows
    flags = L.RM_WEIGHTED if weighted else 0
    alpha_dev = alpha if torch.is_tensor(alpha) else None
    a = L.RaymarchArgs(B=M, N=N, min_deg=min_deg, max_deg=max_deg, flags=flags, alpha=0.0 if alpha_dev is not None else float(alpha),
                       origins=ptr(f32(origins)), dirs=ptr(f32(dirs)), radii=ptr(f32(radii).reshape(-1)), near=None, far=None,
                       t_rand=None, t_vals=ptr(f32(t_vals)), ray_mult=ptr(ray_mult), ray_index=ptr(ray_index), count=ptr(count),
                       features=ptr(d_features), means=None, cov_diag=None, alpha_dev=ptr(alpha_dev))
    check(L.load().durf_raymarch_bwd(stream_ptr(), C.byref(a), ptr(f32(d_features)), ptr(d_origins), ptr(d_dirs)),
          "durf_raymarch_bwd")


def viewdir_enc(viewdirs: torch.Tensor, deg: int = 4) -> torch.Tensor:
    viewdirs = f32(viewdirs)
    B = viewdirs.shape[0]
    enc = torch.empty(B, 3 + 6 * deg, device=_dev(viewdirs))
    check(L.load().durf_viewdir_enc_fwd(stream_ptr(), B, deg, ptr(viewdirs), ptr(enc)), "durf_viewdir_enc_fwd")
    return enc


# ---- K2 ---------------------------------------------------------------------------------------------
def mlp_param_count(topo) -> int:
    return int(L.load().durf_mlp_param_count(C.byref(topology(topo))))


def mlp_layer_views(topo, blob: torch.Tensor):
    """[(kernel[in,out], bias[out])] views into a flat fp32 parameter blob (Dense_0 .. Dense_{depth+3})."""
    t = topology(topo)
    lib = L.load()
    out = []
    for i in range(t.depth + 4):
        ki, ko = C.c_int32(), C.c_int32()
        off = int(lib.durf_mlp_param_offset(C.byref(t), i, C.byref(ki), C.byref(ko)))
        w = blob[off: off + ki.value * ko.value].view(ki.value, ko.value)
        b = blob[off + ki.value * ko.value: off + ki.value * ko.value + ko.value]
        out.append((w, b))
    return out


def mlp_pack(topo, blob: torch.Tensor, packed: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 blob -> tensor-core weight image (bf16, pre-tiled, pre-swizzled)."""
    t = topology(topo)
    nbytes = int(L.load().durf_mlp_packed_bytes(C.byref(t)))
    if nbytes <= 0:
        raise L.DurfError("this MLP topology has no tensor-core path")
    if packed is None:
        packed = torch.empty(nbytes, device=_dev(blob), dtype=torch.uint8)
    check(L.load().durf_mlp_pack_weights(stream_ptr(), C.byref(t), ptr(f32(blob)), ptr(packed)), "durf_mlp_pack_weights")
    return packed


def mlp_pack_multi(topos, blobs, packed_list):
    """`mlp_pack` for several networks in ONE kernel launch (the background MLP and every BoxMLP after an optimizer step).
    `packed_list` entries may be None (allocated here).  Returns the list of weight images."""
    lib = L.load()
    ts = [topology(t) for t in topos]
    out = []
    for t, blob, pk in zip(ts, blobs, packed_list):
        if pk is None:
            nbytes = int(lib.durf_mlp_packed_bytes(C.byref(t)))
            if nbytes <= 0:
                raise L.DurfError("this MLP topology has no tensor-core path")
            pk = torch.empty(nbytes, device=_dev(blob), dtype=torch.uint8)
        out.append(pk)
    n = len(ts)
    if n == 0:
        return out
    arr_t = (L.MlpTopology * n)(*ts)
    blobs = [f32(b) for b in blobs]
    arr_p = (C.c_void_p * n)(*[b.data_ptr() for b in blobs])
    arr_k = (C.c_void_p * n)(*[k.data_ptr() for k in out])
    check(lib.durf_mlp_pack_weights_multi(stream_ptr(), n, arr_t, arr_p, arr_k), "durf_mlp_pack_weights_multi")
    return out


def _mlp_args(topo, precision, M, N, features, cond, blob, packed, ray_index, count, accumulate, raw_rgb, raw_density, saved,
              workspace, fused=None, saved_offset=0, saved_total=0):
    return L.MlpArgs(topo=topology(topo), precision=precision, M=M, N=N, features=ptr(features), cond=ptr(cond),
                     params=ptr(blob), packed=ptr(packed), ray_index=ptr(ray_index), count=ptr(count),
                     accumulate=int(accumulate), raw_rgb=ptr(raw_rgb), raw_density=ptr(raw_density), saved=ptr(saved),
                     workspace=ptr(workspace), workspace_bytes=0 if workspace is None else workspace.numel() * workspace.element_size(),
                     fused_raymarch=None if fused is None else C.addressof(fused), saved_tile_offset=int(saved_offset),
                     saved_total_tiles=int(saved_total))


def fused_raymarch_args(origins, dirs, radii, N: int, *, t_vals=None, near=None, far=None, t_rand=None, contract=False,
                        weighted=False, alpha=0.0, min_deg=0, max_deg=10, ray_shape='cone', integrate=True, ray_mult=None,
                        store_t_vals=True, ray_mult_is_nhit=False):
    """The arguments of `raymarch` as a struct for `mlp_fwd(..., fused=...)` (the tcgen05 MLP kernel generates its
    own input tiles).  Returns (struct, t_vals, keepalive): t_vals is allocated here when it is to be sampled.
    `store_t_vals=False` (sampling only): the fenceposts are formed in registers and not written (t_vals is None) -- an
    object network that runs next to the background network, whose call stores them."""
    if ray_shape not in ('cone', 'cylinder'):
        raise AssertionError("ray_shape must be 'cone' or 'cylinder'")
    origins, dirs = f32(origins), f32(dirs)
    radii = f32(radii).reshape(-1)
    B = origins.shape[0]
    flags = 0
    if t_vals is None:
        flags |= L.RM_SAMPLE
        if store_t_vals:
            t_vals = torch.empty(B, N + 1, device=_dev(origins))
        else:
            flags |= L.RM_NO_TVALS_OUT
        near, far = f32(near).reshape(-1), f32(far).reshape(-1)
        if t_rand is not None:
            flags |= L.RM_RANDOMIZED
            t_rand = f32(t_rand)
    else:
        t_vals = f32(t_vals)
    if contract: flags |= L.RM_CONTRACT
    if weighted: flags |= L.RM_WEIGHTED
    if ray_shape == 'cylinder': flags |= L.RM_CYLINDER
    if not integrate: flags |= L.RM_NO_INTEGRATE
    flags |= L.RM_OUT_BF16_TILE
    if ray_mult_is_nhit and ray_mult is not None: flags |= L.RM_MULT_IS_NHIT     # multiplier = 1 - nhit (obbpose_model.py:205)
    alpha_dev = alpha if torch.is_tensor(alpha) else None
    keep = (origins, dirs, radii, near, far, t_rand, t_vals, ray_mult, alpha_dev)
    a = L.RaymarchArgs(B=B, N=N, min_deg=min_deg, max_deg=max_deg, flags=flags, alpha=0.0 if alpha_dev is not None else float(alpha),
                       origins=ptr(origins), dirs=ptr(dirs), radii=ptr(radii), near=ptr(near), far=ptr(far), t_rand=ptr(t_rand),
                       t_vals=ptr(t_vals), ray_mult=ptr(ray_mult), ray_index=None, count=None, features=None, means=None,
                       cov_diag=None, alpha_dev=ptr(alpha_dev))
    return a, t_vals, keep


def mlp_fwd(topo, features, cond, blob, *, M: int, N: int, precision=L.PREC_BF16, packed=None, ray_index=None, count=None,
            accumulate=False, raw_rgb=None, raw_density=None, num_rays_out=None, save=False, fused=None, saved_buf=None,
            saved_offset=0, saved_total=0):
    """Returns (raw_rgb[B,N,3], raw_density[B,N], saved-or-None).  `fused` = the struct of `fused_raymarch_args`: the kernel
    generates its input tiles itself (`features` may be None, or a tile buffer the generated tiles are also stored to).
    `saved_buf` / `saved_offset` / `saved_total` (tensor-core training): write this call's records into a buffer shared by
    several calls (`mlp_saved_buffer`), so that ONE backward call covers them all."""
    t = topology(topo)
    dev = _dev(features if features is not None else cond)
    Bout = num_rays_out if num_rays_out is not None else M
    if raw_rgb is None:
        raw_rgb = torch.empty(Bout, N, 3, device=dev)
        raw_density = torch.empty(Bout, N, device=dev)
    lib = L.load()
    saved = ws = None
    if precision == L.PREC_FP32:
        if save:
            saved = torch.empty(int(lib.durf_mlp_saved_bytes(C.byref(t), precision, M, N)) // 4, device=dev)
        else:
            ws = torch.empty(int(lib.durf_mlp_workspace_bytes(C.byref(t), precision, M, N, 0)) // 4, device=dev)
    else:
        ws = torch.empty(max(int(lib.durf_mlp_workspace_bytes(C.byref(t), precision, M, N, 0)) // 4, 1), device=dev)
        if save:      # training on the tensor-core path: every layer's bf16 activations as tile images
            saved = saved_buf if saved_buf is not None else mlp_saved_buffer(t, M, N, dev)
    a = _mlp_args(t, precision, M, N, features, f32(cond), f32(blob), packed, ray_index, count, accumulate, raw_rgb, raw_density,
                  saved, ws, fused, saved_offset=saved_offset, saved_total=saved_total if saved_buf is not None else 0)
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.durf_mlp_fwd(stream_ptr(), C.byref(a)), "durf_mlp_fwd")
    if PROFILE is not None:
        e1.record()
        PROFILE['mlp'].append((e0, e1))
    return raw_rgb, raw_density, saved


def mlp_saved_buffer(topo, M: int, N: int, dev) -> torch.Tensor:
    """The buffer `durf_mlp_fwd` keeps a tensor-core training forward's activations and masks in, for M ray-levels."""
    t = topology(topo)
    return torch.empty(max(int(L.load().durf_mlp_saved_bytes(C.byref(t), L.PREC_BF16, M, N)), 16), device=dev, dtype=torch.uint8)


def mlp_bwd(topo, features, cond, blob, saved, d_raw_rgb, d_raw_density, d_blob, *, M: int, N: int, ray_index=None,
            want_d_features=False, precision=L.PREC_FP32, packed=None, count=None):
    """Accumulates into d_blob; returns d_features [M*N, in_dim] (fp32 path only) or None."""
    t = topology(topo)
    dev = _dev(features)
    lib = L.load()
    nbytes = int(lib.durf_mlp_workspace_bytes(C.byref(t), precision, M, N, 1))
    ws = torch.empty(max(nbytes, 16), device=dev, dtype=torch.uint8)
    if want_d_features and precision != L.PREC_FP32 and t.width != 128:
        raise L.DurfError("the tensor-core backward produces the input gradient for width-128 networks (BoxMLP) only; "
                          "use precision='fp32'")
    dfeat = torch.empty(M * N, t.in_dim, device=dev) if want_d_features else None
    a = _mlp_args(t, precision, M, N, features, f32(cond), f32(blob), packed, ray_index, count, False, d_raw_rgb, d_raw_density,
                  saved, ws)
    _poison(ws)
    cap = int(os.environ.get('DURF_WGRAD_CTAS', '0'))
    if cap > 0 and precision == L.PREC_BF16:
        # diagnostic: the serial backward with the weight-gradient kernel on `cap` CTAs (the partition of the tiles over
        # accumulators that OverlappedBackward uses), so that the two can be compared reduction for reduction
        g_rgb, g_den = f32(d_raw_rgb), f32(d_raw_density)
        check(lib.durf_mlp_bwd_data(stream_ptr(), C.byref(a), ptr(g_rgb), ptr(g_den), ptr(dfeat), None, 0), "durf_mlp_bwd_data")
        check(lib.durf_mlp_bwd_weights(stream_ptr(), C.byref(a), ptr(g_rgb), ptr(g_den), ptr(d_blob), None, cap), "durf_mlp_bwd_weights")
        return dfeat
    check(lib.durf_mlp_bwd(stream_ptr(), C.byref(a), ptr(f32(d_raw_rgb)), ptr(f32(d_raw_density)), ptr(d_blob), ptr(dfeat)),
          "durf_mlp_bwd")
    return dfeat


def _poison(ws: torch.Tensor) -> None:
    """DURF_BWD_POISON=1 (tests): fill the dZ workspace with NaN patterns, so that a block consumed before it was produced
    shows up as a non-finite gradient instead of as the (identical) stale data of the previous call."""
    if os.environ.get('DURF_BWD_POISON', '0') == '1':
        ws.fill_(0xFF)


class OverlappedBackward:
    """The tensor-core backward of one large network with its two kernels running CONCURRENTLY: the dZ chain
    (durf_mlp_bwd_data) on the current stream with `data_ctas` CTAs, the weight gradients (durf_mlp_bwd_weights) on a side
    stream with the remaining SMs, consuming every dZ block out of L2 as soon as the chain has published it
    (tile_done counters, include/durf_b200.h).  Several calls may be queued (one per level); `join()` makes the current
    stream wait for the side stream and releases the buffers that were kept alive for it."""

    _side = {}

    def __init__(self, data_ctas: Optional[int] = None):
        dev = torch.cuda.current_device()
        if dev not in OverlappedBackward._side:
            OverlappedBackward._side[dev] = torch.cuda.Stream(device=dev)
        self.side = OverlappedBackward._side[dev]
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        if data_ctas is None:
            data_ctas = int(os.environ.get('DURF_BWD_DATA_CTAS', str(round(sms * 0.65))))
        self.data_ctas = max(1, min(sms - 1, data_ctas))
        self.weight_ctas = sms - self.data_ctas
        self.keep = []

    def __call__(self, topo, features, cond, blob, saved, d_raw_rgb, d_raw_density, d_blob, *, M: int, N: int, packed):
        t = topology(topo)
        dev = _dev(features)
        lib = L.load()
        prec = L.PREC_BF16
        ws = torch.empty(max(int(lib.durf_mlp_workspace_bytes(C.byref(t), prec, M, N, 1)), 16), device=dev, dtype=torch.uint8)
        flags = torch.zeros(max(int(lib.durf_mlp_bwd_flags_bytes(C.byref(t), M)) // 4, 1), device=dev, dtype=torch.int32)
        _poison(ws)
        g_rgb, g_den = f32(d_raw_rgb), f32(d_raw_density)
        a = _mlp_args(t, prec, M, N, features, f32(cond), f32(blob), packed, None, None, False, g_rgb, g_den, saved, ws)
        cur = torch.cuda.current_stream()
        self.side.wait_stream(cur)                   # the counters are zero, the upstream gradients written
        check(lib.durf_mlp_bwd_data(cur.cuda_stream, C.byref(a), ptr(g_rgb), ptr(g_den), None, ptr(flags), self.data_ctas),
              "durf_mlp_bwd_data")
        if os.environ.get('DURF_BWD_SERIALIZE', '0') == '1':      # diagnosis: same launches, one stream
            check(lib.durf_mlp_bwd_weights(cur.cuda_stream, C.byref(a), ptr(g_rgb), ptr(g_den), ptr(d_blob), ptr(flags),
                                           self.weight_ctas), "durf_mlp_bwd_weights")
        else:
            with torch.cuda.stream(self.side):
                check(lib.durf_mlp_bwd_weights(self.side.cuda_stream, C.byref(a), ptr(g_rgb), ptr(g_den), ptr(d_blob), ptr(flags),
                                               self.weight_ctas), "durf_mlp_bwd_weights")
        self.keep.append((ws, flags, g_rgb, g_den, a))

    def join(self) -> None:
        torch.cuda.current_stream().wait_stream(self.side)
        self.keep.clear()


# ---- K3 ---------------------------------------------------------------------------------------------
def _comp_args(raw_rgb, raw_density, t_vals, dirs, white_bkgd, rand_bkgd, activated, density_bias, outs):
    B, N = raw_density.shape[0], raw_density.shape[1]
    return L.CompositeArgs(B=B, N=N, white_bkgd=int(white_bkgd), rand_bkgd=int(rand_bkgd), activated=int(activated),
                           density_bias=float(density_bias), raw_rgb=ptr(raw_rgb), raw_density=ptr(raw_density),
                           t_vals=ptr(t_vals), dirs=ptr(dirs), comp_rgb=ptr(outs.get('comp_rgb')), depth=ptr(outs.get('depth')),
                           acc=ptr(outs.get('acc')), weights=ptr(outs.get('weights')), t_mids=ptr(outs.get('t_mids')),
                           t_dists=ptr(outs.get('t_dists')))


def composite(raw_rgb, raw_density, t_vals, dirs, *, white_bkgd=False, rand_bkgd=False, activated=False, density_bias=-1.0,
              want_mids=True):
    raw_rgb, raw_density, t_vals, dirs = f32(raw_rgb), f32(raw_density), f32(t_vals), f32(dirs)
    raw_density = raw_density.reshape(raw_density.shape[0], -1)
    B, N = raw_density.shape
    dev = _dev(raw_rgb)
    outs = dict(comp_rgb=torch.empty(B, 3, device=dev), depth=torch.empty(B, device=dev), acc=torch.empty(B, device=dev),
                weights=torch.empty(B, N, device=dev))
    if want_mids:
        outs['t_mids'] = torch.empty(B, N, device=dev)
        outs['t_dists'] = torch.empty(B, N, device=dev)
    a = _comp_args(raw_rgb, raw_density, t_vals, dirs, white_bkgd, rand_bkgd, activated, density_bias, outs)
    check(L.load().durf_composite_fwd(stream_ptr(), C.byref(a)), "durf_composite_fwd")
    return outs


def composite_bwd(raw_rgb, raw_density, t_vals, dirs, d_comp_rgb, d_depth, d_weights, *, d_acc=None, white_bkgd=False,
                  rand_bkgd=False, activated=False, density_bias=-1.0, want_d_dirs=False, out_rgb=None, out_density=None):
    raw_rgb, raw_density, t_vals, dirs = f32(raw_rgb), f32(raw_density), f32(t_vals), f32(dirs)
    raw_density = raw_density.reshape(raw_density.shape[0], -1)
    B, N = raw_density.shape
    dev = _dev(raw_rgb)
    g_rgb = out_rgb if out_rgb is not None else torch.empty(B, N, 3, device=dev)
    g_den = out_density if out_density is not None else torch.empty(B, N, device=dev)
    g_dirs = torch.empty(B, 3, device=dev) if want_d_dirs else None
    a = _comp_args(raw_rgb, raw_density, t_vals, dirs, white_bkgd, rand_bkgd, activated, density_bias, {})
    check(L.load().durf_composite_bwd(stream_ptr(), C.byref(a), ptr(f32(d_comp_rgb)), ptr(f32(d_depth)), ptr(d_acc),
                                      ptr(f32(d_weights)), ptr(g_rgb), ptr(g_den), ptr(g_dirs)), "durf_composite_bwd")
    return g_rgb, g_den, g_dirs


# ---- K4 ---------------------------------------------------------------------------------------------
def resample(t_vals, weights, *, u_rand=None, padding=0.01, blurpool=True, num_samples=None):
    t_vals, weights = f32(t_vals), f32(weights)
    B, N = weights.shape
    S = N + 1 if num_samples is None else num_samples
    out = torch.empty(B, S, device=_dev(t_vals))
    check(L.load().durf_resample_fwd(stream_ptr(), B, N, ptr(t_vals), ptr(weights), ptr(None if u_rand is None else f32(u_rand)),
                                     float(padding), int(blurpool), S, ptr(out)), "durf_resample_fwd")
    return out


# ---- KL / KA ----------------------------------------------------------------------------------------
def losses_reduce_ws(B: int, device) -> torch.Tensor:
    """Zero-initialised workspace of the deterministic loss reduction (allocate once, reuse every step)."""
    return torch.zeros(int(L.load().durf_losses_reduce_ws_floats(B)), device=device)


def loss_args(level, num_levels, eps, cfg, lv, batch, depth_mask, partials, grads, reduce_ws):
    B, N = lv['weights'].shape
    g = grads
    eps_dev = eps if torch.is_tensor(eps) else None
    return L.LossArgs(B=B, N=N, level=level, num_levels=num_levels, eps=0.0 if eps_dev is not None else float(eps),
                      reduce_ws=ptr(reduce_ws), eps_dev=ptr(eps_dev),
                      coarse_loss_mult=cfg.coarse_loss_mult, box_loss_mult=cfg.box_loss_mult,
                      depth_loss_mult=cfg.depth_loss_mult, near_loss_mult=cfg.near_loss_mult,
                      empty_loss_mult=cfg.empty_loss_mult, sky_loss_mult=cfg.sky_loss_mult, distortion_mult=1e-6,
                      comp_rgb=ptr(lv['comp_rgb']), depth=ptr(lv['depth']), weights=ptr(lv['weights']), t_vals=ptr(lv['t_vals']),
                      pixels=ptr(batch['pixels']), depth_gt=ptr(batch['depth']), sky=ptr(batch['sky']),
                      lossmult=ptr(batch['lossmult']), dyn_mask=ptr(batch['dyn_mask']), zo=ptr(batch['zo']),
                      depth_mask=ptr(depth_mask), partials=ptr(partials), d_comp_rgb=ptr(g.get('comp_rgb')),
                      d_depth=ptr(g.get('depth')), d_weights=ptr(g.get('weights')))


def losses_prepare(args: L.LossArgs, norms: torch.Tensor):
    check(L.load().durf_losses_prepare(stream_ptr(), C.byref(args), ptr(norms)), "durf_losses_prepare")


def losses_fwd_bwd(args: L.LossArgs, norms: torch.Tensor):
    check(L.load().durf_losses_fwd_bwd(stream_ptr(), C.byref(args), ptr(norms)), "durf_losses_fwd_bwd")


def grad_sanitize(grad: torch.Tensor, max_val: float, scale: float, sumsq: torch.Tensor):
    check(L.load().durf_grad_sanitize(stream_ptr(), grad.numel(), ptr(grad), float(max_val), float(scale), ptr(sumsq)),
          "durf_grad_sanitize")


def adam_step(params, grad, m, v, sumsq, *, max_norm, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, advance_step=True):
    """flax.optim.Adam.  `lr` / `step` are Python numbers, or device tensors (float32[1] / int32[1]) for a captured graph;
    in the device form the kernel increments the step counter itself when advance_step."""
    if torch.is_tensor(lr) or torch.is_tensor(step):
        check(L.load().durf_adam_step_dev(stream_ptr(), params.numel(), ptr(params), ptr(grad), ptr(m), ptr(v), ptr(sumsq),
                                          float(max_norm), ptr(lr), ptr(step), int(advance_step), beta1, beta2, eps),
              "durf_adam_step_dev")
        return
    check(L.load().durf_adam_step(stream_ptr(), params.numel(), ptr(params), ptr(grad), ptr(m), ptr(v), ptr(sumsq),
                                  float(max_norm), float(lr), beta1, beta2, eps, int(step)), "durf_adam_step")
