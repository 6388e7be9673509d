"""Mirror of the reference's internal/box_helpers.py functions the model calls (same names, argument order and shapes).

The model itself uses the fused front-end (`ops.obb_frontend`: Rodrigues + transform + slab test + scene-graph merge in
one launch); the functions here expose the reference's individual steps on the device."""
import torch

from . import _lib as L
from . import ops


def aa2matrix(angles):
    """box_helpers.py:148-167: axis-angle [K,3] -> rotation [K,3,3]."""
    return ops.aa2matrix(angles)


def _per_ray(t: torch.Tensor, per_object_dims: int):
    """(contiguous tensor, per_ray flag): [K,..] and broadcast views of it ([B,K,..] with stride 0) are per object."""
    if t.dim() == per_object_dims:
        return ops.f32(t), 0
    if t.stride(0) == 0:
        return ops.f32(t[0]), 0
    return ops.f32(t), 1


def world2object_rpy(pts, dirs, pose, rot, dim=None, inverse=False):
    """box_helpers.py:286-341 as the model calls it (`dim=None, inverse=False`, obbpose_model.py:110):
    pts, dirs [B,3]; pose [B,K,3]; rot [B,K,3,3] (rotation matrices, e.g. `jnp.broadcast_to(aa2matrix(..), [B,K,3,3])`)
    -> [pts_o, dirs_o], each [B,K,3], dirs_o normalised."""
    if inverse or dim is not None:
        raise NotImplementedError("world2object_rpy: only the forward, unscaled transform the model uses is built "
                                  "(the reference's scale_frames / inverse branches are dead code on the hot path)")
    pts, dirs = ops.f32(pts), ops.f32(dirs)
    B = pts.shape[0]
    pose_c, pose_pr = _per_ray(pose, 2)
    rot_c, rot_pr = _per_ray(rot, 3)
    K = rot_c.shape[-3]
    pts_o = torch.empty(B, K, 3, device=pts.device)
    dirs_o = torch.empty(B, K, 3, device=pts.device)
    L.check(L.load().durf_world2object_fwd(L.stream_ptr(), B, K, L.ptr(pts), L.ptr(dirs), L.ptr(pose_c), pose_pr, L.ptr(rot_c),
                                           rot_pr, L.ptr(pts_o), L.ptr(dirs_o)), "durf_world2object_fwd")
    return [pts_o, dirs_o]


def ray_box_intersection(ray_o, ray_d, aabb_min=None, aabb_max=None):
    """box_helpers.py:59-106: ray_o, ray_d [B,K,3] (object frames), bounds [B,K,3] or None (unit box)
    -> (z_ray_in [B,K], z_ray_out [B,K], intersection_map [B,K] int32)."""
    ray_o, ray_d = ops.f32(ray_o), ops.f32(ray_d)
    shape = ray_o.shape[:-1]
    n = ray_o.numel() // 3
    if n == 0:
        return None, None, None                                                   # box_helpers.py:103-104
    mn = None if aabb_min is None else ops.f32(aabb_min.expand_as(ray_o))
    mx = None if aabb_max is None else ops.f32(aabb_max.expand_as(ray_o))
    zi = torch.empty(shape, device=ray_o.device)
    zo = torch.empty(shape, device=ray_o.device)
    hit = torch.empty(shape, device=ray_o.device, dtype=torch.int32)
    L.check(L.load().durf_ray_box_intersection_fwd(L.stream_ptr(), n, L.ptr(ray_o), L.ptr(ray_d), L.ptr(mn), L.ptr(mx), L.ptr(zi),
                                                   L.ptr(zo), L.ptr(hit)), "durf_ray_box_intersection_fwd")
    return zi, zo, hit


def ray_box_intersection_world(origins, dirs, box, ext):
    """world2object_rpy + ray_box_intersection against [-ext, +ext] in one launch (the fused front-end):
    -> (z_in [B,K], z_out [B,K], intersection [B,K] int32)."""
    o = ops.obb_frontend(origins, dirs, box, ext)
    return o['zi'], o['zo'], o['hit']
