"""Mirror of the reference's internal/box_helpers.py functions the model calls."""
import torch

from . import ops


def aa2matrix(angles):
    """box_helpers.py:148-167: axis-angle [K,3] -> rotation [K,3,3]."""
    return ops.aa2matrix(angles)


def world2object_rpy(pts, dirs, pose, rot, dim=None, inverse=False, *, angles=None):
    """box_helpers.py:286-341 (dim=None, inverse=False).  The CUDA front-end takes the axis-angle box parameters
    (it forms the rotation itself), so pass `angles=[K,3]`; `pose` is [B,K,3] or [K,3] (rows are identical per ray,
    obbpose_model.py:99).  Returns [pts_o, dirs_o], each [B,K,3]."""
    if inverse or dim is not None or angles is None:
        raise NotImplementedError("only the forward, unscaled transform used by the model is built (pass angles=)")
    p = pose[0] if pose.dim() == 3 else pose
    box = torch.cat([ops.f32(p), ops.f32(angles)], dim=-1).contiguous()
    ext = torch.ones(box.shape[0], 3, device=box.device)
    o = ops.obb_frontend(pts, dirs, box, ext, want_object_rays=True)
    return [o['origins_o'], o['dirs_o']]


def ray_box_intersection_world(origins, dirs, box, ext):
    """world2object_rpy + ray_box_intersection (box_helpers.py:59-106) against [-ext, +ext] in one launch:
    -> (z_in [B,K], z_out [B,K], intersection [B,K] int32)."""
    o = ops.obb_frontend(origins, dirs, box, ext)
    return o['zi'], o['zo'], o['hit']
