"""Synthetic inputs with the reference's shapes: pinhole rays, boxes, targets, weights.

Pure numpy, no device work.  Everything is seeded with numpy.random.default_rng(20200823)
(the reference seeds its PRNG with 20200823, train_boxpose.py:325).

Ray construction follows the reference loader `_generate_rays_multi`
(internal/obbpose_dataset.py:613-661): camera_dirs = ((x - W/2)/f, -(y - H/2)/f, -1),
directions = c2w[:3,:3] @ camera_dirs (un-normalised), viewdirs = directions/|directions|,
radii = |dir(y) - dir(y+1)| * 2/sqrt(12), lossmult = 1, near/far constants.
"""
from __future__ import annotations

import math
from typing import Dict, List, NamedTuple, Optional, Tuple

import numpy as np

SEED = 20200823
WAYMO_W, WAYMO_H = 1920, 1280
FOCAL = 2058.72  # 4 x 514.68 (carla/carla_data.ipynb cell 8), full-resolution frame


class RayBatch(NamedTuple):
    """Field order of the reference's BoxRays namedtuple (internal/utils.py:84-86)."""
    origins: np.ndarray     # [B,3]
    directions: np.ndarray  # [B,3] un-normalised
    viewdirs: np.ndarray    # [B,3] unit
    radii: np.ndarray       # [B,1]
    lossmult: np.ndarray    # [B,1]
    near: np.ndarray        # [B,1]
    far: np.ndarray         # [B,1]


def random_c2w(rng: np.random.Generator) -> np.ndarray:
    """Random rigid camera-to-world [3,4]: rotation from a QR of a Gaussian matrix, translation U(-1,1)^3
    (the reference pre-scales scenes by 1/5, obbpose_dataset.py:437)."""
    q, r = np.linalg.qr(rng.standard_normal((3, 3)))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    t = rng.uniform(-1.0, 1.0, size=(3, 1))
    return np.concatenate([q, t], axis=1).astype(np.float32)


def pixel_rays(c2w: np.ndarray, px: np.ndarray, py: np.ndarray, width: int = WAYMO_W, height: int = WAYMO_H,
               focal: float = FOCAL, near: float = 0.0, far: float = 200.0) -> RayBatch:
    """Rays through pixel centres (px, py) (float32 pixel indices, any shape [B])."""
    px = px.astype(np.float32)
    py = py.astype(np.float32)
    f = np.float32(focal)

    def world_dir(x, y):
        cam = np.stack([(x - np.float32(width * 0.5)) / f, -(y - np.float32(height * 0.5)) / f,
                        -np.ones_like(x)], axis=-1)
        return (cam[..., None, :] * c2w[:3, :3]).sum(axis=-1).astype(np.float32)

    d = world_dir(px, py)
    # neighbour along y; the last row gets v[-2:-1] of the (H-1)-row difference array, i.e. |dir(H-3) - dir(H-2)|
    # (obbpose_dataset.py:640-643)
    y0 = np.where(py >= height - 1, py - 2, py)
    dn = np.sqrt(((world_dir(px, y0) - world_dir(px, y0 + 1)) ** 2).sum(-1))
    radii = ((dn[..., None] * np.float32(2)) / np.float32(np.sqrt(12))).astype(np.float32)   # float32 like the reference's numpy 1.x
    o = np.broadcast_to(c2w[:3, -1], d.shape).astype(np.float32).copy()
    v = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    ones = np.ones_like(radii)
    return RayBatch(o, d, v, radii, ones, (near * ones).astype(np.float32), (far * ones).astype(np.float32))


def frame_rays(c2w: np.ndarray, width: int = WAYMO_W, height: int = WAYMO_H, focal: float = FOCAL,
               near: float = 0.0, far: float = 200.0, row0: int = 0, row1: Optional[int] = None) -> RayBatch:
    """All rays of rows [row0,row1) of one camera, row-major (the order render_image flattens to)."""
    row1 = height if row1 is None else row1
    xs, ys = np.meshgrid(np.arange(width, dtype=np.float32), np.arange(row0, row1, dtype=np.float32), indexing='xy')
    return pixel_rays(c2w, xs.reshape(-1), ys.reshape(-1), width, height, focal, near, far)


def random_rays(rng: np.random.Generator, n: int, c2w: Optional[np.ndarray] = None, near: float = 0.0,
                far: float = 200.0) -> Tuple[RayBatch, np.ndarray]:
    """n random pixels of one camera (BASELINE.json configs[0] / configs[2] batches)."""
    c2w = random_c2w(rng) if c2w is None else c2w
    px = rng.integers(0, WAYMO_W, size=n).astype(np.float32)
    py = rng.integers(0, WAYMO_H, size=n).astype(np.float32)
    return pixel_rays(c2w, px, py, near=near, far=far), c2w


def boxes_in_view(rng: np.random.Generator, c2w: np.ndarray, num_objects: int, timesteps: int = 5,
                  behind: bool = False, width: int = WAYMO_W, height: int = WAYMO_H,
                  focal: float = FOCAL) -> Tuple[np.ndarray, np.ndarray]:
    """box_centers [T,K,6] (xyz + axis-angle, world->object; obbpose_model.py:88) and half-extents ext [K,3].

    Boxes sit 4-6 units in front of the camera (or behind it: `behind`=True -> no ray hits, config C1), one
    per cell of a grid over the image so that no ray crosses two boxes: the reference sums multiple hits
    (obbpose_model.py:120-122 "assumes that objects do not occlude each other") and such rays come out NaN."""
    cols = int(math.ceil(math.sqrt(num_objects * 1.5)))
    rows = int(math.ceil(num_objects / cols))
    near_dist = 4.0
    slot_w = 2 * (width / 2 / focal) * near_dist / cols
    slot_h = 2 * (height / 2 / focal) * near_dist / rows
    ext_max = min(0.5, 0.4 * min(slot_w, slot_h) / math.sqrt(3.0))
    centers = np.zeros((timesteps, num_objects, 6), np.float32)
    fwd = -c2w[:3, 2]
    right, up = c2w[:3, 0], c2w[:3, 1]
    for k in range(num_objects):
        dist = rng.uniform(near_dist, 6.0)
        cx = ((k % cols) + 0.5) / cols - 0.5 + rng.uniform(-0.05, 0.05) / cols
        cy = ((k // cols) + 0.5) / rows - 0.5 + rng.uniform(-0.05, 0.05) / rows
        lateral = cx * 2 * (width / 2 / focal) * dist
        vertical = cy * 2 * (height / 2 / focal) * dist
        base = c2w[:3, 3] + (fwd * dist + right * lateral + up * vertical) * (-1.0 if behind else 1.0)
        aa = rng.uniform(-0.6, 0.6, size=3)
        for t in range(timesteps):
            centers[t, k, :3] = base + rng.uniform(-0.02, 0.02, size=3)
            centers[t, k, 3:] = aa + rng.uniform(-0.02, 0.02, size=3)
    ext = (rng.uniform(0.5, 1.0, size=(num_objects, 3)) * ext_max).astype(np.float32)
    return centers, ext


def targets(rng: np.random.Generator, n: int) -> Dict[str, np.ndarray]:
    """pixels U(0,1); depth U(1,40) on 60 % of rays else 0; sky 0.995 on 15 % of rays."""
    pixels = rng.uniform(0.0, 1.0, size=(n, 3)).astype(np.float32)
    depth = np.where(rng.uniform(size=(n, 1)) < 0.6, rng.uniform(1.0, 40.0, size=(n, 1)), 0.0).astype(np.float32)
    sky = np.where(rng.uniform(size=(n, 1)) < 0.15, 0.995, 0.0).astype(np.float32)
    return dict(pixels=pixels, depth=depth, sky=sky)


def layer_shapes(in_dim: int, width: int, depth: int = 8, skip: int = 4, cond_dim: int = 27,
                 cond_width: int = 128) -> List[Tuple[int, int]]:
    """[in,out] of Dense_0..Dense_{depth+3} in flax creation order (obbpose_model.py:329-353):
    trunk (skip concat after layer `skip`), density, bottleneck, condition, rgb."""
    shapes, k = [], in_dim
    for i in range(depth):
        shapes.append((k, width))
        k = width + in_dim if (i % skip == 0 and i > 0) else width
    shapes += [(k, 1), (k, width), (width + cond_dim, cond_width), (cond_width, 3)]
    return shapes


def glorot_mlp(rng: np.random.Generator, in_dim: int, width: int, bias_scale: float = 0.0, **kw) -> List[Tuple[np.ndarray, np.ndarray]]:
    """Random-init weights of the reference architecture: glorot-uniform kernels [in,out], zero biases
    (obbpose_model.py:326-327).  bias_scale>0 draws small biases so tests exercise the bias path."""
    out = []
    for fi, fo in layer_shapes(in_dim, width, **kw):
        lim = math.sqrt(6.0 / (fi + fo))
        out.append((rng.uniform(-lim, lim, size=(fi, fo)).astype(np.float32),
                    (rng.uniform(-1, 1, size=(fo,)) * bias_scale).astype(np.float32)))
    return out
