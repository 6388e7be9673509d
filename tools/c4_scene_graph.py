#!/usr/bin/env python
"""BASELINE.json configs[3] (C4): dynamic scene graph -- background + 8 per-object NeRFs, per-ray OBB intersection and
merged compositing over Waymo-shape frames.  Times one 1920x1280 frame (rays generated on the device) and reports the share
of rays that hit a box (parity of this path: tests/test_gpu_model.py).  Usage: python tools/c4_scene_graph.py [rows]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from durf_b200 import synthetic as S  # noqa: E402
from durf_b200.obbpose_model import MipNerfModel, Variables, render_camera  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
    K = 8
    rng = np.random.default_rng(S.SEED)
    c2w = S.random_c2w(rng)
    centers, ext = S.boxes_in_view(rng, c2w, K)
    model = MipNerfModel(num_objects=K, precision='bf16')
    v = Variables.allocate(model, K, centers.shape[0], 'cuda')
    v.load_mlp('MLP_0', S.glorot_mlp(rng, 60, 256, 0.0))
    for k in range(K):
        v.load_mlp(f'BoxMLP_{k}', S.glorot_mlp(rng, 63, 128, 0.0))
    v.box_centers.copy_(torch.from_numpy(centers).cuda())
    v.mark_dirty()
    ext_t = torch.from_numpy(ext).cuda()
    fn = lambda r, b: model.apply(v, r, b['rays'], None, b['ext'], b['ts'], False, False, False, b['alpha'])
    W = S.WAYMO_W
    for _ in range(2):
        out = render_camera(fn, c2w, W, rows, S.FOCAL, 0.0, 40.0, None, ext_t, 0, None, 10.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_it = 3
    for _ in range(n_it):
        out = render_camera(fn, c2w, W, rows, S.FOCAL, 0.0, 40.0, None, ext_t, 0, None, 10.0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n_it
    n = rows * W
    from durf_b200 import ops
    fe = ops.obb_frontend(*[t for t in ops.generate_rays(c2w, W, rows, S.FOCAL, 0.0, 40.0)[:2]], v.box_centers[0].contiguous(), ext_t)
    hit = float((fe['nhit'] > 0).float().mean())
    print(f"C4: {n} rays, K={K} objects, {100 * hit:.1f}% of rays hit a box: {ms:.1f} ms/frame = {n / ms * 1e3 / 1e6:.3f} M rays/s "
          f"(finite: {bool(torch.isfinite(out[0]).all())})")


if __name__ == "__main__":
    main()
