"""A synthetic on-disk CARLA (or, with waymo=True, Waymo) scene in the layout internal/obbpose_dataset.py reads (test infrastructure): 3 timesteps x 5
cameras of 12 x 16 pixels at factor 4, two cars, LIDAR depth with holes, sky masks, instance masks.  File names are
zero-padded so that natural and lexicographic order agree (the test-only natsort stand-in sorts lexicographically)."""
import os

import numpy as np


def make_scene(root: str, T: int = 3, cams: int = 5, H: int = 12, W: int = 16, factor: int = 4, seed: int = 5,
               waymo: bool = False) -> str:
    from PIL import Image
    rng = np.random.default_rng(seed)
    n = T * cams
    os.makedirs(os.path.join(root, f'images_{factor}'), exist_ok=True)
    for i in range(n):
        img = (rng.uniform(size=(H, W, 4)) * 255).astype(np.uint8)
        Image.fromarray(img, 'RGBA').save(os.path.join(root, f'images_{factor}', f'img_{i:03d}.png'))
    poses = []
    for i in range(n):
        ang = rng.uniform(-0.6, 0.6)
        c, s = np.cos(ang), np.sin(ang)
        rot = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]) @ np.diag([1, -1, -1.0])
        t = rng.uniform(-20, 20, size=3) + np.array([0, 0, 5.0 * (i // cams)])
        hwf = np.array([H * factor, W * factor, 50.0 * factor])
        p = np.concatenate([rot, t[:, None], hwf[:, None]], 1)
        row = np.concatenate([p.reshape(-1), [1.0, 400.0]])
        if waymo:      # two more columns: the principal point (cx, cy) in full-resolution pixels (obbpose_dataset.py:1637)
            row = np.concatenate([row, [W * factor * 0.5 + rng.uniform(-3, 3), H * factor * 0.5 + rng.uniform(-3, 3)]])
        poses.append(row)
    np.save(os.path.join(root, 'poses_bounds.npy'), np.array(poses))
    boxes = {}
    for ts in range(1, T + 1):
        for car in (1, 2):
            ang = rng.uniform(-3, 3)
            c, s = np.cos(ang), np.sin(ang)
            m = np.eye(4)
            m[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
            m[:3, 3] = rng.uniform(-30, 30, size=3)
            boxes[f'{ts}_{car}_center'] = m
            boxes[f'{ts}_{car}_ext'] = rng.uniform(4, 12, size=3)
    np.save(os.path.join(root, '3D_boxes.npy'), boxes, allow_pickle=True)
    depth = rng.uniform(0, 300, size=(n, H, W)).astype(np.float32)
    depth[rng.uniform(size=depth.shape) < 0.2] = 0.0
    np.savez(os.path.join(root, 'depth_images.npz'), depth)
    np.savez(os.path.join(root, 'sky_masks.npz'), (rng.uniform(size=(n, H, W)) < 0.15).astype(np.float32))
    np.savez(os.path.join(root, '2D_boxes.npz'), rng.integers(0, 3, size=(n, H, W)).astype(np.int32))
    return root
