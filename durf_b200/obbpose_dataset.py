"""Host-side mirror of the reference's three loaders (internal/obbpose_dataset.py:44-832 class `Carla`, the loader
`configs/carla_dyn.gin` selects; :833-1460 `Carla_Seq`; :1461-2087 `Waymo`, `configs/waymo.gin`): the on-disk scene -> the
batch dicts `train_step` / `render_image` consume.

    <data_dir>/images_<factor>/*.png|jpg   RGB(A) frames, 5 cameras per timestep, natural file order
    <data_dir>/poses_bounds.npy            [n, 17]: 3x5 LLFF pose (rotation | translation | h, w, focal) + 2 depth bounds
    <data_dir>/3D_boxes.npy                dict '<ts>_<car>_center' -> 4x4 box pose, '<ts>_<car>_ext' -> half extents
    <data_dir>/depth_images.npz, sky_masks.npz, 2D_boxes.npz     'arr_0': [n, h, w] LIDAR depth / sky mask / instance ids

What is kept exactly (checked against the reference's own loader executed on a synthetic scene, tests/test_dataset.py):
the recentring of camera and box poses by the average camera (`_recenter_poses`, :709-738), the scene scale 1/5, the
world-to-object axis-angle of every box (`scipy Rotation.from_matrix(inv(R)).as_rotvec()`), the optional box / yaw noise
and its draw order, depth / 5 on valid pixels, sky -> 0.995, the hard-coded hold-out `i_test = [10, 11]` (:538), the
pinhole rays of `_generate_rays_multi` (:613-661: pixel centres at integer coordinates, radii from the x-neighbour distance),
'timestep' batching (all cameras of one timestep pooled, :235-262) and the seeded draw sequence of `_next_train`
(np.random.seed(20201473) at :208, then one timestep draw and one ray-index draw per batch).

What differs: no loader thread and no global RNG - batches are produced on demand from a private `RandomState` with the
reference's seed (same batches as the reference's thread, which is the only user of the global generator there), `peek()`
returns the next batch without consuming it; the NDC branch (`spherify = False`) is refused: in the reference it builds
`utils.Rays` with 7 of its 8 fields (:693-700) and cannot run.  Rays can also be generated on the device
(`durf_generate_rays`, bit-exact with this function) when a whole camera is rendered: `camera(i)` returns what
`obbpose_model.render_camera` needs.
"""
from __future__ import annotations

import os
import re
from typing import Dict, List

import numpy as np

from .utils import Config, Rays


def _natural_key(s: str):
    return [int(t) if t.isdigit() else t for t in re.split(r'(\d+)', s)]


def _load_npz(path: str):
    with open(path, 'rb') as fp:
        return np.load(fp, allow_pickle=True)['arr_0']


class Carla:
    """`Carla(split, data_dir, config)`; iterate for batches.  split: 'train' | 'test' | 'render'."""

    CAMERAS_PER_TIMESTEP = 5          # FRONT, FRONT_LEFT, SIDE_LEFT, FRONT_RIGHT, SIDE_RIGHT (:515)
    SCENE_SCALE = 5.0                 # far plane 1000 -> 200 (:446)
    SEED = 20201473                   # :208
    I_TEST = (10, 11)                 # the hard-coded hold-out (:538)
    EXT_DIVISOR = 5.0                 # box extents are half extents in scene units (:474)
    SKY_VALUE = 0.995                 # "large distance but not infinity" (:596)
    RENDER_SPLIT_IS_TRAIN = True      # split == 'render' uses the training frames (:544)
    POSE_COLUMNS = 17                 # 15 pose + 2 depth bounds

    def __init__(self, split: str, data_dir: str, config: Config):
        if split not in ('train', 'test', 'render'):
            raise ValueError("the split argument should be either 'train' or 'test', set to {} here.".format(split))
        self.split, self.data_dir = split, data_dir
        self.near, self.far = config.near, config.far
        self.batch_size, self.batching, self.render_path = config.batch_size, config.batching, config.render_path
        self._rng = np.random.RandomState(self.SEED)
        self._load_renderings(config)
        self._generate_rays()
        self.it = 0
        self._peeked = None
        if split == 'train':
            if config.batching != 'timestep':
                raise NotImplementedError(f"{config.batching} batching: configs/carla_dyn.gin uses 'timestep'")
            for name in ('images', 'depth', 'sky_mask', 'masks2d'):
                setattr(self, name, self._flatten_time(getattr(self, name)))
            self.rays = Rays(*[self._flatten_time(r) for r in self.rays])

    # -- iteration --------------------------------------------------------------------------------------
    def __iter__(self):
        return self

    def __next__(self) -> Dict:
        if self._peeked is not None:
            b, self._peeked = self._peeked, None
            return b
        return self._next_train() if self.split == 'train' else self._next_test()

    def peek(self) -> Dict:
        if self._peeked is None:
            self._peeked = self._next_train() if self.split == 'train' else self._next_test()
        return self._peeked

    @property
    def size(self) -> int:
        return self.n_examples

    # -- loading ------------------------------------------------------------------------------------------
    def _load_renderings(self, config: Config) -> None:
        factor = config.factor if config.factor > 0 else 1
        imgdir = os.path.join(self.data_dir, 'images' + ('_{}'.format(config.factor) if config.factor > 0 else ''))
        if not os.path.exists(imgdir):
            raise ValueError('Image folder {} does not exist.'.format(imgdir))
        from PIL import Image
        files = sorted((f for f in os.listdir(imgdir) if f.endswith(('JPG', 'jpg', 'png'))), key=_natural_key)
        images = np.array([np.array(Image.open(os.path.join(imgdir, f)), dtype=np.float32)[:, :, :3] / 255. for f in files])

        poses_arr = np.load(os.path.join(self.data_dir, 'poses_bounds.npy'))
        poses = poses_arr[:, :15].reshape([-1, 3, 5]).transpose([1, 2, 0])
        bds = poses_arr[:, 15:17].transpose([1, 0])
        principal_point = poses_arr[:, 17:] * 1. / factor if self.POSE_COLUMNS > 17 else None      # Waymo: (cx, cy) per frame
        if poses.shape[-1] != len(images):
            raise RuntimeError('Mismatch between imgs {} and poses {}'.format(len(images), poses.shape[-1]))
        masks3d = np.load(os.path.join(self.data_dir, '3D_boxes.npy'), allow_pickle=True).item()
        centers = [k for k in masks3d if 'center' in k]
        box_pose = np.array([masks3d[k] for k in centers])
        box_ext = np.array([masks3d[k] for k in masks3d if 'ext' in k])

        poses[:2, 4, :] = np.floor(poses[:2, 4, :] * 1. / factor)           # h, w of the down-sampled frames
        poses[2, 4, :] = poses[2, 4, :] * 1. / factor                       # focal
        poses = np.moveaxis(poses, -1, 0).astype(np.float32)
        self.bds = np.moveaxis(bds, -1, 0).astype(np.float32)

        self.random_box = False
        if config.centering:
            poses, c2w = self._recenter_poses(poses)
            poses[:, :3, 3] /= self.SCENE_SCALE
            c2w_inv = np.linalg.inv(c2w)
            random_box = None
            if config.random_box:
                self.random_box = True
                random_box = box_pose.copy()
                random_box[:, :3, 3] += self._rng.uniform(-config.box_noise, config.box_noise, size=[box_pose.shape[0], 3])
                random_box = c2w_inv @ random_box
                random_box[:, :3, 3] /= self.SCENE_SCALE
            box_pose = c2w_inv @ box_pose
            box_pose[:, :3, 3] /= self.SCENE_SCALE
            from scipy.spatial.transform import Rotation
            yaw = np.array(Rotation.from_matrix(np.linalg.inv(box_pose[:, :3, :3])).as_rotvec())    # world -> object
            if config.random_yaw and config.random_box:
                rand_yaw = yaw + self._rng.uniform(-config.yaw_noise, config.yaw_noise, size=yaw.shape) * (np.pi / 180.0)
                rand_pose = np.concatenate([random_box[:, :3, 3], rand_yaw], axis=-1)
            elif config.random_box:
                rand_pose = np.concatenate([random_box[:, :3, 3], yaw], axis=-1)
            else:
                rand_pose = np.concatenate([box_pose[:, :3, 3], yaw], axis=-1)
            obbpose = np.concatenate([box_pose[:, :3, 3], yaw], axis=-1)
            box_ext = box_ext / self.EXT_DIVISOR
        rel_pose, can_pose = {}, None
        for i, key in enumerate(centers):
            ts, car, _ = key.split('_')
            if '1_' in key:                              # the reference's test (:471): any key containing '1_'
                can_pose = box_pose[i]
                rel_pose[ts + '_' + car + '_rel'] = np.eye(4)
            else:
                rel_pose[ts + '_' + car + '_rel'] = np.matmul(can_pose, np.linalg.inv(box_pose[i]))
            if config.centering:
                masks3d[key] = obbpose[i]
                masks3d[ts + '_' + car + '_off'] = rand_pose[i]
                masks3d[ts + '_' + car + '_ext'] = box_ext[i]

        depth = _load_npz(os.path.join(self.data_dir, 'depth_images.npz'))
        sky = _load_npz(os.path.join(self.data_dir, 'sky_masks.npz'))
        masks2d = _load_npz(os.path.join(self.data_dir, '2D_boxes.npz'))
        for name, arr in (('depth', depth), ('depth', sky), ('masks2d', masks2d)):
            if len(arr) != len(images):
                raise RuntimeError('Mismatch between imgs {} and {} {}'.format(len(images), name, len(arr)))
        n_ts = int(len(masks2d) / self.CAMERAS_PER_TIMESTEP)
        timesteps = np.repeat(np.arange(1, n_ts + 1), self.CAMERAS_PER_TIMESTEP)
        self.total_timesteps = timesteps[-1]
        if not config.spherify:
            raise NotImplementedError("spherify = False: the reference's NDC branch builds utils.Rays with a missing field "
                                      "(obbpose_dataset.py:693-700) and cannot run; configs/carla_dyn.gin sets spherify = True")
        self.spherify = True

        i_train, i_test = self._hold_out(len(images), config)
        if self.split == 'test':
            indices = i_test
        elif self.split == 'render' and not self.RENDER_SPLIT_IS_TRAIN:
            indices = np.sort(np.concatenate([i_train, i_test]))
        else:
            indices = i_train
        images, depth, sky, poses, masks2d = images[indices], depth[indices], sky[indices], poses[indices], masks2d[indices]
        self.timesteps = timesteps[indices]
        self.rel_poses, self.box_pose = rel_pose, masks3d
        self.principal_point = None if principal_point is None else principal_point[indices]
        self.obj_ids = self._object_ids(masks2d)
        timestep_batches = config.batching == 'timestep'
        self.images = list(images)
        self.depth = [np.where(d > 0.0, d / self.SCENE_SCALE, d).astype(d.dtype) for d in depth]
        self.sky_mask = [np.where(s > 0.0, np.asarray(self.SKY_VALUE, s.dtype), s) for s in sky]
        self.masks2d = list(masks2d)
        if timestep_batches:
            self.depth = [d[..., None] for d in self.depth]
            self.sky_mask = [s[..., None] for s in self.sky_mask]
            self.masks2d = [m[..., None] for m in self.masks2d]
        self.camtoworlds = poses[:, :3, :4]
        self.focal, self.h, self.w = poses[:, -1, -1], poses[:, 0, -1], poses[:, 1, -1]
        self.resolution = self.h * self.w
        self.n_examples = len(self.images)

    def _hold_out(self, n: int, config: Config):
        """(training frames, held-out frames): frames 10 and 11 are held out (:538-540)."""
        i_test = np.array(self.I_TEST)
        return np.array([i for i in np.arange(n) if i not in i_test]), i_test

    def _object_ids(self, masks2d):
        """:567-574: the instance ids that occur in the selected frames' 2D masks, in order of first appearance."""
        ids: List = []
        for u in masks2d:
            for i in np.unique(u):
                if i != 0 and i not in ids:
                    ids.append(i)
        return np.array(ids)

    def _recenter_poses(self, poses):
        """:709-720 (the original NeRF recentring): world := the average camera's frame."""
        out = poses.copy()
        bottom = np.reshape([0, 0, 0, 1.], [1, 4])
        c2w = np.concatenate([self._poses_avg(poses)[:3, :4], bottom], -2)
        full = np.concatenate([poses[:, :3, :4], np.tile(bottom[None], [poses.shape[0], 1, 1])], -2)
        out[:, :3, :4] = (np.linalg.inv(c2w) @ full)[:, :3, :4]
        return out, c2w

    @staticmethod
    def _poses_avg(poses):
        norm = lambda x: x / np.linalg.norm(x)
        center = poses[:, :3, 3].mean(0)
        vec2 = norm(norm(poses[:, :3, 2].sum(0)))
        up = poses[:, :3, 1].sum(0)
        vec0 = norm(np.cross(up, vec2))
        vec1 = norm(np.cross(vec2, vec0))
        return np.concatenate([np.stack([vec0, vec1, vec2, center], 1), poses[0, :3, -1:]], 1)

    # -- rays ---------------------------------------------------------------------------------------------
    def _generate_rays(self) -> None:
        """`_generate_rays_multi` (:613-661): per image, pinhole rays through integer pixel coordinates."""
        fields = [[] for _ in range(7)]
        for i in range(len(self.images)):
            w, h, f = self.w[i], self.h[i], self.focal[i]
            x, y = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32), indexing='xy')
            cx, cy = (w * 0.5, h * 0.5) if self.principal_point is None else self.principal_point[i]
            cam_dirs = np.stack([(x - cx) / f, -(y - cy) / f, -np.ones_like(x)], axis=-1)
            directions = np.squeeze((cam_dirs[..., None, :] * self.camtoworlds[i, :3, :3]).sum(axis=-1))
            origins = np.broadcast_to(self.camtoworlds[i, :3, -1], directions.shape)
            viewdirs = directions / np.linalg.norm(directions, axis=-1, keepdims=True)
            dx = np.sqrt(np.sum((directions[:-1, :, :] - directions[1:, :, :]) ** 2, -1))
            dx = np.concatenate([dx, dx[-2:-1, :]], 0)
            radii = dx[..., None] * 2 / np.sqrt(12)
            ones = np.ones_like(origins[..., :1])
            for lst, v in zip(fields, (origins, directions, viewdirs, radii, ones, self.near * ones, self.far * ones)):
                lst.append(v)
        self.rays = Rays(*fields)

    def camera(self, i: int) -> Dict:
        """Pinhole parameters of image i for `obbpose_model.render_camera` / `ops.generate_rays`: the same rays as
        `self.rays[...][i]`, generated on the device (durf_generate_rays, image-centre variant: (x - w/2) / focal)."""
        pp = None if self.principal_point is None else (float(self.principal_point[i, 0]), float(self.principal_point[i, 1]))
        return dict(c2w=self.camtoworlds[i], width=int(self.w[i]), height=int(self.h[i]), focal=float(self.focal[i]),
                    near=float(self.near), far=float(self.far), principal_point=pp)

    def _flatten_time(self, x):
        """:235-253: flatten every image and pool the cameras of one timestep."""
        flat = [y.reshape([-1, y.shape[-1]]) for y in x]
        _, counts = np.unique(self.timesteps, return_counts=True)
        bounds = np.concatenate([[0], np.cumsum(counts)])
        return [np.concatenate(flat[bounds[i]:bounds[i + 1]], axis=0) for i in range(len(counts))]

    # -- batches ------------------------------------------------------------------------------------------
    def _boxes(self, ts_key: int, suffix: str, width: int):
        cars = self.obj_ids[self.obj_ids != 0]
        return np.array([np.asarray(self.box_pose['{}_{}_{}'.format(ts_key, c, suffix)]).reshape(-1) for c in cars]).reshape(-1, width)

    def _next_train(self) -> Dict:
        """:296-326 ('timestep'): one timestep, `batch_size` rays drawn with replacement from its pooled cameras."""
        un = np.unique(self.timesteps)
        t = self._rng.randint(0, len(un), ())
        idx = self._rng.randint(0, self.rays[0][t].shape[0], (self.batch_size,))
        init = np.array([self._boxes(i + 1, 'off' if self.random_box else 'center', 6) for i in range(len(un))]).reshape(len(un), -1, 6)
        return {'pixels': self.images[t][idx], 'rays': Rays(*[r[t][idx] for r in self.rays]), 'depth': self.depth[t][idx],
                'sky': self.sky_mask[t][idx], 'box': self._boxes(t + 1, 'off', 6), 'ext': self._boxes(t + 1, 'ext', 3),
                'can': self._boxes(1, 'off', 6), 'ts': t, 'target': self._boxes(t + 1, 'center', 6), 'init': init}

    def _next_test(self) -> Dict:
        """:330-371: the next held-out frame with the boxes of its timestep."""
        idx = self.it
        self.it = (self.it + 1) % self.n_examples
        if self.render_path:
            raise NotImplementedError("render_path: the spiral path of the reference is generated for spherify = False only")
        t = self.timesteps[idx]
        init = np.array([self._boxes(i + 1, 'center', 6) for i in range(self.total_timesteps)]).reshape(self.total_timesteps, -1, 6)
        return {'pixels': self.images[idx], 'rays': Rays(*[r[idx] for r in self.rays]), 'depth': self.depth[idx],
                'sky': self.sky_mask[idx], 'box': self._boxes(t, 'off', 6), 'init': init, 'ext': self._boxes(t, 'ext', 3),
                'can': self._boxes(1, 'off', 6), 'ts': t - 1, 'target': self._boxes(t, 'center', 6)}


class Waymo(Carla):
    """internal/obbpose_dataset.py:1461-2087, the loader `configs/waymo.gin` selects.  The reference's class is a copy of `Carla`
    with these differences, all of which are the class attributes and hooks below: `poses_bounds.npy` carries the principal
    point (cx, cy) in two extra columns and the rays go through it (:1635-1637, 1882-1885); box extents are FULL extents
    (:1731, / 10); the hold-out is frames 10 and 12 and the 'render' split is every frame (:1806-1813); the object ids are
    1..n from the box dictionary, not from the 2D masks (:1828-1830); sky pixels become 0.975 (:1853)."""

    I_TEST = (10, 12)
    EXT_DIVISOR = 10.0
    SKY_VALUE = 0.975
    RENDER_SPLIT_IS_TRAIN = False
    POSE_COLUMNS = 19

    def _object_ids(self, masks2d):
        last_ts = list(self.box_pose.keys())[-1].split('_')[0]
        return np.arange(1, int(len(self.box_pose) / 3 / int(last_ts)) + 1)


class Carla_Seq(Carla):
    """internal/obbpose_dataset.py:833-1460 ('carla_seq'): a single-camera sequence.  The reference's class is a copy of `Carla`
    with one frame per timestep (:1160-1161), every `llffhold`-th frame as the test split and ALL frames - the held-out ones
    included - as the training split (:1175-1178)."""

    CAMERAS_PER_TIMESTEP = 1

    def _hold_out(self, n: int, config: Config):
        return np.arange(n), np.arange(n)[::config.llffhold]


dataset_dict = {'carla_dyn': Carla, 'carla_seq': Carla_Seq, 'waymo': Waymo}


def get_dataset(split: str, train_dir: str, config: Config):
    """internal/obbpose_dataset.py:17-18."""
    if config.dataset_loader not in dataset_dict:
        raise NotImplementedError(f"dataset_loader {config.dataset_loader!r}: the reference has 'carla_dyn', 'carla_seq' and 'waymo'")
    return dataset_dict[config.dataset_loader](split, train_dir, config)
