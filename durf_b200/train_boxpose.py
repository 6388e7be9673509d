"""Host driver mirroring the reference's `train_boxpose.py:main` (324-581) on top of the CUDA hot path:

    python -m durf_b200.train_boxpose --gin_file=configs/carla_dyn.gin --train_dir=/tmp/run [--max_steps=N]

It reads the reference's .gin files unchanged (`utils.load_gin`), builds the model (`construct_mipnerf`), restores the
newest checkpoint (`init_step = state.step + 1`, :404-406), runs the schedules of :347-367 (log-lerp learning rate with
sine warm-up, eps, BARF alpha), the train step, the rays/sec logging of :518-528, periodic checkpoints (:529-532) and a
final test render.  With `--data_dir` batches come from the CARLA loader (`durf_b200.obbpose_dataset.Carla`, the mirror of
internal/obbpose_dataset.py's `Carla`); no dataset ships with the reference, so by default they come from
`SyntheticTimestepDataset`, which reproduces the loader's batch contract ('timestep' batching: all rays of a batch share one
`ts`; fields rays/pixels/depth/sky/ext/init/ts) on synthetic pinhole rays."""
from __future__ import annotations

import argparse
import os
import time
from typing import Dict, Iterator

import numpy as np
import torch

from . import checkpoint, math as dmath, obbpose_dataset, parallel, synthetic as S
from .obbpose_model import MipNerfModel, Variables, render_camera
from .train import GraphedTrainStep, TrainState, train_step
from .utils import Config, Rays, load_gin


class SyntheticTimestepDataset:
    """Batches shaped like `Carla._next_train` with batching='timestep' (obbpose_dataset.py:293-328), synthetic content."""

    def __init__(self, config: Config, num_objects: int, device, seed: int = S.SEED, rank: int = 0):
        self.cfg, self.K, self.dev = config, num_objects, device
        # The SCENE (cameras, boxes, hence the initial box_centers parameter) is the same on every rank: the reference
        # replicates one state to all devices (flax.jax_utils.replicate, train_boxpose.py:407) and shards only the batch.
        scene_rng = np.random.default_rng(seed)
        self.c2w = [S.random_c2w(scene_rng) for _ in range(config.timesteps)]
        self.centers, self.ext = S.boxes_in_view(scene_rng, self.c2w[0], num_objects, timesteps=config.timesteps)
        # the per-rank stream only draws pixels / targets; the timestep sequence is shared (one `ts` per global batch)
        self.ts_rng = np.random.default_rng(seed + 7919)
        self.rng = np.random.default_rng(seed + 104729 * (rank + 1))

    def peek(self) -> Dict:
        return dict(init=self.centers, ext=self.ext)

    def __iter__(self) -> Iterator[Dict]:
        B = self.cfg.batch_size
        while True:
            ts = int(self.ts_rng.integers(0, self.cfg.timesteps))
            rays, _ = S.random_rays(self.rng, B, c2w=self.c2w[ts], near=self.cfg.near, far=self.cfg.far)
            tg = S.targets(self.rng, B)
            to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev, non_blocking=True)
            yield dict(rays=Rays(*[to(a) for a in rays]), pixels=to(tg['pixels']), depth=to(tg['depth']), sky=to(tg['sky']),
                       ext=to(self.ext), ts=ts, init=self.centers)


class DiskDataset:
    """The CARLA loader's numpy batches (durf_b200.obbpose_dataset.Carla) as device batches: pinned staging, non-blocking
    copies on the current stream (what utils.shard / device_put does in the reference, train_boxpose.py:424)."""

    def __init__(self, loader, device):
        self.loader, self.device = loader, device

    def _to_dev(self, b):
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).pin_memory().to(self.device, non_blocking=True)
        return dict(rays=Rays(*[up(r) for r in b['rays']]), pixels=up(b['pixels']), depth=up(b['depth']), sky=up(b['sky']),
                    ext=up(b['ext']), init=np.asarray(b['init'], np.float32), ts=int(b['ts']))

    def peek(self):
        return self._to_dev(self.loader.peek())

    def __iter__(self):
        return self

    def __next__(self):
        return self._to_dev(next(self.loader))


def main(argv=None) -> Dict:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--gin_file", required=True)
    ap.add_argument("--train_dir", required=True)
    ap.add_argument("--max_steps", type=int, default=None)
    ap.add_argument("--batch_size", type=int, default=None)
    ap.add_argument("--save_every", type=int, default=50000)
    ap.add_argument("--print_every", type=int, default=100)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--render_rows", type=int, default=0, help="rows of the final 1920-wide test render (0: none)")
    ap.add_argument("--no_graph", action="store_true", help="launch every step from Python instead of replaying it from a CUDA graph")
    ap.add_argument("--data_dir", default=None, help="an on-disk CARLA scene (internal/obbpose_dataset.py layout); default: synthetic batches")
    args = ap.parse_args(argv)

    rank, world, local = parallel.env_rank_world()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=dev)
    cfg_kw, model_kw = load_gin(args.gin_file)
    if args.max_steps is not None:
        cfg_kw["max_steps"] = args.max_steps
    if args.batch_size is not None:
        cfg_kw["batch_size"] = args.batch_size
    config = Config(**cfg_kw)
    model = MipNerfModel(precision=args.precision, timesteps=config.timesteps,
                         **{k: v for k, v in model_kw.items() if k in MipNerfModel.__dataclass_fields__})
    if args.data_dir:
        dataset = DiskDataset(obbpose_dataset.get_dataset('train', args.data_dir, config), dev)      # train_boxpose.py:331
    else:
        dataset = SyntheticTimestepDataset(config, model.num_objects, dev, seed=S.SEED, rank=rank)
    variables = model.init(np.random.default_rng(20200823), dataset.peek()["init"], device=dev)       # train_boxpose.py:325
    state = checkpoint.restore_checkpoint(args.train_dir, TrainState.create(variables))
    if world > 1:
        # replicas must start from identical parameters and moments whatever was restored on each rank
        for t in (variables.flat, state.m, state.v):
            torch.distributed.broadcast(t, 0)
    init_step = state.step + 1 if state.step > 0 else 1
    lr_fn = lambda s: dmath.learning_rate_decay(s, config.lr_init, config.lr_final, config.max_steps, config.lr_delay_steps,
                                                config.lr_delay_mult)
    eps_fn = lambda s: dmath.learning_rate_decay(s, config.eps_init, config.eps_final, config.eps_max_steps, config.eps_delay_steps,
                                                 config.lr_delay_mult)
    alpha_fn = lambda s: dmath.freq_alpha_rate(s, config.alpha_init, config.alpha_final, config.alpha_delay_steps,
                                               config.alpha_max_steps)
    prevs = torch.from_numpy(np.asarray(dataset.peek()["init"], np.float32)).to(dev)
    t0, losses, last = time.time(), [], {}
    # The step is captured once in a CUDA graph and replayed (at the reference's shipped batch of 512 rays launching its ~40
    # kernels from Python takes longer than running them); the fp32 parity mode sizes its GEMMs on the host and is not capturable.
    graphed, use_graph = None, not args.no_graph and model.step_is_capturable()
    for step, batch in zip(range(init_step, config.max_steps + 1), dataset):
        ts = batch["ts"]
        prev = prevs[ts + 1 if ts == 0 else ts - 1][None]                        # train_boxpose.py:453-456
        if use_graph and graphed is None:
            graphed = GraphedTrainStep(model, config, state, batch['rays'][0].shape[0], variables.K, world_size=world, device=dev,
                                       use_prev=True)
        if graphed is not None:
            stats = graphed(batch, lr_fn(step), eps_fn(step), alpha_fn(step), prev=prev)
        else:
            state, stats = train_step(model, config, None, state, batch, lr_fn(step), eps_fn(step), alpha_fn(step), prev=prev,
                                      world_size=world)
        prevs[ts, :, :3] = stats['pose']              # the FORWARD pass's pose (stats.pose), not the post-Adam one (:461)
        state.step = step
        if step % args.print_every == 0 or step == config.max_steps:
            torch.cuda.synchronize()
            dt = time.time() - t0
            n = min(args.print_every, step - init_step + 1)
            last = dict(step=step, loss=float(stats["loss"]), lr=lr_fn(step), rays_per_sec=config.batch_size * world * n / max(dt, 1e-9))
            losses.append(last["loss"])
            if rank == 0:
                print(f"{step:>7d}/{config.max_steps}: loss={last['loss']:.4f} lr={last['lr']:.2e} "
                      f"{last['rays_per_sec']:.0f} r/s")                                                 # train_boxpose.py:518-528
            t0 = time.time()
        if rank == 0 and (step % args.save_every == 0 or step == config.max_steps):
            checkpoint.save_checkpoint(args.train_dir, state, step, keep=100)
    out = dict(last=last, losses=losses, step=state.step)
    if args.render_rows > 0 and rank == 0:
        fn = lambda rng, b: model.apply(variables, rng, b["rays"], None, b["ext"], b["ts"], False, False, False, b["alpha"])
        if args.data_dir:      # the first held-out frame of the scene, rays generated on the device
            test = obbpose_dataset.get_dataset('test', args.data_dir, config)
            tb, cam = test.peek(), test.camera(0)
            rgb, dist, acc = render_camera(fn, cam['c2w'], cam['width'], min(args.render_rows, cam['height']), cam['focal'], cam['near'],
                                           cam['far'], None, torch.from_numpy(np.asarray(tb['ext'], np.float32)).to(dev), int(tb['ts']),
                                           None, alpha_fn(state.step), principal_point=cam['principal_point'])
        else:
            rgb, dist, acc = render_camera(fn, dataset.c2w[0], S.WAYMO_W, args.render_rows, S.FOCAL, config.near, config.far, None,
                                           torch.from_numpy(dataset.ext).to(dev), 0, None, alpha_fn(state.step))
        out["render_mean_rgb"] = float(rgb.mean())
    return out


if __name__ == "__main__":
    main()
