"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): ray sharding, the gradient mean that replaces
jax.lax.pmean (train_boxpose.py:253), max-over-ranks timing and frame assembly."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from durf_b200 import parallel
from durf_b200.utils import Rays


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert parallel.env_rank_world() == (rank, world, rank)
        n = 1001                                            # ragged: not divisible by the world size
        g = torch.Generator().manual_seed(5)
        rays = Rays(*[torch.rand(n, d, generator=g) for d in (3, 3, 3, 1, 1, 1, 1)])
        mine = parallel.shard_rays(rays, rank, world)
        s, e = parallel.shard_range(n, rank, world)
        assert mine.origins.shape[0] == e - s and torch.equal(mine.radii, rays.radii[s:e])
        # per-rank "gradient": a deterministic function of the rank's rays, like a loss summed over the local batch
        grad = torch.stack([mine.origins.sum(), mine.directions.sum(), (mine.radii ** 2).sum()]).repeat(7)
        local = grad.clone()
        scale = parallel.allreduce_gradients(grad)
        mean = grad * scale
        # frame assembly: every rank renders its rows, rank order restores the frame
        counts = [parallel.shard_range(n, r, world)[1] - parallel.shard_range(n, r, world)[0] for r in range(world)]
        frame = parallel.gather_rows(mine.origins * 2.0, counts)
        assert torch.allclose(frame, rays.origins * 2.0)
        t = parallel.max_over_ranks(10.0 + rank, torch.device("cpu"))
        assert t == 10.0 + world - 1
        torch.save(dict(local=local, mean=mean), os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_mean_and_sharding(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    want = sum(o['local'] for o in outs) / world               # pmean: mean over devices of the per-device gradient
    for o in outs:
        assert torch.allclose(o['mean'], want, rtol=1e-6)
    assert torch.equal(outs[0]['mean'], outs[1]['mean'])      # every rank steps with the identical gradient


def test_shard_range_covers_every_ray_once():
    for n in (0, 1, 7, 1000, 2457600):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
