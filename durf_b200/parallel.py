"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed; NCCL on the GPU box, gloo in CPU tests).

The path shards by RAYS: every op of the forward pass is per ray, parameters are replicated.  The only data-path
collective is the gradient mean of the train step (`jax.lax.pmean(grad, 'batch')`, train_boxpose.py:253); rendering
needs none (each rank writes its own pixel range; the reference's `all_gather` at train_boxpose.py:378 only hands the
chunks back to one host, which `gather_rows` reproduces when a caller wants the full frame on every rank)."""
from __future__ import annotations

import os
from typing import Sequence, Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process per GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) of `n` rays owned by `rank`; the first n % world ranks get one extra ray."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_rays(rays, rank: int, world: int):
    """Slice every field of a Rays tuple to this rank's contiguous range (utils.shard reshapes to [ndev, -1, ...],
    internal/utils.py:193-196: the same contiguous split when ndev divides the batch)."""
    n = rays[0].shape[0]
    s, e = shard_range(n, rank, world)
    return type(rays)(*[r[s:e] for r in rays])


def allreduce_gradients(d_flat: torch.Tensor) -> float:
    """SUM all-reduce of the flat gradient in place; returns the 1/world factor that turns it into pmean's mean (the
    caller folds it into durf_grad_sanitize, so no extra pass over the gradient is needed)."""
    w = world_size()
    if w > 1:
        dist.all_reduce(d_flat, op=dist.ReduceOp.SUM)
    return 1.0 / w


class GradientBuckets:
    """The gradient mean of the train step (jax.lax.pmean(grad, 'batch'), train_boxpose.py:253) as one SUM all-reduce per
    network, launched on a side stream as soon as that network's slice of the flat gradient is final, so it overlaps the
    rest of the backward pass; `finish()` joins the side stream and returns the 1/world factor (folded into
    durf_grad_sanitize).  The buckets are the contiguous slices of `Variables.flat`: MLP_0 | BoxMLP_k | box_centers."""

    _streams = {}

    def __init__(self, variables, d_flat: torch.Tensor, world: int):
        self.v, self.d_flat, self.world = variables, d_flat, world
        self.works = []
        self.cuda = d_flat.is_cuda
        if self.cuda:
            key = d_flat.device.index
            if key not in GradientBuckets._streams:
                GradientBuckets._streams[key] = torch.cuda.Stream(device=d_flat.device)
            self.side = GradientBuckets._streams[key]

    def network_done(self, name: str) -> None:
        if self.world <= 1:
            return
        bucket = self.v.blob_of(self.d_flat, name)
        if not self.cuda:
            dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
            return
        ev = torch.cuda.Event()
        ev.record()                                   # everything enqueued so far on the compute stream
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            self.works.append(dist.all_reduce(bucket, op=dist.ReduceOp.SUM, async_op=True))

    def finish(self) -> float:
        for w in self.works:
            w.wait()                                  # the compute stream waits for the collective, the host does not
        if self.cuda and self.works:
            torch.cuda.current_stream().wait_stream(self.side)
        self.works = []
        return 1.0 / self.world


def max_over_ranks(value: float, device) -> float:
    """Timing rule: a multi-GPU number is the MAX over ranks of the device-side time."""
    if world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_rows(local: torch.Tensor, counts: Sequence[int]) -> torch.Tensor:
    """All ranks' row blocks concatenated in rank order (ragged last shard allowed)."""
    w = world_size()
    if w == 1:
        return local
    m = max(counts)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(out, pad)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)
