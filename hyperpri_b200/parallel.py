"""Data-parallel gradient exchange: the one collective the path has (reference PLTrainer.py:434-442,
Lightning "ddp": NCCL bucketed all-reduce overlapped with backward; plain per-rank BatchNorm).

One process per GPU.  The engine lays every parameter gradient out in one flat fp32 arena ordered by
backward completion; as each bucket (a contiguous arena slice) is finished the hook below starts an
asynchronous NCCL all-reduce on it, which runs on NCCL's stream concurrently with the remaining
backward kernels.  Averaging (DDP divides by world size) is folded into the loss-gradient scale
(`grad_scale = 1 / world_size`), so no extra pass touches the gradients.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class BucketedAllReduce:
    def __init__(self, group=None, engine=None):
        self.group = group
        self.engine = engine
        self.pending: List = []
        self.bytes = 0

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def hook(self, flat_slice: torch.Tensor):
        """Called by the engine when a gradient bucket is complete (stream-ordered)."""
        if self.world_size == 1:
            return
        self.bytes += flat_slice.numel() * flat_slice.element_size()
        self.pending.append(dist.all_reduce(flat_slice, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Make the current stream wait for every outstanding bucket."""
        for w in self.pending:
            w.wait()
        self.pending.clear()
        if self.engine is not None and hasattr(self.engine, "finalize_grads"):
            self.engine.finalize_grads()       # unscale the loss-scaled gradients once every bucket is reduced

    def grad_scale(self) -> float:
        return 1.0 / self.world_size


def attach(engine, group=None) -> BucketedAllReduce:
    r = BucketedAllReduce(group, engine)
    # a single process has nothing to exchange: leave the engine on its one-launch gradient unpack / unscale path
    engine.bucket_hook = r.hook if r.world_size > 1 else None
    return r


class PixelParallel:
    """The model-sharded option for SpectralUNET (`train_net(..., model_parallel=True)`; the reference runs that model as
    DeepSpeed ZeRO-2 + bf16-mixed over >= 2 GPUs, PLTrainer.py:409-433, because one image's activations do not fit
    its GPUs, README.md:82).

    SpectralUNET is a per-pixel MLP (models.py:117-145): the only coupling between pixels is the per-image
    BatchNorm1d statistics in forward, the two BatchNorm reductions in backward and the sum over pixels in the weight
    gradients.  So every rank holds a horizontal strip of EVERY image (rows R*rank/world .. R*(rank+1)/world): the
    activation memory -- the binding constraint -- divides by the world size at any batch size, the weights stay
    replicated.  Exchange per Linear->BatchNorm->ReLU block and image: one all-reduce of 2 x F doubles in forward (sum,
    sum of squares -> exact per-image statistics), one of 3 x F doubles in backward; one all-reduce of the gradient
    arena per step; the logits strips (R x C floats per image) are all-gathered for the loss.  Results equal the
    single-GPU run up to summation order."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bytes = 0

    def rows(self, R: int, rank=None):
        """[r0, r1) of this rank's strip of an R-row image."""
        k = self.rank if rank is None else rank
        return (R * k) // self.world, (R * (k + 1)) // self.world

    def all_reduce_(self, t: torch.Tensor) -> torch.Tensor:
        """In-place SUM over the group, ordered on the current stream."""
        if self.world > 1:
            self.bytes += t.numel() * t.element_size()
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def gather_rows(self, local: torch.Tensor, R: int, out: torch.Tensor = None) -> torch.Tensor:
        """local: [n, 1, r1 - r0, C] strip of every image -> [n, 1, R, C] on every rank."""
        n, one, _, C = local.shape
        if out is None:
            out = torch.empty((n, one, R, C), dtype=local.dtype, device=local.device)
        if self.world == 1:
            out.copy_(local)
            return out
        rmax = max(self.rows(R, k)[1] - self.rows(R, k)[0] for k in range(self.world))
        pad = torch.zeros((n, one, rmax, C), dtype=local.dtype, device=local.device)
        pad[:, :, : local.shape[2]] = local
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(parts, pad, group=self.group)
        for k, p in enumerate(parts):
            r0, r1 = self.rows(R, k)
            out[:, :, r0:r1] = p[:, :, : r1 - r0]
        return out
