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
