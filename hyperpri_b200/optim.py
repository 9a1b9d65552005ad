"""Multi-tensor Adam on the native library: one launch per step for every parameter (reference
src/PLTrainer.py:171-174 builds ``optim.Adam(params, lr, weight_decay)``).

Drop-in for ``torch.optim.Adam`` with its default flags (no amsgrad / maximize / capturable): same update rule, same
``state_dict`` layout (``step``, ``exp_avg``, ``exp_avg_sq`` per parameter), so checkpoints move both ways.
Parameters must be fp32 CUDA tensors with fp32 ``.grad``; anything else raises (no fallback to a torch loop).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, found_inf=None):
        """found_inf: optional int32 CUDA tensor (the engine's overflow flag, hyperpri_b200.engine._EngineBase.overflow).
        While it is non-zero the launch leaves parameters and moments untouched (the step is skipped on the device,
        without a host synchronisation); `skipped_steps()` reads how often that happened."""
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}
        self.found_inf = found_inf
        self._skipped = None

    def _table(self, gi, plist):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr()) for p in plist)
        t = self._tables.get(gi)
        if t is None or t[0] != key:
            arr = (_lib.AdamJob * len(plist))()
            blocks = 0
            for i, p in enumerate(plist):
                st = self.state[p]
                arr[i].param, arr[i].grad = p.data_ptr(), p.grad.data_ptr()
                arr[i].exp_avg, arr[i].exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                arr[i].numel, arr[i].block0 = p.numel(), blocks
                blocks += (p.numel() + 1023) // 1024
            dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(plist[0].device)
            t = (key, dev, len(plist), blocks)
            self._tables[gi] = t
        return t

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                if not (p.is_cuda and p.dtype == torch.float32 and p.grad.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters and gradients")
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            step = int(self.state[plist[0]]["step"].item()) + 1
            _, dev, n, blocks = self._table(gi, plist)
            b1, b2 = group["betas"]
            check(_lib.lib().hpri_adam_step(C.c_void_p(dev.data_ptr()), n, blocks, float(group["lr"]), float(b1),
                                            float(b2), float(group["eps"]), float(group["weight_decay"]), step,
                                            C.c_void_p(0 if self.found_inf is None else self.found_inf.data_ptr()),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "hpri_adam_step")
            if self.found_inf is not None:            # device-side tally of skipped steps (no synchronisation)
                if self._skipped is None:
                    self._skipped = torch.zeros(1, dtype=torch.int64, device=self.found_inf.device)
                self._skipped += (self.found_inf != 0)
            for p in plist:
                self.state[p]["step"] += 1
                torch.autograd.graph.increment_version(p)     # the engine re-packs operands of changed parameters
        return loss

    def skipped_steps(self) -> int:
        """Steps the device skipped because the loss-scaled fp16 gradients overflowed (synchronises)."""
        return 0 if self._skipped is None else int(self._skipped.item())
