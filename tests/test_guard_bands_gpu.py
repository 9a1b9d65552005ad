"""Out-of-bounds and uninitialised-memory check of a whole training step, without a sanitizer.

Every device tensor the engines allocate (hyperpri_b200/engine.py `_z` / `_e`: activations, gradients, concat buffers,
logits, loss scalars; and, through a stand-in for the `torch` factory functions engine.py and ops.py call, the packed
16-bit weights, gradient packs, the flat gradient arena and the statistics slots) is replaced by the interior of a larger byte buffer with a 64 KiB guard band on either
side.  The guard bands, and the interior of every buffer the engine asks for UNINITIALISED, are filled with 0xFF bytes:
NaN as fp16 / bf16 / fp32 / fp64.  One forward + BCE + backward then has to

  * leave every guard byte untouched (no kernel stores outside the tensor it was given),
  * produce finite logits and gradients (nothing read from a guard band or from a never-written element reaches a
    result -- a NaN survives every multiply, including the one by a zero weight),
  * and, in the deterministic-statistics mode, give bit-identical logits to the same step on ordinary allocations
    (what the kernels compute does not depend on what the allocator left in memory).

Odd image sizes on purpose (partial tiles, floor pooling, Up's one-pixel padding), every model family and constructor
flag, the heuristic kernel choice and the halo-reuse kernel on CTA pairs forced.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import hyperpri_oracle as O                                                    # noqa: E402
from hyperpri_b200 import engine as E, ops                                     # noqa: E402
from test_models_gpu import build                                              # noqa: E402

BAND = 1 << 16


class Guarded:
    def __init__(self):
        self.bufs = []

    def alloc(self, shape, dev, dtype, zero):
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        nbytes = math.prod(shape) * torch.empty((), dtype=dtype).element_size()
        tail = BAND + (-nbytes % 256)
        raw = torch.full((BAND + nbytes + tail,), 0xFF, dtype=torch.uint8, device=dev)
        inner = raw[BAND:BAND + nbytes]
        if zero:
            inner.zero_()
        self.bufs.append((raw, nbytes, shape, dtype))
        return inner.view(dtype).view(shape)

    def z(self, shape, dev, dtype=E.ACT):
        return self.alloc(shape, dev, dtype, True)

    def e(self, shape, dev, dtype=E.ACT):
        return self.alloc(shape, dev, dtype, False)

    def check(self):
        assert self.bufs
        for raw, nbytes, shape, dtype in self.bufs:
            lo, hi = raw[:BAND], raw[BAND + nbytes:]
            assert bool((lo == 0xFF).all()), ("store below the tensor", shape, dtype,
                                              int((lo != 0xFF).nonzero()[-1]) - BAND)
            assert bool((hi == 0xFF).all()), ("store past the tensor", shape, dtype, int((hi != 0xFF).nonzero()[0]))


class TorchShim:
    """`torch` as engine.py / ops.py see it, with the factory functions they allocate device buffers with (packed
    weights, gradient packs, the gradient arena, statistics slots, the second ingest buffer) routed to Guarded."""

    def __init__(self, g):
        self._g = g

    def __getattr__(self, k):
        return getattr(torch, k)

    def _new(self, size, dtype, device, zero):
        if device is None or torch.device(device).type != "cuda":
            return (torch.zeros if zero else torch.empty)(*size, dtype=dtype, device=device)
        return self._g.alloc(size[0] if len(size) == 1 else size, device, dtype or torch.float32, zero)

    def zeros(self, *size, dtype=None, device=None):
        return self._new(size, dtype, device, True)

    def empty(self, *size, dtype=None, device=None):
        return self._new(size, dtype, device, False)

    def ones(self, *size, dtype=None, device=None):
        return self._new(size, dtype, device, True).fill_(1)

    def empty_like(self, t):
        return self._g.alloc(tuple(t.shape), t.device, t.dtype, False)


def step(net, xin, mask, fused):
    net.train()
    net.zero_grad(set_to_none=True)
    if fused:
        loss, logits, _ = net.bce_step(xin.cuda(), mask.cuda(), 0.5)
    else:
        logits = net(xin.cuda())
        loss = torch.nn.BCEWithLogitsLoss()(logits, mask.cuda())
    loss.backward()
    torch.cuda.synchronize()
    return logits.detach().clone(), loss.item(), {k: p.grad.detach().clone() for k, p in net.named_parameters()}


CASES = [
    # model, n, bands, h, w, feats, ctor flags, fused bce_step
    ("CubeNET", 1, 238, 37, 51, 0, {}, True),
    ("CubeNET", 2, 238, 33, 47, 0, dict(att=True, fd=32), False),
    ("UNET", 2, 3, 35, 41, 0, dict(bil=True), False),
    ("UNET", 1, 3, 50, 34, 0, dict(bil=True, att=True), True),
    ("SpectralUNET", 2, 238, 5, 7, 96, {}, True),
    ("SpectralUNET", 1, 238, 9, 13, 40, dict(bnorm=False), False),
]


@pytest.mark.parametrize("algo", [-1, 2], ids=["default", "halo_pair"])
@pytest.mark.parametrize("model,n,bands,h,w,feats,flags,fused", CASES)
def test_training_step_stays_inside_its_buffers_and_ignores_stale_memory(model, n, bands, h, w, feats, flags, fused,
                                                                         algo, monkeypatch):
    if model == "SpectralUNET" and algo != -1:
        pytest.skip("no 3x3 convolutions")
    x = O.synth_cube(3, n, bands, h, w)
    xin = x[:, None] if model == "CubeNET" else x
    mask = O.synth_mask(3, n, h, w)
    ops.set_deterministic(True, backward=False)
    ops.set_conv_algo(algo)
    try:
        net, _ = build(model, bands, feats or 1650, seed=2, **flags)
        lg0, loss0, g0 = step(net, xin, mask, fused)
        _, _, g0b = step(net, xin, mask, fused)                # run-to-run spread of the gradients on ordinary memory
        g = Guarded()
        monkeypatch.setattr(E, "_z", g.z)
        monkeypatch.setattr(E, "_e", g.e)
        monkeypatch.setattr(E, "torch", TorchShim(g))
        monkeypatch.setattr(ops, "torch", TorchShim(g))
        net, _ = build(model, bands, feats or 1650, seed=2, **flags)
        lg1, loss1, g1 = step(net, xin, mask, fused)
        g.check()
        lg2, loss2, g2 = step(net, xin, mask, fused)          # second step on the same (now dirty) workspaces
        g.check()
    finally:
        ops.set_conv_algo(-1)
        ops.set_deterministic(False)
    assert torch.isfinite(lg1).all() and math.isfinite(loss1)
    for k, v in g1.items():
        assert torch.isfinite(v).all(), k
        assert torch.isfinite(g2[k]).all(), k
    assert torch.equal(lg1, lg0) and loss1 == loss0
    assert torch.equal(lg2, lg1)
    flat0, flat0b, flat1 = (torch.cat([v.flatten() for v in d.values()]) for d in (g0, g0b, g1))
    # the backward pass keeps its atomics (split-K weight gradients, BatchNorm-backward sums) and stores 16-bit
    # gradients: equal up to what two runs on ordinary memory differ by
    noise = ((flat0b - flat0).norm() / flat0.norm()).item()
    diff = ((flat1 - flat0).norm() / flat0.norm()).item()
    print(f"gradient rel. L2 difference: guarded vs plain {diff:.3e}, plain vs plain {noise:.3e}")
    assert diff <= max(1e-3, 4 * noise)
