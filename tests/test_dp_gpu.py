"""Two GPUs, NCCL: the data-parallel gradient exchange on the real engine (reference PLTrainer.py:434-442, Lightning
"ddp": per-rank BatchNorm, gradients averaged over ranks).  Each rank runs CubeNET on its own shard with the bucketed
all-reduce hook attached (weight gradients on the side stream, buckets flushed one late).  Checked: (1) exactly -- the
final arena is the sum over ranks of the local (already unscaled) bucket contents at hook time, every arena element belongs to
one bucket, both ranks end bit-identical; (2) against the mean of the two shards' single-process gradients by cosine
similarity (tiny random-init nets amplify fp16 rounding differences between runs, see test_models_gpu.py).
Skipped on a one-GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _net(dev):
    from hyperpri_b200.src.Experiments.models import CubeNET
    torch.manual_seed(7)
    return CubeNET(24, 1, first_depth=64, bilinear=False).to(dev).train()


def _shard(rank, dev):
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.rand((2, 1, 24, 128, 160), generator=g).to(dev)
    m = (torch.rand((2, 1, 128, 160), generator=g) > 0.7).float().to(dev)
    return x, m


def _grads(eng, x, m, red=None):
    scale = red.grad_scale() if red is not None else 1.0
    eng.invalidate_packed()
    logits = eng.forward(x, True)
    _, dlogit, _ = eng.loss_and_dlogit(logits, m, grad_scale=scale)
    g = eng.backward(dlogit, prescaled=True)
    if red is not None:
        red.finish()
    torch.cuda.synchronize()
    return {k: v.detach().float().clone() for k, v in g.items()}


def _worker(rank, world, port, q):
    from hyperpri_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    net = _net(dev)
    eng = net._get_engine(dev)
    red = parallel.attach(eng)
    assert eng.bucket_hook is not None
    real_hook, local, seen = eng.bucket_hook, [], []

    def spy(flat_slice):                  # the bucket as this rank computed it, before the exchange overwrites it
        local.append((flat_slice.data_ptr(), flat_slice.detach().clone()))
        seen.append(flat_slice.numel())
        real_hook(flat_slice)
    eng.bucket_hook = spy
    x, m = _shard(rank, dev)
    out = None
    for _ in range(2):                    # twice: event / buffer reuse across steps
        local.clear(); seen.clear()
        out = _grads(eng, x, m, red)
    assert sum(seen) == eng.arena.numel(), (sum(seen), eng.arena.numel())
    exact = True
    base = eng.arena.data_ptr()
    for ptr, mine in local:
        both = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        off = (ptr - base) // 4
        want = both[0] + both[1]           # buckets leave the unpack kernels already unscaled (1 / loss scale folded in)
        exact = exact and torch.equal(eng.arena[off:off + mine.numel()], want)
    q.put((rank, {k: v.cpu() for k, v in out.items()}, int(eng.overflow.item()), exact))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_dp_gradients_equal_mean_of_shards():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        rank, g, ovf, exact = q.get(timeout=600)
        assert ovf == 0 and exact
        res[rank] = g
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single-process gradients of each shard (same seed -> same weights), no hook
    dev = torch.device("cuda", 0)
    net = _net(dev)
    eng = net._get_engine(dev)
    assert eng.bucket_hook is None
    per = [_grads(eng, *_shard(r, dev)) for r in range(2)]
    def cos(a, b):
        return float(torch.dot(a.flatten().double(), b.flatten().double()) /
                     (a.double().norm() * b.double().norm() + 1e-300))
    bad, fa, fb = [], [], []
    for k in per[0]:
        want = (0.5 * (per[0][k] + per[1][k])).cpu()
        assert torch.equal(res[0][k], res[1][k]), k      # both ranks hold the same reduced values
        if "bias" in k and "double_conv.0" in k or "double_conv.3.bias" in k or k.endswith("inc.0.bias") or k.endswith("inc2.0.bias") or k == "first_conv.bias":
            continue                                     # conv biases under train-mode BN: gradient identically zero
        fa.append(res[0][k].flatten()); fb.append(want.flatten())
        if want.numel() >= 64 and cos(res[0][k], want) < 0.8:
            bad.append((k, cos(res[0][k], want)))
    assert not bad, bad
    assert cos(torch.cat(fa), torch.cat(fb)) > 0.97
