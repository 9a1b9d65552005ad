"""CPU, world_size 2 (gloo): the bucketed gradient all-reduce hook used by the data-parallel path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hyperpri_b200 import parallel
from _mp import from_plain, to_plain


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


class _FakeEngine:
    def __init__(self, rank):
        self.arena = torch.arange(100, dtype=torch.float32) * (rank + 1)
        self.bucket_bounds = [(0, 10), (10, 60), (60, 100)]
        self.bucket_hook = None

    def backward(self):
        for a, b in self.bucket_bounds:
            if self.bucket_hook is not None:          # engines skip the hook when nothing is attached
                self.bucket_hook(self.arena[a:b])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = _FakeEngine(rank)
    red = parallel.attach(eng)
    assert red.world_size == world and abs(red.grad_scale() - 1.0 / world) < 1e-12
    eng.arena.mul_(red.grad_scale())          # the engine folds 1/world into the loss gradient
    eng.backward()
    red.finish()
    q.put(to_plain((rank, eng.arena.clone(), red.bytes)))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [from_plain(q.get(timeout=120)) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = torch.arange(100, dtype=torch.float32) * (1 + 2) / 2       # mean over ranks
    for rank, arena, nbytes in res:
        assert torch.allclose(arena, expect) and nbytes == 400


def test_single_process_is_a_noop():
    eng = _FakeEngine(0)
    red = parallel.attach(eng)
    before = eng.arena.clone()
    eng.backward(); red.finish()
    assert torch.equal(eng.arena, before) and red.grad_scale() == 1.0
    assert eng.bucket_hook is None                    # one process: the engine keeps its single-launch unpack path


def test_engine_hands_every_bucket_to_the_hook_once_in_order():
    """UNetEngine._bucket_done: with an all-reduce hook attached, bucket i is unpacked and handed to the hook as soon as
    its launches are issued (on the side stream, behind its weight gradients -- here, without CUDA, inline), every bucket
    exactly once and in completion order; without a hook only the final bucket triggers the single unpack launch.
    Host logic only: no CUDA."""
    from hyperpri_b200.engine import UNetEngine
    eng = object.__new__(UNetEngine)
    eng.arena = torch.arange(90, dtype=torch.float32)
    eng.bucket_bounds = [(10 * i, 10 * i + 10) for i in range(9)]
    eng._side, eng._side_events, eng._side_used = None, [], 0
    log = []
    eng._unpack_bucket = lambda idx: log.append(("unpack", idx))
    eng.bucket_hook = lambda flat: log.append(("hook", int(flat[0].item()) // 10))
    for idx in range(9):
        log.append(("launched", idx))
        eng._bucket_done(idx)
    assert [e for e in log if e[0] == "hook"] == [("hook", i) for i in range(9)]
    for i in range(9):
        assert log.index(("launched", i)) < log.index(("unpack", i)) < log.index(("hook", i))
        if i < 8:
            assert log.index(("hook", i)) < log.index(("launched", i + 1))
    log.clear()
    eng.bucket_hook = None
    for idx in range(9):
        eng._bucket_done(idx)
    assert log == [("unpack", 8)]


# ---------------------------------------------------------------------------------------------------------------
# PixelParallel (SpectralUNET's model-sharded option): the host-side protocol on CPU tensors -- strips tile the rows,
# all-reducing the per-strip BatchNorm partial sums gives the whole image's statistics, the logits gather restores the
# image.  The kernels around it need a GPU (tests/test_dp_gpu.py::test_pixel_parallel_spectralunet_equals_single_gpu).
def _pp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pp = parallel.PixelParallel()
    R, C, F = 11, 5, 7                                   # odd row count: strips of 5 and 6 rows
    g = torch.Generator().manual_seed(3)
    img = torch.randn((2, 1, R, C), generator=g)          # "logits" of two images, identical on both ranks
    feat = torch.randn((R * C, F), generator=g, dtype=torch.float64)
    r0, r1 = pp.rows(R)
    strips = [pp.rows(R, k) for k in range(world)]
    loc = feat[r0 * C: r1 * C]
    stats = torch.stack([loc.sum(0), (loc * loc).sum(0)], 1)
    pp.all_reduce_(stats)
    full = pp.gather_rows(img[:, :, r0:r1].contiguous(), R)
    q.put(to_plain((rank, strips, stats, full, img, feat, pp.bytes)))
    dist.destroy_process_group()


def test_pixel_parallel_protocol_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [from_plain(q.get(timeout=120)) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, strips, stats, full, img, feat, nbytes in res:
        assert strips == [(0, 5), (5, 11)]                                  # contiguous, disjoint, cover every row
        want = torch.stack([feat.sum(0), (feat * feat).sum(0)], 1)
        assert torch.allclose(stats, want, rtol=1e-12, atol=1e-12)          # whole-image statistics on every rank
        assert torch.equal(full, img)
        assert nbytes == stats.numel() * 8


def test_worker_results_travel_by_value():
    """tests/_mp.py: nested results with tensors of every dtype the workers return survive the plain round trip."""
    import pickle
    obj = (1, {"a": torch.arange(6, dtype=torch.int64).view(2, 3), "b": [torch.tensor(2.5, dtype=torch.float64)]},
           torch.randn(4, 5).to(torch.bfloat16), [0.25, "x"], torch.zeros((), dtype=torch.int64))
    plain = to_plain(obj)
    assert b"torch" not in pickle.dumps(plain).replace(b"torch.", b"")      # no tensor objects left, only dtype names
    back = from_plain(pickle.loads(pickle.dumps(plain)))
    assert back[0] == 1 and back[3] == [0.25, "x"]
    assert torch.equal(back[1]["a"], obj[1]["a"]) and back[1]["a"].dtype == torch.int64
    assert torch.equal(back[1]["b"][0], obj[1]["b"][0]) and back[1]["b"][0].dtype == torch.float64
    assert torch.equal(back[2], obj[2]) and back[2].dtype == torch.bfloat16
    assert back[4].dim() == 0 and back[4].dtype == torch.int64
