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

from _mp import from_plain, to_plain

pytestmark = pytest.mark.gpu


def _collect(q, procs, timeout=240):
    """One result per process; fails as soon as a worker has died instead of waiting for the full timeout."""
    import queue
    import time
    out, t0 = [], time.time()
    while len(out) < len(procs):
        try:
            out.append(from_plain(q.get(timeout=2)))
        except queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            if dead or time.time() - t0 > timeout:
                for p in procs:
                    if p.is_alive():
                        p.kill()
                raise AssertionError(f"worker failed (exit codes {dead}) or timed out after {time.time() - t0:.0f} s")
    return out


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
    q.put(to_plain((rank, {k: v.cpu() for k, v in out.items()}, int(eng.overflow.item()), exact)))
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
    for rank, g, ovf, exact in _collect(q, procs):
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


# ---------------------------------------------------------------------------------------------------------------
# SpectralUNET, model-sharded option (train_net(..., model_parallel=True); reference: DeepSpeed ZeRO-2 over >= 2 GPUs,
# PLTrainer.py:409-433): every rank holds a row strip of every image, per-image BatchNorm statistics all-reduced per
# block.  Two GPUs must give the single-GPU logits, loss and gradients (up to summation order / fp16 rounding noise).
def _pp_net(dev, feats):
    from hyperpri_b200.src.Experiments.models import SpectralUNET
    torch.manual_seed(11)
    return SpectralUNET(24, 1, bn_feats=feats).to(dev).train()


def _pp_batch(dev):
    g = torch.Generator().manual_seed(5)
    x = torch.rand((2, 24, 37, 64), generator=g).to(dev)           # odd row count: strips of 18 and 19 rows
    m = (torch.rand((2, 1, 37, 64), generator=g) > 0.7).float().to(dev)
    return x, m


def _pp_step(net, x, m):
    net.zero_grad(set_to_none=True)
    loss, logits, counts = net.bce_step(x, m)
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().float().cpu().clone() for k, p in net.named_parameters()}
    bufs = {k: v.detach().cpu().clone() for k, v in net.named_buffers()}
    return loss.item(), logits.cpu(), counts.cpu(), grads, bufs


def _pp_worker(rank, world, port, feats, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    net = _pp_net(dev, feats)
    pp = net.enable_pixel_parallel(None)
    assert pp.world == world and net._get_engine(dev).pp is pp
    x, m = _pp_batch(dev)
    out = None
    for _ in range(2):                       # twice: workspace / event reuse; the second step sees updated running stats
        out = _pp_step(net, x, m)
    q.put(to_plain((rank,) + out + (int(net._get_engine(dev).overflow.item()), pp.bytes)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("feats", [96, 1650])
def test_pixel_parallel_spectralunet_equals_single_gpu(feats):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pp_worker, args=(r, 2, port, feats, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for r in _collect(q, procs):
        res[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    dev = torch.device("cuda", 0)
    net = _pp_net(dev, feats)
    x, m = _pp_batch(dev)
    for _ in range(2):
        loss1, logits1, counts1, grads1, bufs1 = _pp_step(net, x, m)

    def cos(a, b):
        return float(torch.dot(a.flatten().double(), b.flatten().double()) / (a.double().norm() * b.double().norm() + 1e-300))
    for rank in (0, 1):
        loss, logits, counts, grads, bufs, ovf, nbytes = res[rank]
        assert ovf == 0 and nbytes > 0
        assert logits.shape == logits1.shape == (2, 1, 37, 64)
        assert (logits - logits1).abs().max().item() <= 5e-3 * logits1.abs().max().item()
        assert abs(loss - loss1) < 1e-5 and (counts - counts1).abs().max().item() <= 4
        for k in bufs1:
            if "running_" in k:
                assert torch.allclose(bufs[k], bufs1[k], rtol=2e-3, atol=2e-4), k
            elif "num_batches" in k:
                assert int(bufs[k]) == int(bufs1[k]) == 4                      # two images x two steps
        fa = torch.cat([grads[k].flatten() for k in sorted(grads1)])
        fb = torch.cat([grads1[k].flatten() for k in sorted(grads1)])
        assert cos(fa, fb) > 0.995 and abs(fa.norm().item() / fb.norm().item() - 1.0) < 0.02
    for k in res[0][3]:
        assert torch.equal(res[0][3][k], res[1][3][k]), k                      # both ranks hold identical reduced gradients


# ---------------------------------------------------------------------------------------------------------------
# The reference's entry function under torchrun-style launch on two GPUs (PLTrainer.py:333-460): train_net with the
# engine-backed models -- data parallel (CubeNET: sharded sampler, bucketed all-reduce on the side stream, FusedAdam) and
# model_parallel=True (SpectralUNET: pixel-parallel strips).  Replicas must stay bit-identical, the loss must fall,
# checkpoints are written by rank 0 only.
def _write_tiny_dataset(root, n_train=4, n_val=2, h=32, w=48):
    import json as _json
    import numpy as np
    from PIL import Image
    from hyperpri_b200 import envi
    base = os.path.join(root, "Datasets", "HyperPRI", "Peanut_968x608")
    for d in ("rgb_files", "hsi_files", "mask_files"):
        os.makedirs(os.path.join(base, d), exist_ok=True)
    os.makedirs(os.path.join(root, "Datasets", "HyperPRI", "data_splits"), exist_ok=True)
    rs = np.random.RandomState(0)
    for split, cnt, day0 in (("train", n_train, 1), ("val", n_val, 11)):
        dates = [f"202207{day0 + i:02d}" for i in range(cnt)]
        for date in dates:
            stem = f"{date}_box37_ref"
            mask = np.zeros((h, w), np.uint8)
            mask[:, w // 3: w // 3 + 6] = 3
            cube = rs.random_sample((h, w, 299)).astype(np.float32) * 0.2
            cube[mask > 0, 100:180] += 0.6
            envi.save(os.path.join(base, "hsi_files", "hinalea_hsi.hdr"), os.path.join(base, "hsi_files", stem + ".dat"), cube)
            Image.fromarray(mask).save(os.path.join(base, "mask_files", stem + "_mask.png"))
            Image.fromarray((rs.random_sample((h, w, 3)) * 255).astype(np.uint8)).save(os.path.join(base, "rgb_files", stem + ".png"))
        js = {"img_dir": "rgb_files", "hsi_dir": "hsi_files", "mask_dir": "mask_files", "notes": "synthetic",
              "box37": {"plant_folder": "Peanut", "resolution": "968x608", "box_no": 37, "phenotype": 1, "dates": dates,
                        "weights": None}}
        with open(os.path.join(root, "Datasets", "HyperPRI", "data_splits", f"{split}1.json"), "w") as f:
            _json.dump(js, f)


def _train_worker(rank, world, port, root, model_name, model_parallel, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from hyperpri_b200.src import PLTrainer as T
    from hyperpri_b200.src.Experiments.params_HyperPRI import ExpHyperspectralPRI
    torch.manual_seed(1000 + rank)                       # replicas start DIFFERENT: train_net must synchronise them
    p = ExpHyperspectralPRI(root, split_no=1, seed_num=0, comet_logging=False)
    p.epochs, p.patch_size, p.spectral_bn_size = 3, (32, 48), 64
    p.change_network_param(model_name, root, 1)
    trainer = T.train_net(p, model_parallel=model_parallel)
    net = trainer.model.m_network
    flat = torch.cat([t.detach().float().flatten() for t in net.parameters()]).cpu()
    eng = net._get_engine(torch.device("cuda", rank))
    q.put(to_plain((rank, flat, [h["tr_loss"] for h in trainer.history], [h["val_loss"] for h in trainer.history],
                    int(eng.overflow.item()), sorted(os.listdir(os.path.join(p.save_path, "Checkpoints"))))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("model_name,model_parallel", [("CubeNET", False), ("SpectralUNET", True)])
def test_train_net_on_two_gpus(tmp_path, model_name, model_parallel):
    root = str(tmp_path)
    _write_tiny_dataset(root)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, root, model_name, model_parallel, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r[0]: r[1:] for r in _collect(q, procs, timeout=400)}
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert torch.equal(res[0][0], res[1][0])                       # replicas bit-identical after three epochs
    assert res[0][1][-1] < res[0][1][0]                           # training loss falls
    assert res[0][2] == res[1][2]                                  # the monitored loss is the same number on both ranks
    assert res[0][3] == 0 and res[1][3] == 0                       # no fp16 gradient overflow
    assert res[0][4] == ["best.ckpt", "last.ckpt"]
    if model_parallel:                                             # both ranks saw the whole batch: identical training losses
        assert res[0][1] == res[1][1]
