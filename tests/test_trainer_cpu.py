"""CPU: the trainer layer's host logic (reference src/PLTrainer.py:270-460) -- checkpoint discovery and loading of
Lightning-layout files, the built-in loop's monitoring / early stopping / resume, and its "ddp" semantics under a
world-size-2 gloo group (sharded sampler, replicas synchronised from rank 0, rank 0 alone writes checkpoints).
The networks here are small torch modules: the engine-backed models need a GPU and are covered by the -m gpu tests."""
import os
import socket
import sys
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch.utils.data import DataLoader, Dataset

from hyperpri_b200.src import PLTrainer as T
from _mp import from_plain, to_plain
from hyperpri_b200.src.Experiments.params_HyperPRI import ExpHyperspectralPRI, ExpRedGreenBluePRI


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


class _Params:
    """The attributes _Loop and RootLightningModel read."""
    def __init__(self, root, net, epochs=50, overall=3):
        self.save_path = os.path.join(root, "Saved_Models", "HSI", "Tiny", "Run_1") + "/"
        self.epochs, self.overall, self.run_num = epochs, overall, 1
        self.test_deepspeed, self.optimizer, self.learn_rate, self.weight_decay, self.momentum = False, "sgd", 0.1, 0.0, 0.0
        self.criterion = torch.nn.BCEWithLogitsLoss()
        self._net = net
        self.b_size = {"train": 2, "val": 2, "test": 2}

    def get_network(self):
        return self._net


class _Pixels(Dataset):
    def __init__(self, n, seed, signal=True):
        g = torch.Generator().manual_seed(seed)
        self.x = torch.rand((n, 3, 4, 4), generator=g)
        self.m = (self.x[:, :1] > 0.5).float() if signal else (torch.rand((n, 1, 4, 4), generator=g) > 0.5).float()

    def __len__(self):
        return len(self.x)

    def __getitem__(self, i):
        return {"image": self.x[i], "mask": self.m[i], "index": i, "label": str(i)}


def _tiny_net():
    torch.manual_seed(0)
    return torch.nn.Conv2d(3, 1, 1)


def test_loop_logs_monitors_and_stops_early(tmp_path):
    """values reach model.logged whatever `lightning` is; best.ckpt follows val_loss; EarlyStopping with
    patience params.overall ends the run (the reference: EarlyStopping('val_loss', patience=params.overall))."""
    p = _Params(str(tmp_path), _tiny_net(), epochs=60, overall=3)
    model = T.RootLightningModel(p)
    # validation labels are noise: val_loss stops improving quickly while tr_loss keeps falling
    tr = DataLoader(_Pixels(16, 0), batch_size=2, shuffle=True)
    va = DataLoader(_Pixels(8, 1, signal=False), batch_size=2)
    loop = T._Loop(p, p.epochs, device=torch.device("cpu"))
    loop.fit(model, tr, va)
    hist = loop.history
    assert all("val_loss" in h and "tr_loss" in h and "tr_dice" in h for h in hist)
    assert loop.stopped_epoch is not None and len(hist) < 60
    best_epoch = min(range(len(hist)), key=lambda i: hist[i]["val_loss"])
    assert len(hist) - 1 - best_epoch == 3                       # stopped `overall` epochs after the best one
    ck = torch.load(os.path.join(p.save_path, "Checkpoints", "best.ckpt"), weights_only=False)
    assert ck["epoch"] == best_epoch and abs(ck["best"] - hist[best_epoch]["val_loss"]) < 1e-12
    last = torch.load(os.path.join(p.save_path, "Checkpoints", "last.ckpt"), weights_only=False)
    assert last["epoch"] == len(hist) - 1 and last["since_best"] == 3


def test_loop_resume_continues_epoch_and_best(tmp_path):
    p = _Params(str(tmp_path), _tiny_net(), epochs=3, overall=100)
    tr, va = DataLoader(_Pixels(8, 0), batch_size=2), DataLoader(_Pixels(4, 1), batch_size=2)
    a = T._Loop(p, 3, device=torch.device("cpu"))
    a.fit(T.RootLightningModel(p), tr, va)
    best3 = min(h["val_loss"] for h in a.history)
    p2 = _Params(str(tmp_path), _tiny_net(), epochs=5, overall=100)
    b = T._Loop(p2, 5, device=torch.device("cpu"))
    b.fit(T.RootLightningModel(p2), tr, va, ckpt_path=os.path.join(p.save_path, "Checkpoints", "last.ckpt"))
    assert [h["epoch"] for h in b.history] == [3, 4]             # not 0..4 again
    ck = torch.load(os.path.join(p.save_path, "Checkpoints", "last.ckpt"), weights_only=False)
    assert ck["best"] <= best3 + 1e-12                           # `best` carried over, never reset to inf


def test_loop_raises_when_monitored_metric_is_missing(tmp_path):
    p = _Params(str(tmp_path), _tiny_net(), epochs=1)
    model = T.RootLightningModel(p)
    model.validation_step = lambda batch, i: None                # logs nothing
    with pytest.raises(RuntimeError, match="val_loss"):
        T._Loop(p, 1, device=torch.device("cpu")).fit(model, DataLoader(_Pixels(4, 0), batch_size=2),
                                                      DataLoader(_Pixels(4, 1), batch_size=2))


def test_load_val_model_reads_lightning_layout_checkpoints(tmp_path):
    """A Lightning .ckpt pickles hyper_parameters={'params': <params object>} whose class lives in the training
    script's module tree; loading must neither need that class nor accept key mismatches."""
    p = ExpHyperspectralPRI(str(tmp_path), split_no=1, seed_num=0, comet_logging=False)
    p.cube_featmaps = 8
    src = T.RootLightningModel(p)
    with torch.no_grad():
        for t in src.parameters():
            t.add_(0.25)
    ck_dir = os.path.join(p.save_path, "Checkpoints")
    with pytest.raises(FileNotFoundError):
        T.load_val_model(p)
    os.makedirs(ck_dir)
    mod = types.ModuleType("scratch_training_script_params")

    class ExpThatWillNotExistAtLoadTime:
        def __init__(self):
            self.anything = [1, 2, 3]
    ExpThatWillNotExistAtLoadTime.__module__ = mod.__name__
    ExpThatWillNotExistAtLoadTime.__qualname__ = "ExpThatWillNotExistAtLoadTime"
    mod.ExpThatWillNotExistAtLoadTime = ExpThatWillNotExistAtLoadTime
    sys.modules[mod.__name__] = mod
    torch.save({"pytorch-lightning_version": "2.0.7", "epoch": 7, "state_dict": src.state_dict(),
                "hyper_parameters": {"params": ExpThatWillNotExistAtLoadTime()}},
               os.path.join(ck_dir, "epoch=7-val_loss=0.123-val_dice=0.800.ckpt"))
    torch.save({"state_dict": {k: torch.zeros_like(v) for k, v in src.state_dict().items()}},
               os.path.join(ck_dir, "last.ckpt"))
    del sys.modules[mod.__name__]
    got = T.load_val_model(p)                                       # the non-'last' file wins (PLTrainer.py:278-289)
    for (k, a), (_, b) in zip(src.state_dict().items(), got.state_dict().items()):
        assert torch.equal(a, b), k
    # a key mismatch is an error, not a silent partial load
    bad = dict(src.state_dict())
    bad.pop(next(iter(bad)))
    torch.save({"state_dict": bad}, os.path.join(ck_dir, "epoch=9-val_loss=0.1-val_dice=0.9.ckpt"))
    os.utime(os.path.join(ck_dir, "epoch=9-val_loss=0.1-val_dice=0.9.ckpt"), (2e9, 2e9))
    with pytest.raises(RuntimeError):
        T.load_val_model(p)


def test_plain_weight_files_and_rgb_run_directories(tmp_path):
    p = ExpRedGreenBluePRI(str(tmp_path), split_no=1, seed_num=0, comet_logging=False)
    assert p.translate_load_dir() == "UNET"
    p.change_network_param("UNET+", str(tmp_path), 1)
    assert p.translate_load_dir() == "UNET+" and "/UNET+/" in p.save_path     # 'UNET+' runs do not collide with 'UNET'
    p.model_name = "nope"
    with pytest.raises(ValueError):
        p.translate_load_dir()
    p.model_name = "UNET"
    src = T.RootLightningModel(p)
    os.makedirs(p.save_path, exist_ok=True)
    torch.save({"module." + k: v for k, v in src.m_network.state_dict().items()}, os.path.join(p.save_path, "best_wts.pt"))
    got = T.load_val_model(p)
    assert all(torch.equal(a, b) for a, b in zip(src.state_dict().values(), got.state_dict().values()))


def test_model_parallel_is_never_a_silent_synonym_of_dp(tmp_path):
    p = _Params(str(tmp_path), _tiny_net())
    p.device = "gpu"
    p.get_train_data = lambda: _Pixels(4, 0)
    p.get_val_data = lambda: _Pixels(4, 1)
    with pytest.raises(NotImplementedError, match="model-sharded SpectralUNET"):
        T.train_net(p, model_parallel=True)


# ------------------------------------------------------------------------------------------ world size 2, gloo
def _dp_worker(rank, world, port, root, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # replicas start DIFFERENT: the loop must broadcast rank 0's
    net = torch.nn.Conv2d(3, 1, 1)
    p = _Params(root, net, epochs=2, overall=100)
    model = T.RootLightningModel(p)
    seen = []
    orig = model.training_step

    def spy(batch, i):
        seen.extend(int(v) for v in batch["index"])
        return orig(batch, i)
    model.training_step = spy
    tr = DataLoader(_Pixels(8, 0), batch_size=2, shuffle=True)
    va = DataLoader(_Pixels(4, 1), batch_size=2)
    loop = T._Loop(p, 2, device=torch.device("cpu"))
    loop.fit(model, tr, va)
    q.put(to_plain((rank, seen, [t.detach().clone() for t in model.parameters()], loop.history)))
    dist.barrier()
    dist.destroy_process_group()


def test_loop_ddp_semantics_world2(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = {}
    for _ in procs:
        rank, seen, params, hist = from_plain(q.get(timeout=180))
        res[rank] = (seen, params, hist)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    s0, s1 = res[0][0], res[1][0]
    assert len(s0) == len(s1) == 8                              # 2 epochs x 4 samples per rank (8 samples / 2 ranks)
    for e in range(2):                                          # each epoch: the ranks' shards are disjoint and cover the set
        a, b = set(s0[4 * e: 4 * e + 4]), set(s1[4 * e: 4 * e + 4])
        assert not (a & b) and (a | b) == set(range(8))
    assert s0[:4] != s0[4:]                                     # set_epoch reshuffles
    for a, b in zip(res[0][1], res[1][1]):                      # replicas stay identical: same start, averaged gradients
        assert torch.allclose(a, b, atol=1e-7)
    assert res[0][2][-1]["val_loss"] == res[1][2][-1]["val_loss"]   # the monitored loss is a mean over ranks
    files = os.listdir(os.path.join(str(tmp_path), "Saved_Models", "HSI", "Tiny", "Run_1", "Checkpoints"))
    assert sorted(files) == ["best.ckpt", "last.ckpt"]
