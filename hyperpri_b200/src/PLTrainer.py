"""Trainer layer with the reference's entry points (reference src/PLTrainer.py:46-661):
``RootLightningModel(params)`` with ``training_step / validation_step / test_step / predict_step /
configure_optimizers`` and ``train_net / validate_net / test_net``.

PyTorch Lightning, torchmetrics and DeepSpeed are not installable offline, so the module re-hosts the
LightningModule contract on plain torch: when `lightning` is importable the class derives from
``pl.LightningModule`` and can be handed to a ``pl.Trainer`` unchanged; otherwise a small built-in
loop (`_Loop`) drives the same step methods.  Data parallelism is one process per GPU with the
engine's bucketed NCCL all-reduce (hyperpri_b200.parallel) instead of Lightning's "ddp" strategy: the training set is
sharded with a DistributedSampler, replicas start from rank 0's parameters, rank 0 alone writes checkpoints.
`model_parallel=True` (DeepSpeed ZeRO-2 in the reference, :409-433) selects the model-sharded SpectralUNET option
(hyperpri_b200.parallel.PixelParallel: each rank holds a strip of every image's pixels, per-image BatchNorm
statistics are all-reduced per layer) and raises for the other models -- it is never a silent synonym of "ddp".
"""
import io
import os
import pickle
import sys
import types
import warnings

import torch
import torch.optim as optim
from torch.utils.data import DataLoader

from .. import metrics as M
from .. import parallel
from ..prefetch import DevicePrefetcher
from ..zero_ckpt import consolidate_deepspeed_two   # noqa: F401  (the reference exposes it here, PLTrainer.py:186)
from .Experiments.models import *        # noqa: F401,F403  (the reference re-exports the models here, :26)

try:                                     # optional: real Lightning
    import lightning.pytorch as pl
    _Base = pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:                        # pragma: no cover - the offline image has no lightning
    pl = None
    _Base = torch.nn.Module
    HAVE_LIGHTNING = False

torch.backends.cudnn.allow_tf32 = True   # as the reference sets them (:32-34); the hot path does not use cuDNN
torch.set_float32_matmul_precision('high')


class RootLightningModel(_Base):
    def __init__(self, params):
        super().__init__()
        self.exp_params = params
        self.test_deepspeed = params.test_deepspeed
        self.p_optimizer, self.p_learn_rate = params.optimizer, params.learn_rate
        self.p_decay, self.p_momentum = params.weight_decay, params.momentum
        self.f_criterion = params.criterion
        self.m_network = params.get_network()
        self.save_segmaps = False
        self.threshold = 0.5
        self.predict_labels = []
        self.logged = {}
        self._driven_by_loop = False      # set by _Loop: values then go to self.logged even when Lightning is importable
        self._grad_scale = 1.0            # 1 / world size under the built-in data-parallel loop

    # ---- Lightning-compatible logging shim
    def log(self, name, value, **kw):
        if HAVE_LIGHTNING and not self._driven_by_loop and getattr(self, "_trainer", None) is not None:
            return super().log(name, value, **kw)
        self.logged.setdefault(name, []).append(value.detach() if torch.is_tensor(value) else value)

    def _fused_bce_ok(self, batch):
        """The configured criterion (params_HyperPRI.py:60,223: BCEWithLogitsLoss(), mean reduction, no weights) on an
        engine-backed network with CUDA inputs: network + loss + gradient + TP/FP/FN/TN run as one fused node."""
        c = self.f_criterion
        net = self.m_network
        return (type(c) is torch.nn.BCEWithLogitsLoss and c.reduction == 'mean' and c.weight is None
                and c.pos_weight is None and hasattr(net, "bce_step") and not getattr(net, "analyze", False)
                and torch.is_tensor(batch['image']) and batch['image'].is_cuda)

    def _pred(self, batch):
        out = self.m_network(batch['image'])
        if getattr(self.m_network, "analyze", False):
            out = out[0]
        return out

    def _seg_metrics(self, pred, mask, thr):
        seg = torch.sigmoid(pred.detach()) > thr
        c = M.confusion_counts(seg, mask)
        return M.accuracy(*c), M.dice(*c), M.jaccard(*c)

    def _step(self, batch, prefix, thr):
        if self._fused_bce_ok(batch):
            # PLTrainer.py:79-98 in one pass: hpri_bce_fwd_bwd gives the loss, its gradient and the counts the
            # Accuracy / Dice / Jaccard metrics are made of; sigmoid(pred) > thr is evaluated inside the kernel
            loss, pred, counts = self.m_network.bce_step(batch['image'], batch['mask'], thr=float(thr),
                                                         grad_scale=self._grad_scale)
            c = [v.double() for v in counts.unbind(0)]
            acc, dice, iou = M.accuracy(*c), M.dice(*c), M.jaccard(*c)
            ret = loss
        else:
            mask = batch['mask'].to(torch.int32)
            pred = self._pred(batch)
            loss = self.f_criterion(pred, batch['mask'])
            acc, dice, iou = self._seg_metrics(pred, mask, thr)
            ret = loss * self._grad_scale if self._grad_scale != 1.0 else loss
        self.log(f'{prefix}_loss', loss, on_step=False, on_epoch=True, sync_dist=True)
        self.log(f'{prefix}_acc', acc, on_step=False, on_epoch=True, sync_dist=False, prog_bar=False)
        self.log(f'{prefix}_dice', dice, on_step=False, on_epoch=True, sync_dist=prefix != 'tr', prog_bar=True)
        self.log(f'{prefix}_pos_iou', iou, on_step=False, on_epoch=True, sync_dist=False, prog_bar=False)
        return ret, pred

    def training_step(self, batch, batch_idx):
        return self._step(batch, 'tr', self.threshold)[0]

    def validation_step(self, batch, batch_idx):
        self._step(batch, 'val', 0.5)

    def test_step(self, batch, batch_idx):
        return self._step(batch, 'test', self.threshold)[1]

    def predict_step(self, batch, batch_idx, dataloader_idx=0):
        self.predict_labels.append(batch['mask'].cpu())
        return self._pred(batch).cpu()

    def configure_optimizers(self):
        name = self.p_optimizer.upper()
        if name == 'ADAM':
            params = list(self.m_network.parameters())
            if params and params[0].is_cuda:       # one native launch per step (same update rule and state_dict)
                from ..optim import FusedAdam
                found_inf = None
                if hasattr(self.m_network, "_get_engine"):   # skip (on the device) a step whose fp16 gradients overflowed
                    found_inf = self.m_network._get_engine(params[0].device).overflow
                return FusedAdam(params, lr=self.p_learn_rate, weight_decay=self.p_decay, found_inf=found_inf)
            return optim.Adam(params, lr=self.p_learn_rate, weight_decay=self.p_decay)
        if name == 'SGD':
            return optim.SGD(self.m_network.parameters(), lr=self.p_learn_rate, momentum=self.p_momentum,
                             weight_decay=self.p_decay)
        raise ValueError(f'Unknown Optimizer name: {name}')


def _to_device(batch, dev):
    """Batches come out of a DevicePrefetcher already on the device (copied one batch ahead on a copy stream); an fp16
    cube (HyperpriDataset(host_dtype=float16)) stays fp16 -- the ingest kernel reads it as is."""
    def conv(k, v):
        if not torch.is_tensor(v) or k not in ('image', 'mask'):
            return v
        v = v.to(dev, non_blocking=True)
        return v if (k == 'image' and v.dtype == torch.float16) else v.float()
    return {k: conv(k, v) for k, v in batch.items()}


def _dist_on():
    return torch.distributed.is_available() and torch.distributed.is_initialized()


class _Loop:
    """Minimal stand-in for pl.Trainer: fit / predict over DataLoaders on one device per process.  Reproduces the parts
    of the reference's trainer configuration that change results (PLTrainer.py:345-352, 434-450): EarlyStopping on
    val_loss with patience params.overall, best-val_loss and last checkpoints, "ddp" semantics under torchrun (sharded
    sampler, replicas synchronised at start, rank 0 writes)."""

    def __init__(self, params, max_epochs, device=None, strategy="dp"):
        self.params, self.max_epochs, self.strategy = params, max_epochs, strategy
        self.device = device or torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        self.history = []
        self.stopped_epoch = None
        self.rank = torch.distributed.get_rank() if _dist_on() else 0

    def _sharded(self, loader, epoch):
        """Lightning's "ddp" replaces the train sampler by a DistributedSampler (each rank sees 1/world of the data)."""
        if not _dist_on() or self.strategy != "dp" or torch.distributed.get_world_size() == 1:
            return loader
        from torch.utils.data.distributed import DistributedSampler
        if isinstance(getattr(loader, "sampler", None), DistributedSampler):
            loader.sampler.set_epoch(epoch)
            return loader
        sharded = getattr(self, "_sharded_loader", None)
        if sharded is None or sharded[0] is not loader:
            shuffle = isinstance(getattr(loader, "sampler", None), torch.utils.data.RandomSampler)
            smp = DistributedSampler(loader.dataset, shuffle=shuffle, seed=int(getattr(self.params, "run_num", 0)))
            new = DataLoader(loader.dataset, batch_size=loader.batch_size, sampler=smp, num_workers=loader.num_workers,
                             collate_fn=loader.collate_fn, drop_last=loader.drop_last)
            self._sharded_loader = sharded = (loader, new, smp)
        sharded[2].set_epoch(epoch)
        return sharded[1]

    @staticmethod
    def _epoch_means(logged):
        return {k: float(torch.stack([torch.as_tensor(v).float().cpu() for v in vs]).mean()) for k, vs in logged.items()}

    def fit(self, model, train_loader, val_loader=None, ckpt_path=None):
        model.to(self.device)
        model._driven_by_loop = True
        opt = model.configure_optimizers()
        start_epoch, best, since_best = 0, float("inf"), 0
        if ckpt_path and os.path.exists(ckpt_path):
            ck = torch.load(ckpt_path, map_location=self.device, weights_only=False)     # our own checkpoint format
            model.load_state_dict(ck["state_dict"]); opt.load_state_dict(ck["optimizer"])
            start_epoch = int(ck.get("epoch", -1)) + 1
            best, since_best = float(ck.get("best", float("inf"))), int(ck.get("since_best", 0))
        red = None
        if _dist_on():
            with torch.no_grad():           # replicas start from rank 0's weights and buffers, as DDP guarantees
                for t in list(model.parameters()) + list(model.buffers()):
                    torch.distributed.broadcast(t.data, src=0)
            if hasattr(model.m_network, "_get_engine"):
                model.m_network._get_engine(self.device).invalidate_packed()
            if self.strategy == "dp" and hasattr(model.m_network, "_get_engine"):
                red = parallel.attach(model.m_network._get_engine(self.device))
                model._grad_scale = red.grad_scale()
        monitor = "val_loss" if val_loader is not None else "tr_loss"
        ck_dir = os.path.join(self.params.save_path, 'Checkpoints')
        skipped_seen = 0
        for epoch in range(start_epoch, self.max_epochs):
            model.train(); model.logged.clear()
            pf = DevicePrefetcher(self._sharded(train_loader, epoch), self.device)
            for i, batch in enumerate(pf):
                opt.zero_grad(set_to_none=True)
                if hasattr(model.m_network, "set_next_input"):     # the staged next batch is ingested under this backward
                    nb = pf.next_batch
                    model.m_network.set_next_input(nb["image"] if nb is not None else None, pf.next_ready)
                loss = model.training_step(_to_device(batch, self.device), i)
                loss.backward()
                if red is not None:
                    red.finish()
                elif _dist_on() and self.strategy == "dp":        # a network without the engine: plain gradient averaging
                    world = torch.distributed.get_world_size()
                    for p_ in model.parameters():
                        if p_.grad is not None:
                            torch.distributed.all_reduce(p_.grad)
                            p_.grad.div_(world)
                opt.step()
            row = self._epoch_means(model.logged)
            if val_loader is not None:
                model.eval(); model.logged.clear()
                with torch.no_grad():
                    for i, batch in enumerate(DevicePrefetcher(val_loader, self.device)):
                        model.validation_step(_to_device(batch, self.device), i)
                row.update(self._epoch_means(model.logged))
            if _dist_on() and torch.distributed.get_world_size() > 1:      # sync_dist=True: the monitored loss is a mean over ranks
                keys = sorted(k for k in row if k.endswith("_loss") or k == "val_dice")
                t = torch.tensor([row[k] for k in keys], dtype=torch.float64, device=self.device)
                torch.distributed.all_reduce(t)
                for k, v in zip(keys, (t / torch.distributed.get_world_size()).tolist()):
                    row[k] = v
            row["epoch"] = epoch
            self.history.append(row)
            if monitor not in row:
                raise RuntimeError(f"monitored metric '{monitor}' was not logged during epoch {epoch} (logged: {sorted(row)})")
            improved = row[monitor] < best
            if improved:
                best, since_best = row[monitor], 0
            else:
                since_best += 1
            skipped = opt.skipped_steps() if hasattr(opt, "skipped_steps") else 0
            if skipped > skipped_seen:          # fp16 gradient overflow: those steps were skipped; lower the loss scale
                eng = model.m_network._get_engine(self.device)
                eng.loss_scale_shift += 1
                warnings.warn(f"{skipped - skipped_seen} optimizer step(s) skipped in epoch {epoch}: loss-scaled fp16 gradients "
                              f"overflowed; the loss scale is halved (shift {eng.loss_scale_shift})")
                skipped_seen = skipped
            if self.rank == 0:
                os.makedirs(ck_dir, exist_ok=True)
                state = {"state_dict": model.state_dict(), "optimizer": opt.state_dict(), "epoch": epoch, "best": best,
                         "since_best": since_best}
                torch.save(state, os.path.join(ck_dir, 'last.ckpt'))
                if improved:
                    torch.save(state, os.path.join(ck_dir, 'best.ckpt'))
            if _dist_on():
                torch.distributed.barrier()
            if since_best >= int(getattr(self.params, "overall", 0) or self.max_epochs + 1):   # EarlyStopping(patience=overall)
                self.stopped_epoch = epoch
                break
        return self

    def predict(self, model, loader, return_predictions=True):
        model.to(self.device).eval()
        model._driven_by_loop = True
        out = []
        with torch.no_grad():
            for i, batch in enumerate(DevicePrefetcher(loader, self.device)):
                out.append(model.predict_step(_to_device(batch, self.device), i))
        return out if return_predictions else None


class _TensorsOnlyUnpickler(pickle.Unpickler):
    """Lightning checkpoints pickle `hyper_parameters = {'params': <ExpHyperspectralPRI>}` (save_hyperparameters), i.e.
    objects of classes that live in the TRAINING script's module tree (`src.Experiments.params_HyperPRI`, torchvision
    transforms, ...).  Only the tensors are wanted here: any class that cannot be imported is replaced by an inert stub
    instead of failing (torch.load(weights_only=True) rejects such files, weights_only=False needs the classes)."""

    def find_class(self, mod, name):
        try:
            return super().find_class(mod, name)
        except Exception:
            return type(name, (), {"__init__": lambda self, *a, **k: None, "__setstate__": lambda self, st: None,
                                   "__reduce_ex__": None})


_tensor_pickle = types.ModuleType("hyperpri_b200_tensor_pickle")
_tensor_pickle.Unpickler = _TensorsOnlyUnpickler
_tensor_pickle.load = lambda f, **kw: _TensorsOnlyUnpickler(f, **kw).load()
_tensor_pickle.loads = lambda b, **kw: _TensorsOnlyUnpickler(io.BytesIO(b), **kw).load()
_tensor_pickle.__dict__.update({k: getattr(pickle, k) for k in ("dump", "dumps", "Pickler", "PickleError", "UnpicklingError",
                                                                 "HIGHEST_PROTOCOL", "DEFAULT_PROTOCOL")})


def _find_val_checkpoint(params):
    """PLTrainer.py:274-292: the newest non-'last' file under Checkpoints (else last.ckpt), else best_wts.pt."""
    ck_dir = os.path.join(params.save_path, 'Checkpoints')
    if os.path.isdir(ck_dir):
        names = os.listdir(ck_dir)
        rest = sorted((os.path.join(ck_dir, f) for f in names if 'last' not in f), key=os.path.getmtime)
        if rest:
            return rest[-1]
        if 'last.ckpt' in names:
            return os.path.join(ck_dir, 'last.ckpt')
        if names:
            return os.path.join(ck_dir, names[0])
    wts = os.path.join(params.save_path, 'best_wts.pt')
    if os.path.exists(wts):
        return wts
    raise FileNotFoundError(f"no checkpoint under {ck_dir} and no {wts}")


def load_val_model(params, device=None):
    """PLTrainer.py:270-330: the model restored from the best val_loss checkpoint (Lightning .ckpt with `m_network.`
    keys, a plain state dict with optional `module.` prefix, or a DeepSpeed ZeRO-2 directory).  Keys must match exactly
    after the prefix mapping; a missing checkpoint raises FileNotFoundError."""
    path = _find_val_checkpoint(params)
    print(f"   LOADING FROM CKPT FILE: {path}")
    model = RootLightningModel(params)
    if os.path.isdir(path):                       # a DeepSpeed ZeRO-2 checkpoint directory (PLTrainer.py:297-307)
        model.m_network.load_state_dict(consolidate_deepspeed_two(path))
        return model
    raw = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_tensor_pickle)
    if isinstance(raw, dict) and "state_dict" in raw:            # Lightning (or our _Loop) checkpoint
        sd = {k.replace("_forward_module.", "", 1): v for k, v in raw["state_dict"].items()}
    else:                                          # plain weights: `module.` (DataParallel) or bare keys (:316-323)
        sd = {("m_network." + k.replace("module.", "", 1)): v for k, v in raw.items()}
    model.load_state_dict(sd, strict=True)
    return model


def train_net(params, checkpoint=None, model_parallel: bool = False):
    """PLTrainer.py:333-460: loaders (batch b_size, shuffle, num_workers=0), fit for params.epochs."""
    train_loader = DataLoader(params.get_train_data(), batch_size=params.b_size['train'], shuffle=True, num_workers=0)
    val_loader = DataLoader(params.get_val_data(), batch_size=params.b_size['val'], shuffle=False, num_workers=0)
    model = RootLightningModel(params)
    if getattr(params, "deterministic", False) and getattr(params, "device", "gpu") == "gpu":
        # the reference builds every Trainer with deterministic='warn' (:430,439,447); here it is a knob, default off
        from .. import ops as _ops
        _ops.set_deterministic(True)
    ckpt = None
    if checkpoint:
        last = os.path.join(params.save_path, 'Checkpoints', 'last.ckpt')
        ckpt = last if os.path.exists(last) else None
    strategy = "dp"
    if model_parallel and getattr(params, "device", "gpu") == "gpu":
        # the reference's MODEL_SHARD run (DeepSpeed ZeRO-2 + bf16-mixed, :409-433; README: SpectralUNET on >= 2 GPUs)
        if not hasattr(model.m_network, "enable_pixel_parallel"):
            raise NotImplementedError(
                f"model_parallel=True is the model-sharded SpectralUNET option; {type(model.m_network).__name__} fits one "
                "B200 and runs data-parallel: call train_net(params, model_parallel=False) under torchrun")
        if not _dist_on():
            raise RuntimeError("model_parallel=True needs an initialised torch.distributed process group (torchrun, "
                               "one rank per GPU)")
        model.m_network.enable_pixel_parallel(torch.distributed.group.WORLD)
        strategy = "pixel_parallel"
    trainer = _Loop(params, params.epochs, strategy=strategy)
    trainer.fit(model, train_loader, val_loader, ckpt_path=ckpt)
    trainer.model = model
    return trainer


def _best_model(params, pl_trainer):
    """The reference always evaluates the best val_loss checkpoint, whatever trainer object it is handed
    (PLTrainer.py:476, 622: load_val_model(params)); the trainer's in-memory (last-epoch) model is used only when
    nothing was written to disk."""
    try:
        model = load_val_model(params)
    except FileNotFoundError:
        model = getattr(pl_trainer, "model", None)
        if model is None:
            raise
    return model


def _collect(model, loader, trainer):
    model.predict_labels = []
    preds = trainer.predict(model, loader, return_predictions=True)
    logits = torch.cat(preds, dim=0).flatten()
    masks = torch.cat(model.predict_labels, dim=0).flatten()
    return logits, masks


def _sweep_device(model, loader, trainer):
    """Prediction sweep that keeps everything on the GPU: per batch, forward + one histogram kernel
    (metrics.DevicePRCurve).  The reference moves every prediction to the host and concatenates (:142-162, :538)."""
    model.to(trainer.device).eval()
    curve = M.DevicePRCurve(trainer.device, 500)
    with torch.no_grad():
        for batch in DevicePrefetcher(loader, trainer.device):
            b = _to_device(batch, trainer.device)
            curve.update(model._pred(b), b['mask'])
    return curve


def _use_device_sweep(trainer, save_segmaps):
    if save_segmaps:
        warnings.warn("save_segmaps=True: rendering segmentation maps to PNG (reference PLTrainer.py:219-267, matplotlib) is "
                      "outside this path; predictions are collected on the host as in the reference, no figures are written")
    return isinstance(trainer, _Loop) and trainer.device.type == "cuda" and not save_segmaps


def validate_net(val_data, params, pl_trainer=None, save_segmaps=False):
    """PLTrainer.py:463-609: predict, BCE, 500-threshold PR curve, best-Dice threshold, Acc/IoU/AP/confusion.
    Returns (precision, recall, thresholds)."""
    loader = DataLoader(val_data, batch_size=params.b_size['test'], shuffle=False)
    model = _best_model(params, pl_trainer)
    trainer = pl_trainer if isinstance(pl_trainer, _Loop) else _Loop(params, 0)
    curve = None
    if _use_device_sweep(trainer, save_segmaps):
        curve = _sweep_device(model, loader, trainer)
        bce = curve.bce_loss()
        prec, rec, thr = curve.compute()
    else:
        logits, masks = _collect(model, loader, trainer)
        bce = params.criterion(logits, masks.float())
        probs = torch.sigmoid(logits)
        prec, rec, thr = M.binned_pr_curve(probs, masks, 500)
    crop = int(len(prec) // 100)
    tp_, tr_, tt_ = prec[crop:-crop], rec[crop:-crop], thr[crop:-crop]     # top/bottom thresholds excluded (:547-550)
    dice_curve = 2 * tp_ * tr_ / (tp_ + tr_).clamp_min(1e-30)
    bi = torch.argmax(dice_curve)
    best_thr = torch.round(tt_[min(bi, len(tt_) - 1)].float(), decimals=2)
    c = curve.counts_at(best_thr) if curve is not None else M.confusion_counts(probs > best_thr, masks)
    ap = M.average_precision(prec, rec)
    print(f"\n{params.model_name}\n   Best Threshold {best_thr:.3f}:")
    print(f"      BCE Loss : {bce:.3f}\n      Pixel Acc: {M.accuracy(*c):.3f}\n      Precision: {tp_[bi]:.3f}")
    print(f"      Recall   : {tr_[bi]:.3f}\n      DICE     : {dice_curve[bi]:.3f}\n      +IOU     : {M.jaccard(*c):.3f}")
    print(f"      Avg Prec : {ap:.3f}\n")
    tp, fp, fn, tn = [float(v) for v in c]
    print(f"      Conf Mat : {[tn / max(tn + fp, 1), fp / max(tn + fp, 1)]}")
    print(f"                 {[fn / max(fn + tp, 1), tp / max(fn + tp, 1)]}")
    if prec[-2] < 1e-6:
        prec[-2] = (1 + prec[-3]) / 2
    model.threshold = best_thr
    if getattr(pl_trainer, "model", None) is not None:
        pl_trainer.model.threshold = best_thr
    return prec, rec, thr


def test_net(test_data, params, best_threshold, pl_trainer=None, save_segmaps=False):
    """PLTrainer.py:612-661: metrics at a fixed threshold."""
    loader = DataLoader(test_data, batch_size=params.b_size['test'], shuffle=False)
    model = _best_model(params, pl_trainer)
    trainer = pl_trainer if isinstance(pl_trainer, _Loop) else _Loop(params, 0)
    thr2 = round(float(best_threshold), 2)
    if _use_device_sweep(trainer, save_segmaps) and abs(thr2 - float(best_threshold)) < 1e-6:
        curve = _sweep_device(model, loader, trainer)
        c = curve.counts_at(thr2)
        prec, rec, _ = curve.compute()
    else:
        logits, masks = _collect(model, loader, trainer)
        probs = torch.sigmoid(logits)
        c = M.confusion_counts(probs > best_threshold, masks)
        prec, rec, _ = M.binned_pr_curve(probs, masks, 500)
    out = {"acc": float(M.accuracy(*c)), "dice": float(M.dice(*c)), "pos_iou": float(M.jaccard(*c)),
           "avg_prec": float(M.average_precision(prec, rec))}
    print(f"Threshold {float(best_threshold):.3f}:")
    for k, v in out.items():
        print(f"      {k:9s}: {v:.3f}")
    return out
