"""Trainer layer with the reference's entry points (reference src/PLTrainer.py:46-661):
``RootLightningModel(params)`` with ``training_step / validation_step / test_step / predict_step /
configure_optimizers`` and ``train_net / validate_net / test_net``.

PyTorch Lightning, torchmetrics and DeepSpeed are not installable offline, so the module re-hosts the
LightningModule contract on plain torch: when `lightning` is importable the class derives from
``pl.LightningModule`` and can be handed to a ``pl.Trainer`` unchanged; otherwise a small built-in
loop (`_Loop`) drives the same step methods.  Data parallelism is one process per GPU with the
engine's bucketed NCCL all-reduce (hyperpri_b200.parallel) instead of Lightning's "ddp" strategy;
`model_parallel=True` (DeepSpeed ZeRO-2 in the reference, :409-433) is accepted and mapped to the same
data-parallel path (ZeRO-2 is data parallelism with sharded optimizer state).
"""
import os
import sys

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

    # ---- Lightning-compatible logging shim
    def log(self, name, value, **kw):
        if HAVE_LIGHTNING:
            return super().log(name, value, **kw)
        self.logged.setdefault(name, []).append(value.detach() if torch.is_tensor(value) else value)

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
        mask = batch['mask'].to(torch.int32)
        pred = self._pred(batch)
        loss = self.f_criterion(pred, batch['mask'])
        acc, dice, iou = self._seg_metrics(pred, mask, thr)
        self.log(f'{prefix}_loss', loss, on_step=False, on_epoch=True, sync_dist=True)
        self.log(f'{prefix}_acc', acc, on_step=False, on_epoch=True, sync_dist=False, prog_bar=False)
        self.log(f'{prefix}_dice', dice, on_step=False, on_epoch=True, sync_dist=prefix != 'tr', prog_bar=True)
        self.log(f'{prefix}_pos_iou', iou, on_step=False, on_epoch=True, sync_dist=False, prog_bar=False)
        return loss, pred

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
                return FusedAdam(params, lr=self.p_learn_rate, weight_decay=self.p_decay)
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


class _Loop:
    """Minimal stand-in for pl.Trainer: fit / predict over DataLoaders on one device per process."""

    def __init__(self, params, max_epochs, device=None):
        self.params, self.max_epochs = params, max_epochs
        self.device = device or torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        self.history = []

    def fit(self, model, train_loader, val_loader=None, ckpt_path=None):
        model.to(self.device)
        opt = model.configure_optimizers()
        if ckpt_path and os.path.exists(ckpt_path):
            ck = torch.load(ckpt_path, map_location=self.device)
            model.load_state_dict(ck["state_dict"]); opt.load_state_dict(ck["optimizer"])
        red = None
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            red = parallel.attach(model.m_network._get_engine(self.device))
        best = float("inf")
        for epoch in range(self.max_epochs):
            model.train(); model.logged.clear()
            for i, batch in enumerate(DevicePrefetcher(train_loader, self.device)):
                opt.zero_grad(set_to_none=True)
                loss = model.training_step(_to_device(batch, self.device), i)
                if red is not None:
                    loss = loss * red.grad_scale()
                loss.backward()
                if red is not None:
                    red.finish()
                opt.step()
            row = {k: float(torch.stack([torch.as_tensor(v).float().cpu() for v in vs]).mean()) for k, vs in model.logged.items()}
            if val_loader is not None:
                model.eval(); model.logged.clear()
                with torch.no_grad():
                    for i, batch in enumerate(DevicePrefetcher(val_loader, self.device)):
                        model.validation_step(_to_device(batch, self.device), i)
                row.update({k: float(torch.stack([torch.as_tensor(v).float().cpu() for v in vs]).mean()) for k, vs in model.logged.items()})
            row["epoch"] = epoch
            self.history.append(row)
            os.makedirs(os.path.join(self.params.save_path, 'Checkpoints'), exist_ok=True)
            state = {"state_dict": model.state_dict(), "optimizer": opt.state_dict(), "epoch": epoch}
            torch.save(state, os.path.join(self.params.save_path, 'Checkpoints', 'last.ckpt'))
            if row.get("val_loss", row.get("tr_loss", 0.0)) < best:
                best = row.get("val_loss", row.get("tr_loss", 0.0))
                torch.save(state, os.path.join(self.params.save_path, 'Checkpoints', 'best.ckpt'))
        return self

    def predict(self, model, loader, return_predictions=True):
        model.to(self.device).eval()
        out = []
        with torch.no_grad():
            for i, batch in enumerate(DevicePrefetcher(loader, self.device)):
                out.append(model.predict_step(_to_device(batch, self.device), i))
        return out if return_predictions else None


def load_val_model(params, device=None):
    """Newest checkpoint under <save_path>/Checkpoints, else best_wts.pt (PLTrainer.py:270-330)."""
    model = RootLightningModel(params)
    ck_dir = os.path.join(params.save_path, 'Checkpoints')
    cand = []
    if os.path.isdir(ck_dir):
        cand = sorted((os.path.join(ck_dir, f) for f in os.listdir(ck_dir) if 'last' not in f), key=os.path.getmtime)
        if not cand:
            cand = [os.path.join(ck_dir, f) for f in os.listdir(ck_dir)]
    wts = os.path.join(params.save_path, 'best_wts.pt')
    if cand and os.path.isdir(cand[-1]):          # a DeepSpeed ZeRO-2 checkpoint directory (PLTrainer.py:297-307)
        model.m_network.load_state_dict(consolidate_deepspeed_two(cand[-1]))
    elif cand:
        ck = torch.load(cand[-1], map_location="cpu")
        sd = ck.get("state_dict", ck)
        sd = {k.replace("_forward_module.", ""): v for k, v in sd.items()}
        model.load_state_dict(sd, strict=False)
    elif os.path.exists(wts):
        sd = torch.load(wts, map_location="cpu")
        model.m_network.load_state_dict({k.replace("module.", "", 1): v for k, v in sd.items()})
    return model


def train_net(params, checkpoint=None, model_parallel: bool = False):
    """PLTrainer.py:333-460: loaders (batch b_size, shuffle, num_workers=0), fit for params.epochs."""
    train_loader = DataLoader(params.get_train_data(), batch_size=params.b_size['train'], shuffle=True, num_workers=0)
    val_loader = DataLoader(params.get_val_data(), batch_size=params.b_size['val'], shuffle=False, num_workers=0)
    model = RootLightningModel(params)
    ckpt = None
    if checkpoint:
        last = os.path.join(params.save_path, 'Checkpoints', 'last.ckpt')
        ckpt = last if os.path.exists(last) else None
    trainer = _Loop(params, params.epochs)
    trainer.fit(model, train_loader, val_loader, ckpt_path=ckpt)
    trainer.model = model
    return trainer


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
    return isinstance(trainer, _Loop) and trainer.device.type == "cuda" and not save_segmaps


def validate_net(val_data, params, pl_trainer=None, save_segmaps=False):
    """PLTrainer.py:463-609: predict, BCE, 500-threshold PR curve, best-Dice threshold, Acc/IoU/AP/confusion.
    Returns (precision, recall, thresholds)."""
    loader = DataLoader(val_data, batch_size=params.b_size['test'], shuffle=False)
    model = getattr(pl_trainer, "model", None) or load_val_model(params)
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
    return prec, rec, thr


def test_net(test_data, params, best_threshold, pl_trainer=None, save_segmaps=False):
    """PLTrainer.py:612-661: metrics at a fixed threshold."""
    loader = DataLoader(test_data, batch_size=params.b_size['test'], shuffle=False)
    model = getattr(pl_trainer, "model", None) or load_val_model(params)
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
