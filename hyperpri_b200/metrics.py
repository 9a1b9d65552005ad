"""Segmentation metrics restated on plain torch (torchmetrics 1.2.0, which the reference uses at
PLTrainer.py:62-68,538-583, is not installed): counts-based accuracy / Dice / +IoU and the binned
precision-recall curve with `thresholds=500`."""
from __future__ import annotations

import torch


def confusion_counts(seg: torch.Tensor, mask: torch.Tensor):
    seg = seg.reshape(-1).bool()
    m = mask.reshape(-1) > 0
    tp = (seg & m).sum().double(); fp = (seg & ~m).sum().double()
    fn = (~seg & m).sum().double(); tn = (~seg & ~m).sum().double()
    return tp, fp, fn, tn


def accuracy(tp, fp, fn, tn):
    return ((tp + tn) / (tp + fp + fn + tn)).float()


def dice(tp, fp, fn, tn, zero_division=1e-12):
    den = 2 * tp + fp + fn
    return torch.where(den > 0, 2 * tp / den.clamp_min(1e-30), torch.as_tensor(zero_division, dtype=den.dtype)).float()


def jaccard(tp, fp, fn, tn):
    den = tp + fp + fn
    return torch.where(den > 0, tp / den.clamp_min(1e-30), torch.zeros_like(den)).float()


def binned_pr_curve(probs: torch.Tensor, target: torch.Tensor, n_thresholds: int = 500):
    """precision[n+1], recall[n+1], thresholds[n] like torchmetrics' binned binary PR curve:
    thresholds = linspace(0,1,n); prediction positive when prob >= thr; last point is (1, 0)."""
    probs = probs.reshape(-1).float()
    pos = target.reshape(-1) > 0
    thr = torch.linspace(0, 1, n_thresholds, device=probs.device)
    # histogram of probabilities into the threshold bins, separately for positives and negatives
    idx = torch.bucketize(probs, thr, right=True) - 1           # largest i with thr[i] <= p
    idx = idx.clamp_(0, n_thresholds - 1)
    hp = torch.bincount(idx[pos], minlength=n_thresholds).double()
    hn = torch.bincount(idx[~pos], minlength=n_thresholds).double()
    tps = hp.flip(0).cumsum(0).flip(0)                           # count with p >= thr[i]
    fps = hn.flip(0).cumsum(0).flip(0)
    npos = pos.sum().double()
    precision = torch.where(tps + fps > 0, tps / (tps + fps).clamp_min(1e-30), torch.zeros_like(tps))
    recall = torch.where(npos > 0, tps / npos.clamp_min(1e-30), torch.zeros_like(tps))
    precision = torch.cat([precision, torch.ones(1, dtype=precision.dtype, device=probs.device)])
    recall = torch.cat([recall, torch.zeros(1, dtype=recall.dtype, device=probs.device)])
    return precision.float(), recall.float(), thr


def average_precision(precision, recall):
    """-sum((r[i+1]-r[i]) * p[i]) over the curve (torchmetrics' reduction)."""
    return -torch.sum((recall[1:] - recall[:-1]) * precision[:-1])
