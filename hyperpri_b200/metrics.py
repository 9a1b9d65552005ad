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


class DevicePRCurve:
    """Validation maths without moving predictions to the host (reference PLTrainer.py:538-583 concatenates every
    prediction on the CPU): per batch one kernel accumulates the 500-bin probability histograms by mask class, a
    second family of bins at the two-decimal thresholds the best-Dice search can return, and the BCE sum.
    compute() gives the same (precision, recall, thresholds) as binned_pr_curve; counts_at(thr) the confusion
    counts for seg = p > thr at any two-decimal threshold."""

    def __init__(self, device, n_thresholds: int = 500):
        from . import ops
        self._ops = ops
        self.thr = torch.linspace(0, 1, n_thresholds, device=device)
        self.cut = torch.round(torch.arange(0, 101, device=device, dtype=torch.float32) / 100, decimals=2)
        z = lambda n: torch.zeros(n, dtype=torch.int64, device=device)
        self.hp, self.hn = z(n_thresholds), z(n_thresholds)
        self.cp, self.cn = z(102), z(102)
        self.bce = torch.zeros((), dtype=torch.float64, device=device)
        self.numel = 0

    def update(self, logits: torch.Tensor, mask: torch.Tensor):
        lg = logits.detach().reshape(-1).float().contiguous()
        mk = mask.reshape(-1).float().contiguous()
        self._ops.pr_hist(lg, mk, self.thr, self.cut, self.hp, self.hn, self.cp, self.cn, self.bce)
        self.numel += lg.numel()

    def bce_loss(self):
        return (self.bce / max(self.numel, 1)).float()

    def compute(self):
        hp, hn = self.hp.double(), self.hn.double()
        tps = hp.flip(0).cumsum(0).flip(0)
        fps = hn.flip(0).cumsum(0).flip(0)
        npos = hp.sum()
        precision = torch.where(tps + fps > 0, tps / (tps + fps).clamp_min(1e-30), torch.zeros_like(tps))
        recall = torch.where(npos > 0, tps / npos.clamp_min(1e-30), torch.zeros_like(tps))
        one = torch.ones(1, dtype=precision.dtype, device=precision.device)
        return (torch.cat([precision, one]).float(), torch.cat([recall, torch.zeros_like(one)]).float(), self.thr)

    def counts_at(self, threshold):
        """(tp, fp, fn, tn) for seg = sigmoid(logit) > threshold; threshold must be one of 0.00, 0.01, ..., 1.00."""
        k = int(torch.argmin((self.cut - float(threshold)).abs()).item())
        if abs(float(self.cut[k]) - float(threshold)) > 1e-6:
            raise ValueError("counts_at needs a two-decimal threshold")
        cp, cn = self.cp.double(), self.cn.double()
        tp, fp = cp[k + 1:].sum(), cn[k + 1:].sum()          # bins k+1.. hold p > cut[k]
        return tp, fp, cp.sum() - tp, cn.sum() - fp
