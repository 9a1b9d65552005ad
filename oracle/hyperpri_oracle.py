"""CPU oracle for the HyperPRI segmentation hot path -- TEST INFRASTRUCTURE ONLY.

This file is a functional fp32 restatement of the reference's arithmetic for the path
SURVEY.md section 8 names (ingest -> UNet / CubeNET / SpectralUNET forward -> sigmoid-BCE).
It is written with plain torch CPU tensor ops driven by a *state dict* (no nn.Module), so
autograd supplies the backward pass that the CUDA kernels are checked against.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  The product (``hyperpri_b200``) never does.

Pinning: the reference ships no tests or golden vectors.  This oracle is
pinned against the reference's own modules imported from /root/reference in the build
container (``oracle/gen_golden.py`` -> ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
replays them without the reference being present).

Each function cites the reference lines (relative to the reference repository root) it follows.
"""
from __future__ import annotations

import zlib
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5       # torch default, reference passes none (model_parts.py:23,26)
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------
# deterministic synthetic parameters (shared by golden generation, tests and bench)
# --------------------------------------------------------------------------------------
def _rs(key: str, seed: int) -> np.random.RandomState:
    return np.random.RandomState((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def synth_state_dict(schema: Dict[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Fill a {key: shape} schema with reproducible values, independent of module
    construction order and of torch's RNG.  Weights/biases use the kaiming-uniform(a=sqrt 5)
    bound 1/sqrt(fan_in) torch's default init gives Conv/Linear; BatchNorm affine terms are
    drawn away from (1, 0) so parity tests exercise them; running stats are non-trivial."""
    out: Dict[str, torch.Tensor] = {}
    fan_of: Dict[str, int] = {}
    for k, shp in schema.items():
        if k.endswith("weight") and len(shp) >= 2:
            # ConvTranspose2d weight is (Cin, Cout, kh, kw): torch computes fan_in from dim 1
            fan_of[k[: -len("weight")]] = int(np.prod(shp[1:]))
    for k, shp in schema.items():
        r = _rs(k, seed)
        pre = k.rsplit(".", 1)[0] + "."
        if k.endswith("num_batches_tracked"):
            out[k] = torch.tensor(0, dtype=torch.long)
        elif k.endswith("running_mean"):
            out[k] = torch.from_numpy(r.uniform(-0.1, 0.1, shp).astype(np.float32))
        elif k.endswith("running_var"):
            out[k] = torch.from_numpy(r.uniform(0.5, 1.5, shp).astype(np.float32))
        elif pre in fan_of:                        # conv / linear weight or bias
            b = 1.0 / np.sqrt(fan_of[pre])
            out[k] = torch.from_numpy(r.uniform(-b, b, shp).astype(np.float32))
        elif k.endswith("weight"):                 # BN gamma
            out[k] = torch.from_numpy(r.uniform(0.5, 1.5, shp).astype(np.float32))
        else:                                      # BN beta
            out[k] = torch.from_numpy(r.uniform(-0.2, 0.2, shp).astype(np.float32))
    if "first_conv.weight" in out and "inc.0.weight" in out:   # CubeNET registers one Conv3d twice
        out["inc.0.weight"], out["inc.0.bias"] = out["first_conv.weight"], out["first_conv.bias"]
    return out


def _dc_schema(pre: str, cin: int, cout: int, mid: int | None = None):
    mid = mid or cout
    s = {}
    for i, (a, b) in ((0, (cin, mid)), (3, (mid, cout))):
        s[f"{pre}.{i}.weight"] = (b, a, 3, 3)
        s[f"{pre}.{i}.bias"] = (b,)
        for nm in ("weight", "bias", "running_mean", "running_var"):
            s[f"{pre}.{i + 1}.{nm}"] = (b,)
        s[f"{pre}.{i + 1}.num_batches_tracked"] = ()
    return s


def unet_schema(n_channels: int, n_classes: int = 1, first: str = "unet", hsi_depth: int = 0, attention: bool = False,
                first_depth: int = 64, bilinear: bool = False):
    """State-dict schema of UNet / CubeNET-64, bilinear=False.  attention=True:
    Up's DoubleConv takes the product skip*up, i.e. Cin/2 input channels (model_parts.py:65-66).  first_depth != 64
    (CubeNET only): first_conv / inc2 have first_depth maps and the last decoder block is upsample4 / upconv4 over
    cat([x1, up]) whatever `attention` says (models.py:193-199, 229-240)."""
    fd = first_depth if first != "unet" else 64
    if bilinear:
        assert fd == 64, "the reference's bilinear + first_depth != 64 branch does not run (models.py:195-196)"
    s: Dict[str, Tuple[int, ...]] = {}
    if first == "unet":
        s.update(_dc_schema("inc.double_conv", n_channels, 64))
    else:  # CubeNET: Conv3d registered twice (models.py:169-171) -> aliased keys
        for pre in ("first_conv", "inc.0"):
            s[f"{pre}.weight"] = (fd, 1, hsi_depth, 3, 3)
            s[f"{pre}.bias"] = (fd,)
        for blk, c in (("inc.1", fd), ("inc2.1", fd)):
            for nm in ("weight", "bias", "running_mean", "running_var"):
                s[f"{blk}.{nm}"] = (c,)
            s[f"{blk}.num_batches_tracked"] = ()
        s["inc2.0.weight"] = (fd, fd, 3, 3)
        s["inc2.0.bias"] = (fd,)
    chans = [fd, 128, 256, 512, 1024]
    for i in range(1, 5):
        s.update(_dc_schema(f"down{i}.maxpool_conv.1.double_conv", chans[i - 1], chans[i] // (2 if bilinear and i == 4 else 1)))
    for i in range(1, 5):
        cin = chans[5 - i]
        if bilinear:                      # Up(in, out, bilinear): nn.Upsample has no parameters (model_parts.py:56-61)
            out = 128 if i == 4 else cin // 2          # models.py:46-49: up4 = Up(128, 64 * factor)
            s.update(_dc_schema(f"up{i}.conv.double_conv", cin // 2 if attention else cin, out // 2, cin // 2))
            continue
        if i == 4 and fd != 64:
            s["upsample4.weight"] = (cin, 64, 2, 2)
            s["upsample4.bias"] = (64,)
            s.update(_dc_schema("upconv4.double_conv", 64 + fd, 64))
            continue
        s[f"up{i}.up.weight"] = (cin, cin // 2, 2, 2)
        s[f"up{i}.up.bias"] = (cin // 2,)
        s.update(_dc_schema(f"up{i}.conv.double_conv", cin // 2 if attention else cin, cin // 2))
    s["outc.conv.weight"] = (n_classes, 64, 1, 1)
    s["outc.conv.bias"] = (n_classes,)
    return s


def spectral_schema(hsi_depth: int, n_classes: int = 1, bn_feats: int = 16, bnorm: bool = True):
    """bnorm=False: every block is Linear -> ReLU (models.py:105-110), no BatchNorm1d entries."""
    s: Dict[str, Tuple[int, ...]] = {}
    f = bn_feats
    dims = {"tail": (hsi_depth, f), "down1": (f, f), "down2": (f, f), "down3": (f, f), "down4": (f, f),
            "up1": (f, f), "up2": (2 * f, f), "up3": (2 * f, f), "up4": (2 * f, f)}
    for nm, (i, o) in dims.items():
        s[f"{nm}.0.weight"] = (o, i)
        s[f"{nm}.0.bias"] = (o,)
        if not bnorm:
            continue
        for q in ("weight", "bias", "running_mean", "running_var"):
            s[f"{nm}.1.{q}"] = (o,)
        s[f"{nm}.1.num_batches_tracked"] = ()
    s["outc.weight"] = (n_classes, 2 * f)
    s["outc.bias"] = (n_classes,)
    return s


def synth_cube(seed: int, n: int, bands: int, h: int, w: int) -> torch.Tensor:
    """Synthetic reflectance-like cube N x bands x H x W in [0,1)."""
    r = np.random.RandomState(1000 + seed)
    return torch.from_numpy(r.random_sample((n, bands, h, w)).astype(np.float32))


def synth_mask(seed: int, n: int, h: int, w: int, p_root: float = 0.05) -> torch.Tensor:
    r = np.random.RandomState(2000 + seed)
    return torch.from_numpy((r.random_sample((n, 1, h, w)) > 1.0 - p_root).astype(np.float32))


# --------------------------------------------------------------------------------------
# ingest (src/dataset.py:261-271, 283-295)
# --------------------------------------------------------------------------------------
def ingest_hsi(cube_hwb: np.ndarray, hsi_lo: int, hsi_hi: int, unsqueeze: bool) -> np.ndarray:
    """ENVI cubes arrive H x W x bands; the reference moves bands first, slices
    [hsi_lo:hsi_hi] and optionally adds a leading axis for CubeNET (dataset.py:266-270)."""
    img = np.moveaxis(np.asarray(cube_hwb), -1, 0)[hsi_lo:hsi_hi]
    if unsqueeze:
        img = img[None]
    return np.ascontiguousarray(img)


def crop_and_rescale(img: np.ndarray, i: int, j: int, th: int, tw: int) -> np.ndarray:
    """RandomCrop at a drawn (i, j) followed by the '/255 if max > 10' rule that applies
    only when a transform exists (dataset.py:284-289)."""
    out = img[..., i:i + th, j:j + tw]
    if out.max() > 10:
        out = out / 255
    return np.ascontiguousarray(out)


def binarise_mask(label_u8: np.ndarray) -> np.ndarray:
    """ToTensor()*255 then >0 -> 1 (dataset.py:294-295); result float32 {0,1}."""
    lab = label_u8.astype(np.float32) / 255.0 * 255.0
    return np.where(lab > 0, np.ones_like(lab), np.zeros_like(lab)).astype(np.float32)


def band_normalise(x: torch.Tensor, mean_b: torch.Tensor, std_b: torch.Tensor) -> torch.Tensor:
    """Optional per-band normalisation named by north_star; NOT in the reference
     -- restated as (x - mu_b) / sigma_b over the band axis -3."""
    return (x - mean_b[:, None, None]) / std_b[:, None, None]


# --------------------------------------------------------------------------------------
# storage-precision emulation
# --------------------------------------------------------------------------------------
class _RoundBF16(torch.autograd.Function):
    """Round to fp16 in forward AND round the incoming gradient to bf16 in backward: models a
    tensor the B200 path stores in fp16 between kernels and whose gradient it stores in bf16."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.float16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        if _GRAD_SCALE:        # loss-scaled fp16 gradient storage
            return (g * _GRAD_SCALE).to(torch.float16).to(torch.float32) / _GRAD_SCALE
        return g.to(torch.bfloat16).to(torch.float32)


_QUANT = False
_GRAD_SCALE = 0.0


def emulate_bf16_storage(on: bool, grad_scale: float = 0.0):
    """With emulation on, the oracle rounds weights, conv/linear outputs and activations to fp16 and
    their gradients to bf16 at exactly the points where hyperpri_b200 stores them (fp32 accumulation
    and fp32 BatchNorm statistics are kept).  The fp32 oracle (off, default) is the parity
    reference; the emulated oracle separates implementation errors from storage-precision
    effects in tests."""
    global _QUANT, _GRAD_SCALE
    _QUANT = bool(on)
    _GRAD_SCALE = float(grad_scale)   # 0: gradients rounded to bf16; S: gradients stored as fp16(S * g)


def q(x):
    return _RoundBF16.apply(x) if _QUANT else x


# --------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------
def batch_norm(x, sd, pre, training, stats_out=None, ch_axis=1):
    """BatchNormNd restated by hand: biased variance to normalise, unbiased for the running
    update, eps 1e-5, momentum 0.1."""
    red = [d for d in range(x.dim()) if d != ch_axis]
    shp = [1] * x.dim()
    shp[ch_axis] = -1
    if training:
        mean = x.mean(dim=red)
        var = x.var(dim=red, unbiased=False)
        if stats_out is not None:
            m = x.numel() // x.shape[ch_axis]
            stats_out[pre + ".running_mean"] = ((1 - BN_MOMENTUM) * sd[pre + ".running_mean"]
                                               + BN_MOMENTUM * mean.detach())
            stats_out[pre + ".running_var"] = ((1 - BN_MOMENTUM) * sd[pre + ".running_var"]
                                              + BN_MOMENTUM * var.detach() * m / max(m - 1, 1))
            stats_out[pre + ".num_batches_tracked"] = stats_out.get(
                pre + ".num_batches_tracked", sd[pre + ".num_batches_tracked"]) + 1
    else:
        mean, var = sd[pre + ".running_mean"], sd[pre + ".running_var"]
    xh = (x - mean.view(shp)) * torch.rsqrt(var.view(shp) + BN_EPS)
    return xh * sd[pre + ".weight"].view(shp) + sd[pre + ".bias"].view(shp)


def _cbr(x, w, b, sd, bn, training, stats_out, quant_out=True):
    """conv3x3 pad 1 -> BatchNorm -> ReLU.  (Emulation: the bias-free conv output is what is stored.)"""
    x = q(F.conv2d(x, q(w), None, padding=1)) + b.view(1, -1, 1, 1)
    x = torch.relu(batch_norm(x, sd, bn, training, stats_out))
    return q(x) if quant_out else x


def double_conv(x, sd, pre, training, stats_out=None, quant_out=True):
    """(conv3x3 pad 1 -> BN -> ReLU) x 2  (model_parts.py:14-31).  With quant_out=False the last
    activation stays fp32 (the B200 path never stores the tensor the 1x1 head or a pool+skip pair
    consumes; see emulate_bf16_storage)."""
    x = _cbr(x, sd[f"{pre}.0.weight"], sd[f"{pre}.0.bias"], sd, f"{pre}.1", training, stats_out)
    return _cbr(x, sd[f"{pre}.3.weight"], sd[f"{pre}.3.bias"], sd, f"{pre}.4", training, stats_out, quant_out)


def down(x, sd, pre, training, stats_out=None, quant_out=True):
    """MaxPool2d(2) (floor) -> DoubleConv  (model_parts.py:34-45).  x is the fp32 activation of the
    level above; the pooled copy and the skip copy are separately stored tensors."""
    return double_conv(q(F.max_pool2d(x, 2)), sd, pre + ".maxpool_conv.1.double_conv", training, stats_out,
                       quant_out)


def up(x1, x2, sd, pre, training, stats_out=None, quant_out=True, attention=False, up_key=None, conv_key=None):
    """ConvTranspose2d(k2,s2) -> zero pad to the skip's size (left/top floor(d/2), rest
    right/bottom) -> cat([skip, up]) (use_attention: skip * up, model_parts.py:84-85) -> DoubleConv
    (model_parts.py:71-90)."""
    up_key, conv_key = up_key or pre + ".up", conv_key or pre + ".conv.double_conv"
    if up_key + ".weight" in sd:
        x1 = q(F.conv_transpose2d(x1, q(sd[up_key + ".weight"]), sd[up_key + ".bias"], stride=2))
    else:                                 # bilinear=True: nn.Upsample(scale_factor=2, bilinear, align_corners=True)
        x1 = q(F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True))
    dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    x = q(q(x2) * x1) if attention else torch.cat([q(x2), x1], dim=1)
    return double_conv(x, sd, conv_key, training, stats_out, quant_out)


def _unet_body(x1, sd, training, stats_out, attention=False):
    """x1 and the other encoder outputs are kept fp32 here: their stored copies are the pooled
    tensor (down) and the skip half of the concat buffer (up)."""
    x2 = down(x1, sd, "down1", training, stats_out, quant_out=False)
    x3 = down(x2, sd, "down2", training, stats_out, quant_out=False)
    x4 = down(x3, sd, "down3", training, stats_out, quant_out=False)
    x5 = down(x4, sd, "down4", training, stats_out)
    x = up(x5, x4, sd, "up1", training, stats_out, attention=attention)
    x = up(x, x3, sd, "up2", training, stats_out, attention=attention)
    x = up(x, x2, sd, "up3", training, stats_out, attention=attention)
    if "upsample4.weight" in sd:              # CubeNET first_depth != 64 (models.py:229-240): always concatenated
        x = up(x, x1, sd, "up4", training, stats_out, quant_out=False, up_key="upsample4", conv_key="upconv4.double_conv")
    else:
        x = up(x, x1, sd, "up4", training, stats_out, quant_out=False, attention=attention)
    return F.conv2d(x, sd["outc.conv.weight"], sd["outc.conv.bias"])      # model_parts.py:93-99


def unet_forward(x, sd, training=True, stats_out=None, attention=False):
    """UNet.forward (models.py:53-68), bilinear=False."""
    return _unet_body(double_conv(q(x), sd, "inc.double_conv", training, stats_out, quant_out=False), sd, training,
                      stats_out, attention)


def cubenet_forward(x, sd, training=True, stats_out=None, attention=False):
    """CubeNET.forward (models.py:202-247), any first_depth.  The Conv3d with a kernel
    spanning every band and pad (0,1,1) is restated as the 2-D conv it equals
    (SURVEY.md appendix B.3): x is N x 1 x D x R x C."""
    w = sd["first_conv.weight"]
    x1 = _cbr(q(x[:, 0]), w[:, 0], sd["first_conv.bias"], sd, "inc.1", training, stats_out)
    x1 = _cbr(x1, sd["inc2.0.weight"], sd["inc2.0.bias"], sd, "inc2.1", training, stats_out, quant_out=False)
    return _unet_body(x1, sd, training, stats_out, attention)


def spectralunet_forward(x, sd, training=True, stats_out=None):
    """SpectralUNET.forward (models.py:117-145): per-pixel MLP U-Net, one image at a time so
    every BatchNorm1d sees that image's pixels only (models.py:132)."""
    n, d, r, c = x.shape
    rast = x.reshape(n, d, r * c).permute(0, 2, 1)
    outs = []
    cur = dict(sd)

    def blk(t, nm):
        so = {} if stats_out is not None else None
        t = q(F.linear(t, q(sd[nm + ".0.weight"]), None)) + sd[nm + ".0.bias"]
        if nm + ".1.weight" not in sd:            # bnorm=False: Linear -> ReLU (models.py:105-110)
            return q(torch.relu(t))
        t = q(torch.relu(batch_norm(t, cur, nm + ".1", training, so, ch_axis=1)))
        if so:
            cur.update(so)        # running stats advance once per image
        return t

    for i in range(n):
        x0 = blk(q(rast[i]), "tail")
        x1 = blk(x0, "down1")
        x2 = blk(x1, "down2")
        x3 = blk(x2, "down3")
        x4 = blk(x3, "down4")
        t = blk(x4, "up1")
        t = blk(torch.cat((x3, t), -1), "up2")
        t = blk(torch.cat((x2, t), -1), "up3")
        t = blk(torch.cat((x1, t), -1), "up4")
        t = F.linear(torch.cat((x0, t), -1), sd["outc.weight"], sd["outc.bias"])
        outs.append(t.reshape(-1, r, c))          # plain reshape, as models.py:144 does
    if stats_out is not None:
        for k in cur:
            if "running_" in k or "num_batches" in k:
                stats_out[k] = cur[k]
    return torch.stack(outs, 0)


def spectralunet_forward_streaming(x, sd, chunk: int = 16384):
    """Train-mode SpectralUNET forward (models.py:117-145) for full-width patches without an autograd graph:
    every block is evaluated over row chunks of the pixel matrix, the per-image BatchNorm1d statistics are
    accumulated chunk-wise in fp64 (two passes: statistics, then normalise in place), so the peak memory is the
    five skip tensors plus one pre-activation buffer (~17 GB fp32 per 608 x 700 image at 1650 features) instead of the
    146 GB the autograd version needs for batch 2.  Returns logits N x 1 x R x C; running statistics are not
    advanced (forward check only).  tests/test_oracle_golden.py holds it to spectralunet_forward on small cases."""
    n, d, r, c = x.shape
    m = r * c
    outs = []
    with torch.no_grad():
        def blk(parts, nm):
            """parts: list of [m, f_i] tensors whose concatenation (models.py:140-143) is the block input."""
            w, b = sd[nm + ".0.weight"], sd[nm + ".0.bias"]
            ws, o = [], 0
            for t in parts:
                ws.append(w[:, o:o + t.shape[1]])
                o += t.shape[1]
            pre = torch.empty((m, w.shape[0]), dtype=torch.float32)
            s1 = torch.zeros(w.shape[0], dtype=torch.float64)
            s2 = torch.zeros(w.shape[0], dtype=torch.float64)
            for i in range(0, m, chunk):
                y = b + sum(t[i:i + chunk] @ wi.t() for t, wi in zip(parts, ws))
                pre[i:i + chunk] = y
                yd = y.double()
                s1 += yd.sum(0)
                s2 += (yd * yd).sum(0)
            if nm + ".1.weight" not in sd:        # bnorm=False
                scale, shift = torch.ones(w.shape[0]), torch.zeros(w.shape[0])
            else:
                mean = s1 / m
                var = (s2 / m - mean * mean).clamp_min(0)
                scale = (sd[nm + ".1.weight"].double() / torch.sqrt(var + BN_EPS)).float()
                shift = (sd[nm + ".1.bias"].double() - mean * scale.double()).float()
            for i in range(0, m, chunk):
                pre[i:i + chunk] = torch.relu(pre[i:i + chunk] * scale + shift)
            return pre

        rast = x.reshape(n, d, m).permute(0, 2, 1)
        for i in range(n):
            x0 = blk([rast[i].contiguous()], "tail")
            x1 = blk([x0], "down1")
            x2 = blk([x1], "down2")
            x3 = blk([x2], "down3")
            x4 = blk([x3], "down4")
            t = blk([x4], "up1")
            del x4
            t = blk([x3, t], "up2")
            del x3
            t = blk([x2, t], "up3")
            del x2
            t = blk([x1, t], "up4")
            del x1
            wo = sd["outc.weight"]
            f = x0.shape[1]
            lg = x0 @ wo[:, :f].t() + t @ wo[:, f:].t() + sd["outc.bias"]
            outs.append(lg.reshape(-1, r, c))
    return torch.stack(outs, 0)


def bce_with_logits(pred, target):
    """BCEWithLogitsLoss(), mean reduction (params_HyperPRI.py:60,223; PLTrainer.py:86):
    mean(max(x,0) - x t + log1p(exp(-|x|)))."""
    return (pred.clamp_min(0) - pred * target + torch.log1p(torch.exp(-pred.abs()))).mean()


def seg_counts(pred, mask, thr=0.5):
    """TP/FP/FN/TN of sigmoid(pred) > thr against the int mask (PLTrainer.py:80,88-91)."""
    seg = torch.sigmoid(pred) > thr
    m = mask > 0.5
    tp = (seg & m).sum().item(); fp = (seg & ~m).sum().item()
    fn = (~seg & m).sum().item(); tn = (~seg & ~m).sum().item()
    return tp, fp, fn, tn


FORWARDS = {"UNET": unet_forward, "CubeNET": cubenet_forward, "SpectralUNET": spectralunet_forward}


def forward_backward(model: str, x: torch.Tensor, mask: torch.Tensor, sd: Dict[str, torch.Tensor],
                     training: bool = True, attention: bool = False):
    """One training-step body (PLTrainer.py:79-98 without metrics): logits, loss, grads per
    state-dict key, updated BN buffers."""
    leaf = {}
    for k, v in sd.items():
        if v.dtype.is_floating_point and "running_" not in k:
            leaf[k] = v.detach().clone().requires_grad_(True)
        else:
            leaf[k] = v
    if model == "CubeNET":                      # aliased module: one tensor, two keys
        leaf["inc.0.weight"] = leaf["first_conv.weight"]
        leaf["inc.0.bias"] = leaf["first_conv.bias"]
    stats: Dict[str, torch.Tensor] = {}
    kw = {"attention": True} if attention else {}
    logits = FORWARDS[model](x, leaf, training, stats if training else None, **kw)
    loss = bce_with_logits(logits, mask)
    loss.backward()
    grads = {k: v.grad for k, v in leaf.items() if isinstance(v, torch.Tensor) and v.requires_grad
             and v.grad is not None and not k.startswith("inc.0.")}
    return logits.detach(), loss.detach(), grads, stats
