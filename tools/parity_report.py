"""Parity of the B200 path against the CPU oracle on identical synthetic inputs and weights.

    python tools/parity_report.py --model CubeNET --n 2 --h 152 --w 242 [--json out.json]

Prints: logits max-abs error relative to max|logit| (north_star tolerance 1e-2), loss error,
agreement of thresholded masks (>= 99.9 %), per-parameter gradient relative L2 errors.
Runs the oracle on the host CPU cores of the box; needs the CUDA extension and a GPU.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hyperpri_oracle as O                                             # noqa: E402
from hyperpri_b200.src.Experiments.models import UNet, CubeNET, SpectralUNET   # noqa: E402


def build(model, bands, feats):
    if model == "UNET":
        return UNet(bands, 1, bilinear=False), O.unet_schema(bands, 1, "unet")
    if model == "CubeNET":
        return CubeNET(bands, 1, first_depth=64, bilinear=False), O.unet_schema(1, 1, "cube", hsi_depth=bands)
    return SpectralUNET(bands, 1, bn_feats=feats), O.spectral_schema(bands, 1, feats)


def run(model, n, h, w, bands, feats=1650, seed=0, training=True, verbose=True, emulate=False):
    import math
    O.emulate_bf16_storage(emulate, grad_scale=2.0 ** (math.ceil(math.log2(n * h * w)) - 4))
    net, schema = build(model, bands, feats)
    sd = O.synth_state_dict(schema, seed)
    net.load_state_dict(sd)
    net = net.cuda().train(training)
    x = O.synth_cube(seed, n, bands, h, w)
    xin = x[:, None] if model == "CubeNET" else x
    mask = O.synth_mask(seed, n, h, w)
    t0 = time.time()
    ologits, oloss, ograds, ostats = O.forward_backward(model, xin, mask, sd, training=training)
    t_cpu = time.time() - t0
    logits = net(xin.cuda())
    loss = torch.nn.BCEWithLogitsLoss()(logits, mask.cuda())
    if training:
        loss.backward()
    torch.cuda.synchronize()
    lg = logits.detach().cpu()
    scale = ologits.abs().max().item()
    res = {
        "model": model, "shape": [n, bands, h, w], "training": training, "oracle_cpu_s": round(t_cpu, 2),
        "oracle": "bf16-storage emulation" if emulate else "fp32",
        "logit_max_abs": scale,
        "logit_max_rel_err": (lg - ologits).abs().max().item() / scale,
        "logit_rms_rel_err": ((lg - ologits).pow(2).mean().sqrt() / ologits.pow(2).mean().sqrt()).item(),
        "loss": loss.item(), "loss_oracle": oloss.item(), "loss_abs_err": abs(loss.item() - oloss.item()),
        "mask_agreement": ((lg > 0) == (ologits > 0)).float().mean().item(),
    }
    if training:
        rel = {}
        for k, p in net.named_parameters():
            g, og = p.grad.detach().cpu(), ograds[k]
            rel[k] = ((g - og).norm() / (og.norm() + 1e-30)).item() if og.norm() > 1e-12 else float(g.norm())
        big = {k: v for k, v in rel.items() if not k.endswith(".bias") or "bn" in k or True}
        # conv/linear biases in front of a train-mode BN have an identically-zero gradient: report apart
        zero_bias = [k for k in rel if k.endswith(".bias") and ograds[k].norm() < 1e-6 * max(1e-30, max(
            ograds[q].norm().item() for q in ograds))]
        vals = sorted(v for k, v in big.items() if k not in zero_bias)
        res["grad_rel_l2_max"] = vals[-1]
        res["grad_rel_l2_median"] = vals[len(vals) // 2]
        res["grad_worst"] = sorted(((v, k) for k, v in big.items() if k not in zero_bias), reverse=True)[:3]
        res["zero_bias_grad_abs_max"] = max([net.get_parameter(k).grad.abs().max().item() for k in zero_bias] or [0])
        bufs = dict(net.named_buffers())
        errs = []
        for k, v in ostats.items():
            if "running_" in k and k in bufs:
                errs.append(((bufs[k].cpu() - v).abs().max() / (v.abs().max() + 1e-12)).item())
        res["running_stat_max_rel_err"] = max(errs) if errs else None
    O.emulate_bf16_storage(False)
    if verbose:
        print(json.dumps(res))
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="CubeNET")
    ap.add_argument("--n", type=int, default=2)
    ap.add_argument("--h", type=int, default=152)
    ap.add_argument("--w", type=int, default=242)
    ap.add_argument("--bands", type=int, default=None)
    ap.add_argument("--feats", type=int, default=1650)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--eval", action="store_true")
    ap.add_argument("--json", default=None)
    ap.add_argument("--emulate", action="store_true", help="oracle rounds stored tensors to bf16 like the GPU path")
    a = ap.parse_args()
    bands = a.bands or (3 if a.model == "UNET" else 238)
    torch.set_num_threads(os.cpu_count())
    r = run(a.model, a.n, a.h, a.w, bands, a.feats, a.seed, training=not a.eval, emulate=a.emulate)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(r, f, indent=1)
