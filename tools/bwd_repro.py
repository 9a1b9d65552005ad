"""Which buffers of the backward pass differ between two runs of the same step (deterministic forward)?  Prints, per
workspace tensor, the relative L2 difference and the fraction of elements that differ, to tell fp32 summation-order
noise (1e-7) from anything larger.

    python tools/bwd_repro.py [CubeNET|UNET] [h] [w]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import hyperpri_oracle as O                                            # noqa: E402
from hyperpri_b200 import ops                                          # noqa: E402
from test_models_gpu import build                                      # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "CubeNET"
h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (37, 51)
bands = 238 if model == "CubeNET" else 3
ops.set_deterministic(True, backward=os.environ.get("DET_BWD", "0") == "1")
net, _ = build(model, bands, seed=2)
x = O.synth_cube(3, 2, bands, h, w)
xin = (x[:, None] if model == "CubeNET" else x).cuda()
mask = O.synth_mask(3, 2, h, w).cuda()


def snap():
    net.train()
    net.zero_grad(set_to_none=True)
    logits = net(xin)
    loss = torch.nn.BCEWithLogitsLoss()(logits, mask)
    loss.backward()
    torch.cuda.synchronize()
    eng = net._get_engine(xin.device)
    out = {}
    for k, v in eng.ws.items():
        if torch.is_tensor(v) and v.is_floating_point():
            out["ws." + k] = v.detach().float().clone()
    for k, v in eng.grads.items():
        out["grad." + k] = v.detach().float().clone()
    if hasattr(eng, "_bw_sums") and torch.is_tensor(eng._bw_sums):
        out["bw_sums"] = eng._bw_sums.detach().double().clone()
    return out


snap()
a, b = snap(), snap()
seen = {}
for k in a:
    d = (a[k] - b[k]).double()
    n = a[k].double().norm().item()
    rel = d.norm().item() / (n + 1e-300)
    frac = (d != 0).double().mean().item()
    ptr = k
    print(f"{k:58s} rel L2 {rel:9.2e}   differing {100 * frac:7.3f} %   max|d|/max|a| {d.abs().max().item() / (a[k].abs().max().item() + 1e-300):9.2e}")
