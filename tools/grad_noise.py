"""Run-to-run spread of the gradients of one training step on identical inputs (deterministic forward statistics), and
which parameters carry it.  Used to tell summation-order noise (present without any concurrency) from a race between
the main and the weight-gradient stream (would vanish with HPRI_WGRAD_STREAM=0 / CUDA_LAUNCH_BLOCKING=1).

    python tools/grad_noise.py [CubeNET|UNET] [h] [w]
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
runs = []
for it in range(5):
    net.train()
    net.zero_grad(set_to_none=True)
    logits = net(xin)
    loss = torch.nn.BCEWithLogitsLoss()(logits, mask)
    loss.backward()
    torch.cuda.synchronize()
    runs.append({k: p.grad.detach().clone() for k, p in net.named_parameters()})
ref = runs[0]
flat = lambda d: torch.cat([v.flatten() for v in d.values()])
tag = f"{model} {h}x{w} WGRAD_STREAM={os.environ.get('HPRI_WGRAD_STREAM', '1')} BLOCKING={os.environ.get('CUDA_LAUNCH_BLOCKING', '0')}"
for i in range(1, 5):
    print(tag, f"run {i} vs 0: rel L2 {((flat(runs[i]) - flat(ref)).norm() / flat(ref).norm()).item():.3e}")
for i in range(1, 5):
    for j in range(i + 1, 5):
        print(tag, f"run {j} vs {i}: rel L2 {((flat(runs[j]) - flat(runs[i])).norm() / flat(runs[i]).norm()).item():.3e}")
per = sorted(((((runs[1][k] - ref[k]).norm() / (ref[k].norm() + 1e-30)).item(), k) for k in ref), reverse=True)
print("largest per-parameter differences:", [(f"{v:.2e}", k) for v, k in per[:6]])
print("smallest:", [(f"{v:.2e}", k) for v, k in per[-4:]])
