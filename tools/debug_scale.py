"""Debug: engine gradients at grad_scale 1 vs 0.5 (x2) on one GPU, and against torch autograd fp32."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperpri_b200.src.Experiments.models import CubeNET
dev = torch.device("cuda", 0)
torch.manual_seed(7)
net = CubeNET(24, 1, first_depth=64, bilinear=False).to(dev).train()
g = torch.Generator().manual_seed(100)
x = torch.rand((2, 1, 24, 64, 80), generator=g).to(dev)
m = (torch.rand((2, 1, 64, 80), generator=g) > 0.7).float().to(dev)
eng = net._get_engine(dev)
LOG = []
def grads(scale):
    eng.invalidate_packed()
    logits = eng.forward(x, True)
    LOG.append(logits.detach().clone())
    _, dlogit, _ = eng.loss_and_dlogit(logits, m, grad_scale=scale)
    gr = eng.backward(dlogit, prescaled=True)
    torch.cuda.synchronize()
    return {k: v.detach().float().clone() for k, v in gr.items()}
a = grads(1.0); b = grads(0.5); c = grads(1.0)
print("logits a-b", float((LOG[0]-LOG[1]).abs().max()), "a-c", float((LOG[0]-LOG[2]).abs().max()), "max", float(LOG[0].abs().max()))
worst = 0.0
for k in a:
    den = float(a[k].abs().max()) + 1e-30
    e1 = float((a[k] - 2 * b[k]).abs().max()) / den
    e2 = float((a[k] - c[k]).abs().max()) / den
    worst = max(worst, e1, e2)
    if False:
        print(f"{k:40s} |a-2b|/max {e1:.2e}   |a-c|/max {e2:.2e}  max {den:.3e}")
print("worst rel", worst)
