"""One CubeNET-64 training step (batch 2, 238x608x968) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...`.  Two warm-up steps run outside the profiled range."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperpri_b200.src.Experiments.models import CubeNET, UNet   # noqa: E402

H, W, BANDS = 608, 968, 238


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "CubeNET"
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if model == "CubeNET":
        net = CubeNET(BANDS, 1, first_depth=64, bilinear=False).to(dev).train()
        x = torch.rand((2, 1, BANDS, H, W), device=dev)
    else:
        net = UNet(3, 1, bilinear=False).to(dev).train()
        x = torch.rand((2, 3, H, W), device=dev)
    mask = (torch.rand((2, 1, H, W), device=dev) > 0.95).float()
    eng = net._get_engine(dev)

    def step():
        eng.invalidate_packed()
        logits = eng.forward(x, True)
        _, dlogit, _ = eng.loss_and_dlogit(logits, mask)
        eng.backward(dlogit, prescaled=True)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("ok")


if __name__ == "__main__":
    main()
