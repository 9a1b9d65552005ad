"""One CubeNET-64 training step (batch 2, 238x608x968) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...`.  Two warm-up steps run outside the profiled range."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperpri_b200.src.Experiments.models import CubeNET, SpectralUNET, UNet   # noqa: E402

H, W, BANDS = 608, 968, 238


def misc(dev):
    """The kernels no training step launches: multi-tensor Adam, the validation histogram pass, the bilinear x2
    upsample pair and the attention product, at BASELINE sizes."""
    from hyperpri_b200 import metrics as M, ops
    from hyperpri_b200.optim import FusedAdam
    net = CubeNET(BANDS, 1, first_depth=64, bilinear=False).to(dev)
    for p in net.parameters():
        p.grad = torch.randn_like(p) * 1e-3
    opt = FusedAdam(net.parameters(), lr=1e-3)
    logits = torch.randn((2, 1, H, W), device=dev) * 3
    mask = (torch.rand((2, 1, H, W), device=dev) > 0.95).float()
    curve = M.DevicePRCurve(dev, 500)
    lo = torch.randn((2, H // 2, W // 2, 128), device=dev).half()
    cat = torch.zeros((2, H, W, 256), device=dev, dtype=torch.float16)
    dlo = torch.empty_like(lo)
    prod = torch.empty((2, H, W, 128), device=dev, dtype=torch.float16)

    def run():
        opt.step()
        curve.update(logits, mask)
        ops.upsample2_fwd(lo, cat[..., 128:])
        ops.upsample2_bwd(cat[..., 128:], dlo)
        ops.mul16(cat[..., :128], cat[..., 128:], prod)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("ok")


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "CubeNET"
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if model == "misc":
        return misc(dev)
    if model == "SpectralUNET":
        net = SpectralUNET(BANDS, 1, bn_feats=1650).to(dev).train()
        x = torch.rand((1, BANDS, H, 700), device=dev)
        mask = (torch.rand((1, 1, H, 700), device=dev) > 0.95).float()
        eng = net._get_engine(dev)

        def sstep():
            eng.invalidate_packed()
            logits = eng.forward(x, True)
            _, dlogit, _ = eng.loss_and_dlogit(logits, mask)
            eng.backward(dlogit, prescaled=True)
        for _ in range(2):
            sstep()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        sstep()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print("ok")
        return
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
