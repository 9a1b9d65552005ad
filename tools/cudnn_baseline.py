"""SURVEY.md 8(d) "kernel to beat": the same network as plain torch.nn modules on the same B200 under cuDNN --
(i) fp32 parameters with TF32 matmuls (what PLTrainer.py:32-34 sets), (ii) bf16 autocast + channels_last.
Stand-alone restatement of the architecture (model_parts.py:14-99, models.py:148-247 with the Conv3d written as the
equivalent Conv2d over 238 bands); imports neither the product nor oracle/.  forward -> BCEWithLogits -> backward,
batch 2, 238x608x968 (CubeNET-64) or 3x608x968 (UNET), CUDA-event timed.  Prints one JSON line per mode."""
import json
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F


def dc(ci, co):
    return nn.Sequential(nn.Conv2d(ci, co, 3, padding=1), nn.BatchNorm2d(co), nn.ReLU(inplace=True),
                         nn.Conv2d(co, co, 3, padding=1), nn.BatchNorm2d(co), nn.ReLU(inplace=True))


class Net(nn.Module):
    def __init__(self, cin, cube):
        super().__init__()
        if cube:
            self.inc = nn.Sequential(nn.Conv2d(cin, 64, 3, padding=1), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
                                     nn.Conv2d(64, 64, 3, padding=1), nn.BatchNorm2d(64), nn.ReLU(inplace=True))
        else:
            self.inc = dc(cin, 64)
        C = [64, 128, 256, 512, 1024]
        self.down = nn.ModuleList([dc(C[i], C[i + 1]) for i in range(4)])
        self.upT = nn.ModuleList([nn.ConvTranspose2d(C[i + 1], C[i], 2, stride=2) for i in (3, 2, 1, 0)])
        self.upc = nn.ModuleList([dc(2 * C[i], C[i]) for i in (3, 2, 1, 0)])
        self.outc = nn.Conv2d(64, 1, 1)

    def forward(self, x):
        xs = [self.inc(x)]
        for d in self.down:
            xs.append(d(F.max_pool2d(xs[-1], 2)))
        y = xs[4]
        for k, (t, c) in enumerate(zip(self.upT, self.upc)):
            skip = xs[3 - k]
            y = t(y)
            y = F.pad(y, [0, skip.shape[3] - y.shape[3], 0, skip.shape[2] - y.shape[2]])
            y = c(torch.cat([skip, y], dim=1))
        return self.outc(y)


def run(model, mode, steps=10, warmup=3):
    cube = model == "CubeNET"
    cin = 238 if cube else 3
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = Net(cin, cube).to(dev).train()
    x = torch.rand((2, cin, 608, 968), device=dev)
    mask = (torch.rand((2, 1, 608, 968), device=dev) > 0.95).float()
    crit = nn.BCEWithLogitsLoss()
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    if mode == "bf16_channels_last":
        net = net.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)

    def step():
        net.zero_grad(set_to_none=True)
        if mode == "bf16_channels_last":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = net(x)
            loss = crit(out.float(), mask)
        else:
            loss = crit(net(x), mask)
        loss.backward()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"tool": "cudnn_baseline", "model": model, "mode": mode, "batch": 2, "ms_per_step": ms,
                      "images_per_s": 2000.0 / ms, "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}), flush=True)


if __name__ == "__main__":
    models = sys.argv[1:] or ["CubeNET", "UNET"]
    for m in models:
        for mode in ("tf32", "bf16_channels_last"):
            run(m, mode)
            torch.cuda.empty_cache()
