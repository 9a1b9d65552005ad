"""Isolated timings of the tcgen05 implicit-GEMM kernel on CubeNET layer shapes (CUDA events)."""
import math, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperpri_b200 import ops

DEV = "cuda"
def t_ms(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def conv_case(n, h, w, cin, cout, taps=9, bn=0):
    x = torch.randn((n, h, w, cin), device=DEV).to(ops.ACT)
    wt = torch.randn((cout, cin, 3, 3), device=DEV) * (1 / math.sqrt(cin * 9))
    spec = ops.WeightSpec("conv3x3", cout, cin)
    wp = spec.pack_fwd(wt)
    if taps == 1:
        wp = wp[:, : ops.kpad(cin)].contiguous()
    y = torch.empty((n, h, w, cout), device=DEV, dtype=ops.ACT)
    stats = torch.zeros((cout, 2), dtype=torch.float64, device=DEV)
    gf = 2.0 * n * h * w * cout * cin * taps / 1e9
    res = {}
    res["stats"] = t_ms(lambda: ops.igemm_fwd(x, wp, cout, taps, y, cout, stats=stats, block_n=bn))
    res["nostats"] = t_ms(lambda: ops.igemm_fwd(x, wp, cout, taps, y, cout, block_n=bn))
    # wgrad
    xg = x.to(ops.GRAD); dy = torch.randn((n, h, w, cout), device=DEV).to(ops.GRAD)
    gw = spec.grad_buffer(DEV)
    if taps == 9 and os.environ.get("WGRAD", "0") == "1":
        res["wgrad"] = t_ms(lambda: ops.igemm_wgrad(xg, dy, 1, cout, gw, block_n=bn))
    return gf, res

CASES = [
    ("inc2 64->64 @608x968", 2, 608, 968, 64, 64, 9, 0),
    ("1x1 64->64 @608x968 (epilogue only)", 2, 608, 968, 64, 64, 1, 0),
    ("first 240->64 @608x968", 2, 608, 968, 240, 64, 9, 0),
    ("up4.c1 128->64 @608x968", 2, 608, 968, 128, 64, 9, 0),
    ("down1.c2 128->128 @304x484", 2, 304, 484, 128, 128, 9, 0),
    ("down1.c2 128->128 @304x484 bn64", 2, 304, 484, 128, 128, 9, 64),
    ("down2.c2 256->256 @152x242", 2, 152, 242, 256, 256, 9, 0),
    ("down2.c2 256->256 @152x242 bn128", 2, 152, 242, 256, 256, 9, 128),
    ("up1.c1 1024->512 @76x121", 2, 76, 121, 1024, 512, 9, 0),
    ("down4.c2 1024->1024 @38x60", 2, 38, 60, 1024, 1024, 9, 0),
    ("down4.c2 1024->1024 @38x60 bn128", 2, 38, 60, 1024, 1024, 9, 128),
]
if __name__ == "__main__":
    out = []
    algo = int(os.environ.get("ALGO", "-1"))
    ops.set_conv_algo(algo)
    print("conv algo", algo)
    for name, n, h, w, ci, co, taps, bn in CASES:
        gf, r = conv_case(n, h, w, ci, co, taps, bn)
        line = f"{name:44s} {gf:7.1f} GF | " + " | ".join(f"{k} {v*1e3:7.1f} us {gf/v:6.0f} TF/s" for k, v in r.items())
        print(line); out.append({"name": name, "gflop": gf, **{k: v for k, v in r.items()}})
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
