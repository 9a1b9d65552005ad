"""Interleaved A/B timing of one CubeNET-64 training step (batch 2, 238 x 608 x 968) under engine switches: the
configurations alternate inside ONE process so that clock / power-cap drift hits all of them alike.
    python tools/ab_step.py [rounds] [steps]   -> one JSON line per configuration (median / min ms per step)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperpri_b200 import ops                                          # noqa: E402
from hyperpri_b200.src.Experiments.models import CubeNET              # noqa: E402

H, W, BANDS = 608, 968, 238


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    which = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = CubeNET(BANDS, 1, first_depth=64, bilinear=False).to(dev).train()
    x = torch.rand((2, 1, BANDS, H, W), device=dev)
    mask = (torch.rand((2, 1, H, W), device=dev) > 0.95).float()
    eng = net._get_engine(dev)
    state = {"prefetch": True}

    def step():
        eng.invalidate_packed()
        logits = eng.forward(x, True)
        _, dlogit, _ = eng.loss_and_dlogit(logits, mask)
        if state["prefetch"]:
            eng.set_next_input(x)
        eng.backward(dlogit, prescaled=True)

    def cfg(prefetch=True, overlap=True, conv=-1, wgrad=-1, reverse=False, a_stages=2):
        def apply():
            ops.set_halo_a_stages(a_stages)
            ops.set_reverse_elementwise(reverse)
            state["prefetch"] = prefetch
            eng.set_overlap(overlap)
            eng._no_prefetch = not prefetch
            ops.set_conv_algo(conv)
            ops.set_wgrad_algo(wgrad)
        return apply
    configs = {"default": cfg(), "no_prefetch": cfg(prefetch=False), "no_overlap": cfg(prefetch=False, overlap=False),
               "single_cta_convs": cfg(conv=1), "generic_wgrad": cfg(wgrad=0), "reversed_elementwise": cfg(reverse=True), "halo_a_stages3": cfg(a_stages=3)}
    if which:
        configs = {k: v for k, v in configs.items() if k in which}
    res = {k: [] for k in configs}
    for r in range(rounds):
        for name, apply in configs.items():
            apply()
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            res[name].append(e0.elapsed_time(e1) / steps)
    # ---- timeline of one default step: start / end of every native call relative to the step's first launch, with the
    # stream it ran on (events are recorded on the launching stream, so overlap between the streams is visible)
    out = os.environ.get("HPRI_TIMELINE")
    if out:
        configs_default = cfg()
        configs_default()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ops.PROFILE = []
        t0 = torch.cuda.Event(enable_timing=True)
        t0.record()
        step()
        torch.cuda.synchronize()
        rec, ops.PROFILE = ops.PROFILE, None
        rows = [{"kernel": n, "start_us": t0.elapsed_time(a) * 1e3, "end_us": t0.elapsed_time(b) * 1e3, "args": d[:60]} for n, a, b, d in rec]
        with open(out, "w") as f:
            json.dump(rows, f, indent=0)
    for name, v in res.items():
        v = sorted(v)
        print(json.dumps({"config": name, "median_ms": v[len(v) // 2], "min_ms": v[0], "max_ms": v[-1],
                          "images_per_s_median": 2e3 / v[len(v) // 2], "rounds": rounds, "steps": steps}))


if __name__ == "__main__":
    main()
