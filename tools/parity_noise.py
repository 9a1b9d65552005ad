"""Run-to-run spread of the full-size parity figures in the default (atomic-order) statistics mode, and the value in the
deterministic mode: CubeNET-64 / UNET, 2 x (238 | 3) x 608 x 968, train-mode forward against the fp32 oracle.
    python tools/parity_noise.py [repeats]  -> JSON lines"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hyperpri_oracle as O                                                   # noqa: E402
from hyperpri_b200 import ops                                                 # noqa: E402
from hyperpri_b200.src.Experiments.models import CubeNET, UNet                # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    torch.set_num_threads(os.cpu_count())
    for model, bands in (("CubeNET", 238), ("UNET", 3)):
        n, h, w = 2, 608, 968
        if model == "UNET":
            net, schema = UNet(bands, 1, bilinear=False), O.unet_schema(bands, 1, "unet")
        else:
            net, schema = CubeNET(bands, 1, first_depth=64, bilinear=False), O.unet_schema(1, 1, "cube", hsi_depth=bands)
        sd = O.synth_state_dict(schema, 0)
        net.load_state_dict(sd)
        net = net.cuda().train()
        x = O.synth_cube(0, n, bands, h, w)
        xin = x[:, None] if model == "CubeNET" else x
        with torch.no_grad():
            ol = O.FORWARDS[model](xin, sd, True, None)
        xg = xin.cuda()
        for mode in ("default", "deterministic"):
            ops.set_deterministic(mode == "deterministic")
            agree, err, outs = [], [], []
            for _ in range(reps):
                with torch.no_grad():
                    lg = net(xg).cpu()
                outs.append(lg)
                agree.append(((lg > 0) == (ol > 0)).float().mean().item())
                err.append(((lg - ol).abs().max() / ol.abs().max()).item())
            spread = max((a - b).abs().max().item() for a in outs for b in outs) / ol.abs().max().item()
            print(json.dumps({"model": model, "mode": mode, "mask_agreement": agree, "logit_max_rel_err": err,
                              "max_run_to_run_diff_rel": spread, "flips_min_max": [int((1 - max(agree)) * ol.numel()),
                                                                                  int((1 - min(agree)) * ol.numel())]}))
        ops.set_deterministic(False)


if __name__ == "__main__":
    main()
