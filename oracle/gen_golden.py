"""Generate tests/golden/*.npz by running the REFERENCE modules (imported from
/root/reference, build container only) on the deterministic synthetic inputs/weights of
oracle/hyperpri_oracle.py.  The fixtures hold reference OUTPUTS only; inputs and weights are
regenerated from seeds by the tests.  Also asserts the oracle restatement agrees with the
reference before writing (so a stale oracle cannot be pinned silently).

    python oracle/gen_golden.py            # needs /root/reference
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
import hyperpri_oracle as O                                     # noqa: E402
from src.Experiments.models import UNet, CubeNET, SpectralUNET  # noqa: E402  (reference)

OUT = os.path.join(HERE, "..", "tests", "golden")
torch.manual_seed(0)
torch.set_num_threads(8)

CASES = {
    # name: (model, ctor, schema, input shape builder)
    "unet_2x3x32x40": dict(model="UNET", n=2, h=32, w=40, bands=3, seed=0),
    "cubenet_2x238x32x40": dict(model="CubeNET", n=2, h=32, w=40, bands=238, seed=1),
    "cubenet_1x238x48x72": dict(model="CubeNET", n=1, h=48, w=72, bands=238, seed=2),
    "spectral32_2x238x6x10": dict(model="SpectralUNET", n=2, h=6, w=10, bands=238, seed=3, feats=32),
    "spectral1650_2x238x4x5": dict(model="SpectralUNET", n=2, h=4, w=5, bands=238, seed=4, feats=1650),
    # bnorm=False (models.py:72,105-110): Linear -> ReLU blocks
    "spectral32_nobn_2x238x6x10": dict(model="SpectralUNET", n=2, h=6, w=10, bands=238, seed=11, feats=32, bnorm=False),
    # use_attention=True (model_parts.py:84-85): skip * up instead of cat([skip, up])
    "unet_att_2x3x32x40": dict(model="UNET", n=2, h=32, w=40, bands=3, seed=5, attention=True),
    "cubenet_att_2x238x34x42": dict(model="CubeNET", n=2, h=34, w=42, bands=238, seed=6, attention=True),
    # first_depth != 64 (models.py:193-199, 229-240): upsample4 / upconv4 over cat([x1 (first_depth), up (64)])
    "cubenet_fd32_2x238x32x40": dict(model="CubeNET", n=2, h=32, w=40, bands=238, seed=7, first_depth=32),
    "cubenet_fd128_att_1x238x34x42": dict(model="CubeNET", n=1, h=34, w=42, bands=238, seed=8, first_depth=128, attention=True),
    # bilinear=True (model_parts.py:56-61): nn.Upsample + mid-channel DoubleConvs, down4 / decoder outputs halved
    "unet_bil_2x3x34x42": dict(model="UNET", n=2, h=34, w=42, bands=3, seed=9, bilinear=True),
    "cubenet_bil_att_2x238x32x40": dict(model="CubeNET", n=2, h=32, w=40, bands=238, seed=10, bilinear=True, attention=True),
}


def build(case):
    m = case["model"]
    att = case.get("attention", False)
    bil = case.get("bilinear", False)
    if m == "UNET":
        net = UNet(case["bands"], 1, bilinear=bil, use_attention=att)
        schema = O.unet_schema(case["bands"], 1, "unet", attention=att, bilinear=bil)
    elif m == "CubeNET":
        fd = case.get("first_depth", 64)
        net = CubeNET(case["bands"], 1, first_depth=fd, bilinear=bil, use_attention=att)
        schema = O.unet_schema(1, 1, "cube", hsi_depth=case["bands"], attention=att, first_depth=fd, bilinear=bil)
    else:
        net = SpectralUNET(case["bands"], 1, bn_feats=case["feats"], bnorm=case.get("bnorm", True))
        schema = O.spectral_schema(case["bands"], 1, case["feats"], bnorm=case.get("bnorm", True))
    ref_sd = net.state_dict()
    assert {k: tuple(v.shape) for k, v in ref_sd.items()} == {k: tuple(s) for k, s in schema.items()}, \
        "schema drifted from the reference state_dict"
    sd = O.synth_state_dict(schema, case["seed"])
    if m == "CubeNET":
        sd["inc.0.weight"] = sd["first_conv.weight"]; sd["inc.0.bias"] = sd["first_conv.bias"]
    net.load_state_dict(sd)
    return net, sd


def run(name, case):
    net, sd = build(case)
    x = O.synth_cube(case["seed"], case["n"], case["bands"], case["h"], case["w"])
    if case["model"] == "CubeNET":
        x = x[:, None]
    mask = O.synth_mask(case["seed"], case["n"], case["h"], case["w"])
    out = {}
    for mode in ("train", "eval"):
        net.load_state_dict(sd)
        net.train(mode == "train")
        net.zero_grad()
        logits = net(x)
        loss = torch.nn.BCEWithLogitsLoss()(logits, mask)
        loss.backward()
        ologits, oloss, ograds, ostats = O.forward_backward(case["model"], x, mask, sd, training=(mode == "train"),
                                                            attention=case.get("attention", False))
        scale = logits.abs().max().item()
        err = (ologits - logits).abs().max().item() / scale
        assert err < 2e-4, (name, mode, err)
        assert abs(oloss.item() - loss.item()) < 1e-5
        out[f"{mode}.logits"] = logits.detach().numpy()
        out[f"{mode}.loss"] = np.float64(loss.item())
        gn = {}
        for k, p in net.named_parameters():
            g = p.grad
            gn[k] = float(g.norm())
            og = ograds[k]
            rel = float((og - g).norm() / (g.norm() + 1e-12))
            absd = float((og - g).abs().max())
            assert rel < 3e-2 or absd < 1e-7, (name, mode, k, rel, absd)   # tiny-batch BN is ill-conditioned in fp32
            if g.numel() <= 2048:
                out[f"{mode}.grad.{k}"] = g.numpy().copy()
        out[f"{mode}.gradnorm.keys"] = np.array(list(gn.keys()))
        out[f"{mode}.gradnorm.vals"] = np.array(list(gn.values()), dtype=np.float64)
        if mode == "train":
            for k, v in net.state_dict().items():
                if "running_" in k or "num_batches" in k:
                    ov = ostats[k]
                    assert torch.allclose(ov.float(), v.float(), rtol=1e-4, atol=1e-6), (name, k)
                    if v.numel() <= 2048:
                        out[f"train.buf.{k}"] = v.numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:3]})


def ingest_case():
    r = np.random.RandomState(7)
    cube = r.random_sample((12, 20, 299)).astype(np.float32)          # H x W x bands, ENVI order
    # dataset.py:266-270 restated inline with numpy exactly as the reference writes it
    img = np.moveaxis(np.array(cube), -1, 0)
    img = img[25:263, :, :]
    cube5 = np.expand_dims(img, 0)
    assert np.array_equal(O.ingest_hsi(cube, 25, 263, False), img)
    assert np.array_equal(O.ingest_hsi(cube, 25, 263, True), cube5)
    np.savez_compressed(os.path.join(OUT, "ingest_12x20x299.npz"),
                        sum_per_band=img.sum(axis=(1, 2)).astype(np.float64), first=img[:, 0, 0], last=img[:, -1, -1])
    print("wrote ingest")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]                      # optional: regenerate just the named cases
    for nm, cs in CASES.items():
        if not only or nm in only:
            run(nm, cs)
    if not only:
        ingest_case()
