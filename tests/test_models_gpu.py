"""Model-level parity on a B200, called through the reference-facing nn.Module API (which calls the C ABI):
UNet / CubeNET / SpectralUNET forward + BCE + backward against the CPU oracle on identical synthetic cubes and
weights, and against the committed golden fixtures generated from the reference modules.

Tolerances (north_star): logits max abs error <= 1e-2 * max|logit|; thresholded masks agree >= 99.9 %
(fp16 activations, fp32 accumulation and BatchNorm statistics).  Gradients go through bf16 storage and, with
random-init weights on random cubes, are cancellation-dominated: the fp32 reference itself moves by ~15 % (median
relative L2 per parameter) under bf16/fp16-storage emulation, so they are checked by cosine similarity."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import hyperpri_oracle as O                                                    # noqa: E402
from hyperpri_b200.src.Experiments.models import UNet, CubeNET, SpectralUNET   # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(autouse=True)
def _deterministic_statistics():
    """The model-level tolerances below are asserted in the deterministic-statistics mode (fixed-order BatchNorm partial
    sums): the logits of a given build are then the same on every run, so a case that sits near its bound cannot pass on
    one run and fail on the next because an atomic landed in a different order.  The default mode's run-to-run spread is
    measured separately (tools/parity_noise.py; test_deterministic_statistics_mode_is_bit_reproducible compares the two
    modes; tests/test_zz_full_size_gpu.py also asserts the default mode at full size)."""
    from hyperpri_b200 import ops
    ops.set_deterministic(True, backward=False)
    yield
    ops.set_deterministic(False)


def build(model, bands, feats=1650, seed=0, att=False, fd=64, bil=False, bnorm=True):
    if model == "UNET":
        net = UNet(bands, 1, bilinear=bil, use_attention=att)
        schema = O.unet_schema(bands, 1, "unet", attention=att, bilinear=bil)
    elif model == "CubeNET":
        net = CubeNET(bands, 1, first_depth=fd, bilinear=bil, use_attention=att)
        schema = O.unet_schema(1, 1, "cube", hsi_depth=bands, attention=att, first_depth=fd, bilinear=bil)
    else:
        net, schema = SpectralUNET(bands, 1, bn_feats=feats, bnorm=bnorm), O.spectral_schema(bands, 1, feats, bnorm=bnorm)
    sd = O.synth_state_dict(schema, seed)
    net.load_state_dict(sd)
    return net.cuda(), sd


def run_ours(net, x, mask, train=True):
    net.train(train)
    net.zero_grad(set_to_none=True)
    logits = net(x.cuda())
    loss = torch.nn.BCEWithLogitsLoss()(logits, mask.cuda())
    if train:
        loss.backward()
    torch.cuda.synchronize()
    return logits.detach().cpu(), loss.item()


def cos(a, b):
    return (a.flatten().double() @ b.flatten().double() / (a.norm().double() * b.norm().double() + 1e-300)).item()


def record(**kw):
    """Measured parity figures -> gpurun_out/parity_records.jsonl (when run on the GPU box), for profiles/."""
    import json
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_records.jsonl"), "a") as f:
            f.write(json.dumps(kw) + "\n")


CASES = [("UNET", 2, 3, 96, 136, 1650), ("CubeNET", 2, 238, 96, 136, 1650), ("CubeNET", 1, 238, 80, 104, 1650),
         ("SpectralUNET", 2, 238, 24, 40, 1650), ("SpectralUNET", 1, 238, 16, 33, 96)]


@pytest.mark.parametrize("model,n,bands,h,w,feats", CASES)
def test_train_step_parity_vs_oracle(model, n, bands, h, w, feats):
    net, sd = build(model, bands, feats)
    x = O.synth_cube(0, n, bands, h, w)
    xin = x[:, None] if model == "CubeNET" else x
    mask = O.synth_mask(0, n, h, w)
    torch.set_num_threads(os.cpu_count())
    ol, oloss, og, ostats = O.forward_backward(model, xin, mask, sd, training=True)
    lg, loss = run_ours(net, xin, mask)
    assert lg.shape == ol.shape and lg.dtype == torch.float32
    record(test="train_step_parity_vs_oracle", model=model, shape=[n, bands, h, w], feats=feats,
           logit_max_rel_err=(lg - ol).abs().max().item() / ol.abs().max().item(),
           mask_agreement=((lg > 0) == (ol > 0)).float().mean().item())
    assert (lg - ol).abs().max().item() <= 1e-2 * ol.abs().max().item()
    # north_star's >= 99.9 % is asserted at BASELINE size in tests/test_zz_full_size_gpu.py.  At these small shapes the
    # bound is looser for a reason the oracle itself shows: rounding its own stored tensors to fp16 (emulate_bf16_storage)
    # flips 0.084 % of the 96 x 136 CubeNET masks against its fp32 self (22 of 26112 pixels), so 99.9 % would leave a
    # margin of four pixels; 960-pixel cases cannot even express 99.9 %.  At most 0.2 % (or 3 pixels) may differ here.
    flips = ((lg > 0) != (ol > 0)).sum().item()
    assert flips <= max(3, 2e-3 * lg.numel())
    assert abs(loss - oloss.item()) < 2e-4
    flat_o, flat_g = [], []
    for k, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        flat_o.append(og[k].flatten()); flat_g.append(p.grad.cpu().flatten())
    assert cos(torch.cat(flat_g), torch.cat(flat_o)) > 0.97
    bufs = dict(net.named_buffers())
    for k, v in ostats.items():
        if "running_" in k:
            assert torch.allclose(bufs[k].cpu(), v, rtol=5e-3, atol=5e-4), k
        elif "num_batches" in k:
            assert bufs[k].item() == int(v)


@pytest.mark.parametrize("name,model,n,bands,h,w,seed,feats", [
    ("unet_2x3x32x40", "UNET", 2, 3, 32, 40, 0, 0), ("cubenet_2x238x32x40", "CubeNET", 2, 238, 32, 40, 1, 0),
    ("cubenet_1x238x48x72", "CubeNET", 1, 238, 48, 72, 2, 0), ("spectral32_2x238x6x10", "SpectralUNET", 2, 238, 6, 10, 3, 32),
    ("spectral1650_2x238x4x5", "SpectralUNET", 2, 238, 4, 5, 4, 1650),
    ("spectral32_nobn_2x238x6x10", "SpectralUNET", 2, 238, 6, 10, 11, 32),
    ("unet_att_2x3x32x40", "UNET", 2, 3, 32, 40, 5, 0), ("cubenet_att_2x238x34x42", "CubeNET", 2, 238, 34, 42, 6, 0),
    ("cubenet_fd32_2x238x32x40", "CubeNET", 2, 238, 32, 40, 7, 0), ("cubenet_fd128_att_1x238x34x42", "CubeNET", 1, 238, 34, 42, 8, 0),
    ("unet_bil_2x3x34x42", "UNET", 2, 3, 34, 42, 9, 0), ("cubenet_bil_att_2x238x32x40", "CubeNET", 2, 238, 32, 40, 10, 0)])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_against_reference_golden(name, model, n, bands, h, w, seed, feats, mode):
    """Reference-module outputs (tests/golden, made by oracle/gen_golden.py), held to the north_star 1e-2 of max|logit|
    in train mode and 3e-3 in eval mode (running statistics; measured <= 9.2e-4).  Exception: use_attention=True in train
    mode stays at 3e-2 -- two fp16-stored operands are multiplied at every decoder level and these shapes normalise over
    as few as 8 samples per channel (measured 1.4e-2 .. 2.0e-2; 7.2e-3 at most for every other case)."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    fd = int(name.split("_fd")[1].split("_")[0]) if "_fd" in name else 64
    net, _ = build(model, bands, feats, seed, att="_att_" in name, fd=fd, bil="_bil_" in name, bnorm="_nobn_" not in name)
    x = O.synth_cube(seed, n, bands, h, w)
    xin = x[:, None] if model == "CubeNET" else x
    mask = O.synth_mask(seed, n, h, w)
    with torch.set_grad_enabled(mode == "train"):
        lg, loss = run_ours(net, xin, mask, train=(mode == "train"))
    ref = torch.from_numpy(g[f"{mode}.logits"])
    record(test="against_reference_golden", name=name, mode=mode,
           logit_max_rel_err=(lg - ref).abs().max().item() / ref.abs().max().item(), loss_abs_err=abs(loss - float(g[f"{mode}.loss"])))
    tol = 3e-3 if mode == "eval" else (3e-2 if "_att_" in name else 1e-2)
    assert (lg - ref).abs().max().item() <= tol * ref.abs().max().item()
    assert abs(loss - float(g[f"{mode}.loss"])) < (2e-3 if "_att_" in name else 3e-4)


def test_odd_sizes_pad_and_pool_floor():
    """W/8 odd -> the ConvTranspose output is one column short of the skip and is zero padded on the right
    (model_parts.py:77-80); MaxPool floors."""
    net, sd = build("UNET", 3)
    for h, w in ((48, 72), (40, 88), (34, 50)):
        x = O.synth_cube(5, 1, 3, h, w)
        mask = O.synth_mask(5, 1, h, w)
        ol = O.forward_backward("UNET", x, mask, sd, training=True)[0]
        lg, _ = run_ours(net, x, mask)
        net.load_state_dict(sd)           # undo the running-stat update
        assert (lg - ol).abs().max().item() <= 2e-2 * ol.abs().max().item(), (h, w)


def test_eval_mode_uses_running_stats_and_no_grad_path():
    net, sd = build("CubeNET", 238)
    x = O.synth_cube(1, 2, 238, 64, 80)[:, None]
    mask = O.synth_mask(1, 2, 64, 80)
    ol = O.forward_backward("CubeNET", x, mask, sd, training=False)[0]
    net.eval()
    with torch.no_grad():
        lg = net(x.cuda()).cpu()
    assert (lg - ol).abs().max().item() <= 1e-2 * ol.abs().max().item()
    assert torch.equal(net.state_dict()["inc.1.running_mean"].cpu(), sd["inc.1.running_mean"])


def test_state_dict_roundtrip_and_optimizer_step_changes_output():
    net, sd = build("UNET", 3)
    out = net.state_dict()
    assert set(out) == set(sd) and all(torch.equal(out[k].cpu(), sd[k]) for k in sd if "num_batches" not in k)
    x, mask = O.synth_cube(2, 2, 3, 64, 64), O.synth_mask(2, 2, 64, 64)
    from hyperpri_b200.optim import FusedAdam
    opt = FusedAdam(net.parameters(), lr=1e-3)
    losses = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.BCEWithLogitsLoss()(net(x.cuda()), mask.cuda())
        loss.backward()
        opt.step()                       # in-place update bumps param versions -> operands are re-packed
        losses.append(loss.item())
    assert losses[-1] < losses[0] - 0.05, losses


@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_fused_adam_matches_torch_adam(wd):
    """hpri_adam_step (one launch for all tensors) == torch.optim.Adam step for step; state_dicts interchange."""
    from hyperpri_b200.optim import FusedAdam
    g = torch.Generator(device="cpu").manual_seed(5)
    shapes = [(64, 238, 3, 3), (1000,), (7,), (3, 5, 2, 2), (1025,)]
    pa = [torch.randn(s, generator=g).cuda().requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa, ob = FusedAdam(pa, lr=1e-2, weight_decay=wd), torch.optim.Adam(pb, lr=1e-2, weight_decay=wd)
    for it in range(5):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).cuda()
            a.grad, b.grad = gr.clone(), gr.clone()
        v0 = pa[0]._version
        oa.step(); ob.step()
        assert pa[0]._version > v0                     # version bump -> the engine re-packs its fp16 operands
        for a, b in zip(pa, pb):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (it, (a - b).abs().max())
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa["state"].keys() == sb["state"].keys()
    for k in sa["state"]:
        assert set(sa["state"][k]) == {"step", "exp_avg", "exp_avg_sq"}
        assert torch.allclose(sa["state"][k]["exp_avg_sq"], sb["state"][k]["exp_avg_sq"], rtol=1e-5, atol=1e-9)
    ob.load_state_dict(sa)                             # a FusedAdam checkpoint resumes under torch.optim.Adam


def test_partial_weight_change_and_fp16_host_input():
    """A single changed parameter goes through the per-layer re-pack path (the table-driven launch needs every
    layer stale); an fp16 cube (host-side conversion before H2D) gives bit-identical logits."""
    net, sd = build("CubeNET", 238)
    x = O.synth_cube(3, 1, 238, 48, 40)[:, None].cuda()
    net.eval()
    with torch.no_grad():
        a = net(x)
        assert torch.equal(net(x.half()), a)
        with torch.no_grad():
            net.up3.conv.double_conv[0].weight.mul_(1.5)        # bumps one parameter version
        b = net(x)
        net2, _ = build("CubeNET", 238)
        sd2 = {k: v.clone() for k, v in sd.items()}
        sd2["up3.conv.double_conv.0.weight"] *= 1.5
        net2.load_state_dict(sd2)
        net2.cuda().eval()
        assert not torch.equal(a, b) and torch.equal(net2(x), b)


def test_full_size_properties():
    """BASELINE size (2 x 238 x 608 x 968): size-independent properties instead of an oracle run --
    finite logits, loss consistent with the logits, per-image independence of the forward in eval mode."""
    net, _ = build("CubeNET", 238)
    x = torch.rand((2, 1, 238, 608, 968), device="cuda")
    mask = (torch.rand((2, 1, 608, 968), device="cuda") > 0.95).float()
    net.train()
    lg = net(x)
    loss = torch.nn.BCEWithLogitsLoss()(lg, mask)
    loss.backward()
    assert lg.shape == (2, 1, 608, 968) and torch.isfinite(lg).all() and torch.isfinite(loss)
    assert all(torch.isfinite(p.grad).all() for p in net.parameters())
    net.eval()
    with torch.no_grad():
        both = net(x)
        one = net(x[1:2].contiguous())
    assert torch.equal(both[1:2], one)            # eval-mode BN: images are independent, bit for bit


@pytest.mark.parametrize("model,n,bands,h,w", [("UNET", 2, 3, 192, 272), ("CubeNET", 2, 238, 162, 210)])
def test_use_attention_train_step_parity_vs_oracle(model, n, bands, h, w):
    """use_attention=True (model_parts.py:65-66, 84-85): the decoder blocks convolve skip * up; odd sizes exercise the
    zero padding of the upsampled operand."""
    net, sd = build(model, bands, att=True)
    x = O.synth_cube(1, n, bands, h, w)
    xin = x[:, None] if model == "CubeNET" else x
    mask = O.synth_mask(1, n, h, w)
    torch.set_num_threads(os.cpu_count())
    ol, oloss, og, _ = O.forward_backward(model, xin, mask, sd, training=True, attention=True)
    lg, loss = run_ours(net, xin, mask)
    # the product of two fp16-stored activations at every decoder level compounds the storage rounding, and the
    # worst pixel moves from run to run with the order of the statistics atomics (measured 1.1e-2 .. 2.1e-2 of
    # max|logit| at small sizes, rms 2e-3 .. 3e-3): this flag is held to 3e-2 max / 5e-3 rms (the configured path,
    # without attention, to 1e-2 above)
    err = (lg - ol).abs()
    assert err.max().item() <= 3e-2 * ol.abs().max().item()
    assert err.pow(2).mean().sqrt().item() <= 5e-3 * ol.abs().max().item()
    assert abs(loss - oloss.item()) < 2e-4
    flat_o, flat_g = [], []
    for k, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        assert p.grad.shape == og[k].shape, k
        flat_o.append(og[k].flatten()); flat_g.append(p.grad.cpu().flatten())
    assert cos(torch.cat(flat_g), torch.cat(flat_o)) > 0.97
    for k in ("up1.conv.double_conv.0.weight", "up4.conv.double_conv.0.weight", "up1.up.weight", "down4.maxpool_conv.1.double_conv.3.weight"):
        assert cos(dict(net.named_parameters())[k].grad.cpu(), og[k]) > 0.9, k


@pytest.mark.parametrize("fd,att,h,w", [(32, False, 160, 208), (128, False, 97, 131), (16, True, 128, 160)])
def test_cubenet_first_depth_train_step_parity_vs_oracle(fd, att, h, w):
    """CubeNET first_depth != 64 (models.py:193-199, 229-240): first_conv / inc2 with first_depth maps (below and above
    one 64-channel tile), last decoder block `upsample4` / `upconv4` over cat([x1, up]); with use_attention the other
    three blocks multiply and the last still concatenates."""
    net, sd = build("CubeNET", 238, att=att, fd=fd)
    assert ("upsample4.weight" in sd) and ("up4.up.weight" not in sd)
    x = O.synth_cube(2, 2, 238, h, w)[:, None]
    mask = O.synth_mask(2, 2, h, w)
    torch.set_num_threads(os.cpu_count())
    ol, oloss, og, ostats = O.forward_backward("CubeNET", x, mask, sd, training=True, attention=att)
    lg, loss = run_ours(net, x, mask)
    err = (lg - ol).abs()
    assert err.max().item() <= (3e-2 if att else 1e-2) * ol.abs().max().item()
    assert abs(loss - oloss.item()) < 2e-4
    flat_o, flat_g = [], []
    for k, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.shape == og[k].shape, k
        flat_o.append(og[k].flatten()); flat_g.append(p.grad.cpu().flatten())
    assert cos(torch.cat(flat_g), torch.cat(flat_o)) > 0.97
    for k in ("first_conv.weight", "inc2.0.weight", "upsample4.weight", "upconv4.double_conv.0.weight",
              "down1.maxpool_conv.1.double_conv.0.weight"):
        assert cos(dict(net.named_parameters())[k].grad.cpu(), og[k]) > 0.9, k
    bufs = dict(net.named_buffers())
    for k, v in ostats.items():
        if "running_" in k:
            assert torch.allclose(bufs[k].cpu(), v, rtol=5e-3, atol=5e-4), k


@pytest.mark.parametrize("model,att,h,w", [("UNET", False, 162, 210), ("CubeNET", False, 128, 176), ("UNET", True, 192, 272)])
def test_bilinear_train_step_parity_vs_oracle(model, att, h, w):
    """bilinear=True (model_parts.py:56-61; models.py:33,43-49): nn.Upsample(x2, bilinear, align_corners) in place of
    the ConvTranspose, DoubleConvs with mid channels, down4 and the decoder outputs halved."""
    bands = 3 if model == "UNET" else 238
    net, sd = build(model, bands, att=att, bil=True)
    assert not any(".up." in k for k in sd)                          # nn.Upsample has no parameters
    assert sd["up1.conv.double_conv.3.weight"].shape == (256, 512, 3, 3)
    x = O.synth_cube(3, 2, bands, h, w)
    xin = x[:, None] if model == "CubeNET" else x
    mask = O.synth_mask(3, 2, h, w)
    torch.set_num_threads(os.cpu_count())
    ol, oloss, og, ostats = O.forward_backward(model, xin, mask, sd, training=True, attention=att)
    lg, loss = run_ours(net, xin, mask)
    err = (lg - ol).abs()
    record(test="bilinear_train_step_parity_vs_oracle", model=model, att=att, shape=[2, bands, h, w],
           logit_max_rel_err=err.max().item() / ol.abs().max().item(), logit_rms_rel_err=err.pow(2).mean().sqrt().item() / ol.abs().max().item())
    # bilinear=True is an unconfigured constructor flag.  Its align-corners interpolation reads
    # fp16-stored activations and stores fp16 again at every decoder level, one more rounding per level than the
    # ConvTranspose path, and the worst pixel of these small odd-sized cases lands at 0.9e-2 .. 1.1e-2 of max|logit|
    # (rms 1e-3): held to 1.5e-2 max / 3e-3 rms here (attention: 3e-2, as above); the configured path to 1e-2.
    assert err.max().item() <= (3e-2 if att else 1.5e-2) * ol.abs().max().item()
    assert err.pow(2).mean().sqrt().item() <= (5e-3 if att else 3e-3) * ol.abs().max().item()
    assert abs(loss - oloss.item()) < 2e-4
    flat_o, flat_g = [], []
    for k, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.shape == og[k].shape, k
        flat_o.append(og[k].flatten()); flat_g.append(p.grad.cpu().flatten())
    assert cos(torch.cat(flat_g), torch.cat(flat_o)) > 0.97
    for k in ("up1.conv.double_conv.0.weight", "up4.conv.double_conv.3.weight", "down4.maxpool_conv.1.double_conv.3.weight",
              "down1.maxpool_conv.1.double_conv.0.weight"):
        assert cos(dict(net.named_parameters())[k].grad.cpu(), og[k]) > 0.9, k


@pytest.mark.parametrize("model,n,bands,h,w", [("CubeNET", 2, 238, 96, 136), ("UNET", 2, 3, 128, 160)])
def test_gradients_per_parameter_vs_storage_emulating_oracle(model, n, bands, h, w):
    """Per-parameter gradient error against the oracle run with the B200 path's storage precision emulated (fp16
    activations / weights, fp16 gradients under the engine's power-of-two loss scale), with a bound calibrated by the
    oracle itself: random-init ReLU networks have non-smooth, cancellation-dominated gradients -- perturbing the INPUT
    of the fp32 oracle by 1e-4 (2e-4 relative) already moves its own per-parameter gradients by ~10 % (median), which
    is why no tight absolute bound exists at the whole-network level (the per-kernel tests in test_kernels_gpu.py hold
    dgrad / wgrad / BatchNorm-backward to 2e-4 .. 6e-3).  Asserted here, per parameter tensor:
        relL2(ours, emulated oracle) <= max(0.08, 2.5 * relL2(perturbed fp32 oracle, fp32 oracle)),
    plus a hard ceiling of 0.35 and a global cosine > 0.97."""
    import math
    net, sd = build(model, bands)
    x = O.synth_cube(0, n, bands, h, w)
    xin = x[:, None] if model == "CubeNET" else x
    mask = O.synth_mask(0, n, h, w)
    torch.set_num_threads(os.cpu_count())
    _, _, g32, _ = O.forward_backward(model, xin, mask, sd, training=True)
    g = torch.Generator().manual_seed(7)
    _, _, gpert, _ = O.forward_backward(model, xin + 1e-4 * torch.randn(xin.shape, generator=g), mask, sd, training=True)
    S = float(2.0 ** (math.ceil(math.log2(n * h * w)) - 4))            # engine.loss_scale()
    O.emulate_bf16_storage(True, S)
    try:
        _, _, gemu, _ = O.forward_backward(model, xin, mask, sd, training=True)
    finally:
        O.emulate_bf16_storage(False)
    run_ours(net, xin, mask)

    def rel(a, b):
        return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
    worst, rows = 0.0, []
    for k, p in net.named_parameters():
        if k.startswith("inc.0."):
            continue
        # conv biases in front of train-mode BatchNorm have an exactly-zero gradient here and ~1e-9 noise in the oracle
        if k.endswith(".bias") and og_is_noise(g32[k]):
            assert p.grad.abs().max().item() <= 1e-6
            continue
        e, sens = rel(p.grad.cpu(), gemu[k]), rel(gpert[k], g32[k])
        rows.append((k, e, sens))
        assert e <= max(0.08, 2.5 * sens), (k, e, sens)
        assert e <= 0.35, (k, e)
        worst = max(worst, e)
    es = sorted(r[1] for r in rows)
    record(test="gradients_per_parameter", model=model, shape=[n, bands, h, w], median_rel_l2=es[len(es) // 2], max_rel_l2=worst,
           median_oracle_sensitivity=sorted(r[2] for r in rows)[len(rows) // 2])
    flat_g = torch.cat([p.grad.cpu().flatten() for k, p in net.named_parameters() if not k.startswith("inc.0.")])
    flat_o = torch.cat([gemu[k].flatten() for k, p in net.named_parameters() if not k.startswith("inc.0.")])
    assert cos(flat_g, flat_o) > 0.97


def og_is_noise(g):
    return g.abs().max().item() < 1e-7


def test_bce_step_is_the_fused_training_step_and_gradients_are_zero_copy():
    """nn.Module.bce_step (what RootLightningModel.training_step runs for the configured BCEWithLogitsLoss): same loss
    as the criterion on the module's logits, TP/FP/FN/TN of sigmoid(logits) > thr, gradients equal to the generic
    autograd route, delivered as views of the engine's arena (no copies); accumulation semantics stay exact."""
    net, sd = build("CubeNET", 238)
    x = O.synth_cube(4, 2, 238, 64, 96)[:, None].cuda()
    mask = O.synth_mask(4, 2, 64, 96).cuda()
    net.train()
    net.zero_grad(set_to_none=True)
    loss, logits, counts = net.bce_step(x, mask, thr=0.4)
    loss.backward()
    torch.cuda.synchronize()
    eng = net._get_engine(x.device)
    g_fused = {k: p.grad.clone() for k, p in net.named_parameters()}
    for k, p in net.named_parameters():
        assert p.grad.data_ptr() == eng.grads[k].data_ptr(), k            # arena views, not clones
    ref_loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, mask)
    assert abs(loss.item() - ref_loss.item()) < 1e-6 and not logits.requires_grad
    seg = torch.sigmoid(logits) > 0.4
    tp, fp = (seg & (mask > 0)).sum().item(), (seg & ~(mask > 0)).sum().item()
    assert counts.tolist()[:2] == [tp, fp] and counts.sum().item() == mask.numel()
    # generic route (any criterion): module logits -> torch loss -> autograd
    net.load_state_dict(sd)
    net.zero_grad(set_to_none=True)
    l2 = torch.nn.BCEWithLogitsLoss()(net(x), mask)
    l2.backward()
    torch.cuda.synchronize()
    flat = lambda d: torch.cat([d[k].flatten() for k in sorted(d)])
    g_gen = {k: p.grad.clone() for k, p in net.named_parameters()}
    assert cos(flat(g_fused), flat(g_gen)) > 0.999 and abs(l2.item() - loss.item()) < 1e-6
    # a second backward WITHOUT zero_grad accumulates (p.grad still aliases the arena: snapshot + add keeps it exact)
    net.load_state_dict(sd)
    l3 = torch.nn.BCEWithLogitsLoss()(net(x), mask)
    l3.backward()
    torch.cuda.synchronize()
    g_acc = {k: p.grad.clone() for k, p in net.named_parameters()}
    assert cos(flat(g_acc), flat(g_gen)) > 0.999
    ratio = (flat(g_acc).norm() / flat(g_gen).norm()).item()
    assert abs(ratio - 2.0) < 0.05, ratio
    # zero_grad(set_to_none=False) keeps the aliasing tensors and zeroes them: the next backward must not double
    net.load_state_dict(sd)
    net.zero_grad(set_to_none=False)
    l4, _, _ = net.bce_step(x, mask)
    l4.backward()
    torch.cuda.synchronize()
    g4 = {k: p.grad.clone() for k, p in net.named_parameters()}
    assert abs((flat(g4).norm() / flat(g_gen).norm()).item() - 1.0) < 0.05
    assert int(eng.overflow.item()) == 0
    # eval / no-grad path of the same call
    net.eval()
    with torch.no_grad():
        le, lg_e, ce = net.bce_step(x, mask)
    assert abs(le.item() - torch.nn.functional.binary_cross_entropy_with_logits(lg_e, mask).item()) < 1e-6
    assert ce.sum().item() == mask.numel()


def test_generic_criterion_route_scales_any_loss_gradient():
    """A sum-reduced criterion has |dlogit| up to 1 (2^20 times the mean-reduced bound the static loss scale assumes):
    the generic route picks its power-of-two scale from the measured gradient, so nothing overflows and the
    gradients are those of the mean-reduced run times numel."""
    net, sd = build("UNET", 3)
    x = O.synth_cube(6, 1, 3, 64, 64).cuda()
    mask = O.synth_mask(6, 1, 64, 64).cuda()
    net.train()
    net.zero_grad(set_to_none=True)
    torch.nn.BCEWithLogitsLoss(reduction="sum")(net(x), mask).backward()
    eng = net._get_engine(x.device)
    assert int(eng.overflow.item()) == 0
    gs = torch.cat([p.grad.flatten().clone() for _, p in sorted(net.named_parameters())])
    assert torch.isfinite(gs).all()
    net.load_state_dict(sd)
    net.zero_grad(set_to_none=True)
    torch.nn.BCEWithLogitsLoss()(net(x), mask).backward()
    gm = torch.cat([p.grad.flatten().clone() for _, p in sorted(net.named_parameters())])
    assert cos(gs, gm) > 0.999
    assert abs((gs.norm() / gm.norm()).item() / mask.numel() - 1.0) < 0.05


def test_next_batch_ingest_prefetch_gives_the_same_network_input():
    """set_next_input(x2) before a backward: x2 is ingested on a third stream under that backward and the next forward(x2)
    starts from the ready buffer (bit-identical NHWC input); a forward on any OTHER tensor ignores the prefetched buffer."""
    net, sd = build("CubeNET", 238)
    x1 = O.synth_cube(7, 2, 238, 48, 64)[:, None].cuda()
    x2 = O.synth_cube(8, 2, 238, 48, 64)[:, None].cuda()
    x3 = O.synth_cube(9, 2, 238, 48, 64)[:, None].cuda()
    mask = O.synth_mask(7, 2, 48, 64).cuda()
    net.train()
    eng = net._get_engine(x1.device)
    loss, _, _ = net.bce_step(x1, mask)
    net.set_next_input(x2)
    loss.backward()
    assert eng._prefetched is not None
    other = eng.ws["x_alt"]
    net.eval()                                   # running statistics: the logits depend on the input only
    with torch.no_grad():
        a = net(x2)                              # takes the prefetched buffer
        assert eng.ws["x"] is other and eng._prefetched is None
        xin_pref = eng.ws["x"].clone()
        b = net(x2)                              # plain ingest of the same tensor
        assert torch.equal(eng.ws["x"], xin_pref) and torch.equal(a, b)
    net.train()
    loss, _, _ = net.bce_step(x1, mask)
    net.set_next_input(x2)
    loss.backward()
    net.eval()
    with torch.no_grad():
        c = net(x3)                              # not the registered tensor: ingested normally
        net.load_state_dict(net.state_dict())
        d = net(x3)
    assert torch.equal(c, d) and not torch.equal(c, a)
    torch.cuda.synchronize()


@pytest.mark.parametrize("model,bands,h,w,feats", [("CubeNET", 238, 96, 136, 0), ("UNET", 3, 160, 200, 0), ("SpectralUNET", 238, 24, 40, 1650)])
def test_deterministic_statistics_mode_is_bit_reproducible(model, bands, h, w, feats):
    """ops.set_deterministic(True) / HPRI_DETERMINISTIC=1: per-CTA partial BatchNorm statistics are combined in a fixed
    order instead of by atomics, so repeated train-mode forwards are bit-identical (the default mode differs from run to
    run by ~1e-3 of max|logit| at the worst pixel); the result stays within the parity tolerance of the oracle."""
    from hyperpri_b200 import ops
    net, sd = build(model, bands, feats or 1650)
    x = O.synth_cube(0, 2, bands, h, w)
    xin = (x[:, None] if model == "CubeNET" else x).cuda()
    mask = O.synth_mask(0, 2, h, w)
    ops.set_deterministic(True)
    try:
        net.train()
        outs = []
        for _ in range(3):
            with torch.no_grad():
                outs.append(net(xin).clone())
        torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
        ol = O.forward_backward(model, xin.cpu(), mask, sd, training=True)[0]
        assert (outs[0].cpu() - ol).abs().max().item() <= 1e-2 * ol.abs().max().item()
        # a full training step in the fully deterministic mode leaves finite gradients (its reproducibility:
        # test_deterministic_mode_training_is_bit_reproducible)
        lg, loss = run_ours(net, xin.cpu(), mask)
        assert all(torch.isfinite(p.grad).all() for p in net.parameters())
    finally:
        ops.set_deterministic(False)
    net.load_state_dict(sd)
    with torch.no_grad():
        plain = net(xin)
    assert (plain - outs[0]).abs().max().item() <= 5e-3 * outs[0].abs().max().item()


DET_CASES = [("CubeNET", 238, 2, 37, 51, 0, {}), ("CubeNET", 238, 1, 96, 136, 0, dict(att=True, fd=32)),
             ("UNET", 3, 2, 50, 34, 0, dict(bil=True)), ("UNET", 3, 2, 160, 200, 0, {}),
             ("SpectralUNET", 238, 2, 9, 13, 96, {}), ("SpectralUNET", 238, 2, 24, 40, 1650, dict(bnorm=False))]


@pytest.mark.parametrize("model,bands,n,h,w,feats,flags", DET_CASES)
def test_deterministic_mode_training_is_bit_reproducible(model, bands, n, h, w, feats, flags):
    """ops.set_deterministic(True) (HPRI_DETERMINISTIC=1; the reference's Trainer(deterministic='warn'),
    PLTrainer.py:430,439,447): three training steps -- fused BCE step, backward, FusedAdam -- from the same weights on
    the same batches give bit-identical losses, logits, gradients and updated parameters on every run, and the
    gradients of this mode (unfused BatchNorm-backward reduction, unsplit weight gradients, single-CTA bias sums)
    agree with the production backward pass to the fp16-rounding level of its run-to-run spread."""
    from hyperpri_b200 import ops
    from hyperpri_b200.optim import FusedAdam
    x = O.synth_cube(5, n, bands, h, w)
    xin = (x[:, None] if model == "CubeNET" else x).cuda()
    mask = O.synth_mask(5, n, h, w).cuda()

    def train():
        net, _ = build(model, bands, feats or 1650, seed=4, **flags)
        net.train()
        opt = FusedAdam(net.parameters(), lr=1e-3)
        trace = []
        for it in range(3):
            opt.zero_grad(set_to_none=True)
            loss, logits, counts = net.bce_step(xin.roll(it, -1), mask.roll(it, -1), 0.5)
            loss.backward()
            trace.append((loss.detach().clone(), logits.detach().clone(), counts.clone(),
                          [p.grad.detach().clone() for p in net.parameters()]))
            opt.step()
        torch.cuda.synchronize()
        return trace, [p.detach().clone() for p in net.parameters()], [k for k, _ in net.named_parameters()]

    ops.set_deterministic(True)
    try:
        ta, pa, names = train()
        tb, pb, _ = train()
    finally:
        ops.set_deterministic(False)
    for it, ((la, lga, ca, ga), (lb, lgb, cb, gb)) in enumerate(zip(ta, tb)):
        assert torch.equal(la, lb) and torch.equal(lga, lgb) and torch.equal(ca, cb), it
        for k, u, v in zip(names, ga, gb):
            assert torch.isfinite(u).all(), (it, k)
            assert torch.equal(u, v), (it, k, (u - v).abs().max().item())
    for k, u, v in zip(names, pa, pb):
        assert torch.equal(u, v), k
    ops.set_deterministic(True, backward=False)          # the production backward pass behind the same forward pass
    try:
        tc, _, _ = train()
    finally:
        ops.set_deterministic(False)
    assert torch.equal(tc[0][1], ta[0][1])
    flat = lambda g: torch.cat([t.flatten() for t in g]).double()
    d, c = flat(ta[0][3]), flat(tc[0][3])
    rel = ((d - c).norm() / c.norm()).item()
    record(test="deterministic_mode_training", model=model, shape=[n, bands, h, w], flags=str(flags),
           grad_rel_l2_vs_production_backward=rel)
    assert rel <= 5e-3


@pytest.mark.parametrize("feats,h,w", [(96, 24, 40), (1650, 16, 33)])
def test_spectralunet_without_batchnorm_train_step_parity_vs_oracle(feats, h, w):
    """SpectralUNET(bnorm=False) (models.py:72,105-110): Linear -> ReLU blocks through the same kernels (apply pass with
    scale 1 / shift = bias; backward apply with zero sums = ReLU backward; bias gradients by column sums)."""
    net, sd = build("SpectralUNET", 238, feats, bnorm=False)
    assert not any(".1." in k for k in sd) and len(sd) == 20
    x = O.synth_cube(5, 2, 238, h, w)
    mask = O.synth_mask(5, 2, h, w)
    torch.set_num_threads(os.cpu_count())
    ol, oloss, og, _ = O.forward_backward("SpectralUNET", x, mask, sd, training=True)
    lg, loss = run_ours(net, x, mask)
    assert (lg - ol).abs().max().item() <= 1e-2 * ol.abs().max().item()
    assert abs(loss - oloss.item()) < 2e-4
    fo, fg = [], []
    for k, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.shape == og[k].shape, k
        fo.append(og[k].flatten()); fg.append(p.grad.cpu().flatten())
        if og[k].norm() > 0:
            assert cos(p.grad.cpu(), og[k]) > 0.98, k            # no BatchNorm: smooth, well-conditioned gradients
    assert cos(torch.cat(fg), torch.cat(fo)) > 0.995
    # second step accumulates nothing stale (bias gradients are written with beta = 0 for the first image)
    lg2, loss2 = run_ours(net, x, mask)
    assert abs(loss2 - loss) < 1e-6
