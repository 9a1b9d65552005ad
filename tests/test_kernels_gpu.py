"""Per-kernel parity on a B200: every C-ABI entry point against a plain PyTorch fp32 reference of
the same op on identical (bf16-rounded) inputs.  Tolerances are stated per test."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from hyperpri_b200 import ops  # noqa: E402

DEV = "cuda"


@pytest.fixture(autouse=True, scope="module")
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


BF, FH = torch.bfloat16, torch.float16


def nhwc(x_nchw, cpad=None, dt=BF):
    n, c, h, w = x_nchw.shape
    cpad = cpad or c
    out = torch.zeros((n, h, w, cpad), dtype=dt, device=x_nchw.device)
    out[..., :c] = x_nchw.permute(0, 2, 3, 1).to(dt)
    return out


def nchw(x_nhwc, c=None):
    x = x_nhwc.float().permute(0, 3, 1, 2)
    return x if c is None else x[:, :c]


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def relerr(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


CONV_CASES = [
    # n, h, w, cin, cout, block_n
    (2, 16, 24, 64, 64, 0),
    (1, 20, 40, 240, 64, 0),
    (2, 9, 13, 128, 256, 0),
    (1, 11, 70, 128, 128, 0),
    (1, 38, 60, 512, 1024, 0),
    (1, 16, 16, 8, 64, 0),
    (1, 24, 40, 256, 256, 128),
]


@pytest.fixture(params=[0, 1, 2, -1], ids=["generic", "halo", "halo_pair", "default"])
def conv_algo(request):
    """Every 3x3 kernel the launcher can pick: generic per-tap, halo-reuse on single CTAs, halo-reuse on CTA pairs
    (tcgen05 cta_group::2 -- what the heuristic picks for every production layer), and the heuristic itself."""
    ops.set_conv_algo(request.param)
    yield request.param
    ops.set_conv_algo(-1)


@pytest.fixture(params=[-1, 0], ids=["wgrad_default", "wgrad_generic"])
def wgrad_algo(request):
    """Weight gradients with the halo-reuse kernel allowed (the production choice at Cout 64 / 128) and forced off."""
    ops.set_wgrad_algo(request.param)
    yield request.param
    ops.set_wgrad_algo(-1)


HALO_CASES = [(2, 32, 24, 64, 64, 0), (1, 40, 50, 128, 64, 0), (1, 33, 70, 240, 64, 0), (2, 16, 33, 64, 128, 0),
              (1, 70, 37, 128, 256, 0), (1, 8, 96, 256, 128, 0)]


@pytest.mark.parametrize("dt", [BF, FH])
@pytest.mark.parametrize("n,h,w,cin,cout,bn", CONV_CASES + HALO_CASES)
def test_conv3x3_fwd_and_stats(n, h, w, cin, cout, bn, dt, conv_algo):
    x = rnd(n, cin, h, w, seed=1)
    wt = rnd(cout, cin, 3, 3, scale=1 / math.sqrt(cin * 9), seed=2)
    xb = nhwc(x, dt=dt)
    spec = ops.WeightSpec("conv3x3", cout, cin)
    wp = spec.pack_fwd(wt, dtype=dt)
    y = torch.full((n, h, w, cout), float("nan"), dtype=dt, device=DEV)
    stats = torch.zeros((cout, 2), dtype=torch.float64, device=DEV)
    ops.igemm_fwd(xb, wp, cout, 9, y, cout, stats=stats, block_n=bn)
    torch.cuda.synchronize()
    ref = F.conv2d(nchw(xb), wt.to(dt).float(), padding=1)
    got = nchw(y)
    assert torch.isfinite(got).all()
    # output rounding: half-ulp relative 2^-9 (bf16) / 2^-12 (fp16); fp32 accumulation order differences
    assert relerr(got, ref) < (6e-3 if dt == BF else 8e-4)
    g64 = y.double().reshape(-1, cout)
    assert torch.allclose(stats[:, 0], g64.sum(0), rtol=1e-6, atol=1e-3)
    assert torch.allclose(stats[:, 1], (g64 * g64).sum(0), rtol=1e-6, atol=1e-3)


@pytest.mark.parametrize("n,h,w,cin,cout,bn", CONV_CASES[:5] + HALO_CASES[:3])
def test_conv3x3_dgrad(n, h, w, cin, cout, bn, conv_algo):
    dy = rnd(n, cout, h, w, seed=3)
    wt = rnd(cout, cin, 3, 3, scale=1 / math.sqrt(cout * 9), seed=4)
    dyb = nhwc(dy)
    spec = ops.WeightSpec("conv3x3", cout, cin)
    wp = spec.pack_dgrad(wt, dtype=BF)
    cpad = (cin + 7) // 8 * 8
    dx = torch.full((n, h, w, cpad), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.igemm_fwd(dyb, wp, cin, 9, dx, cpad, block_n=bn)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(nchw(dyb), wt.to(torch.bfloat16).float(), padding=1)
    assert relerr(nchw(dx, cin), ref) < 6e-3


@pytest.mark.parametrize("xdt,gdt", [(BF, BF), (FH, FH)])
@pytest.mark.parametrize("n,h,w,cin,cout,bn", CONV_CASES + HALO_CASES)
def test_conv3x3_wgrad(n, h, w, cin, cout, bn, xdt, gdt, wgrad_algo):
    """Both operands of one tcgen05.mma kind::f16 must share a format (mixed f16 x bf16 faults on B200)."""
    x = rnd(n, cin, h, w, seed=5)
    dy = rnd(n, cout, h, w, seed=6)
    xb, dyb = nhwc(x, dt=xdt), nhwc(dy, dt=gdt)
    spec = ops.WeightSpec("conv3x3", cout, cin)
    dwp = spec.grad_buffer(DEV)
    ops.igemm_wgrad(xb, dyb, 1, cout, dwp, block_n=bn)
    dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=DEV)
    spec.unpack_grad(dwp, dw)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(nchw(xb), (cout, cin, 3, 3), nchw(dyb), padding=1)
    # fp32 accumulate of exact bf16 products; split-K changes summation order only
    assert relerr(dw, ref) < 2e-4


CONVT_CASES = [(2, 8, 12, 128, 64, 16, 25), (1, 19, 30, 1024, 512, 38, 61), (1, 5, 7, 256, 128, 10, 14)]


@pytest.mark.parametrize("n,h,w,cin,cout,H2,W2", CONVT_CASES)
def test_convT(n, h, w, cin, cout, H2, W2):
    x = rnd(n, cin, h, w, seed=7)
    wt = rnd(cin, cout, 2, 2, scale=1 / math.sqrt(cin), seed=8)
    bias = rnd(cout, seed=9)
    xb = nhwc(x)
    spec = ops.WeightSpec("convT2x2", cout, cin)
    # destination: second half of a (2*cout)-channel concat buffer at skip resolution H2 x W2
    cat = torch.zeros((n, H2, W2, 2 * cout), dtype=torch.bfloat16, device=DEV)
    ops.convT_fwd(xb, spec.pack_fwd(wt, dtype=BF), cout, cat[..., cout:], bias=bias)
    torch.cuda.synchronize()
    wq = wt.to(torch.bfloat16).float()
    ref = F.conv_transpose2d(nchw(xb), wq, bias, stride=2)
    ref = F.pad(ref, [0, W2 - 2 * w, 0, H2 - 2 * h])
    assert relerr(nchw(cat[..., cout:]), ref) < 6e-3
    assert (cat[..., :cout] == 0).all()
    # dgrad / wgrad against autograd
    dcat = torch.zeros((n, H2, W2, 2 * cout), dtype=torch.bfloat16, device=DEV)
    dcat[..., cout:] = nhwc(rnd(n, cout, H2, W2, seed=10))
    dyv = dcat[..., cout:]
    dx = torch.full((n, h, w, cin), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.convT_dgrad(dyv, spec.pack_dgrad(wt, dtype=BF), cin, dx)
    dwp = spec.grad_buffer(DEV)
    ops.igemm_wgrad(xb, dyv, 2, 4 * cout, dwp)
    dw = torch.empty_like(wt)
    spec.unpack_grad(dwp, dw)
    db = torch.empty(cout, dtype=torch.float32, device=DEV)
    ops.colsum(dyv[:, :2 * h, :2 * w], db)
    torch.cuda.synchronize()
    xr = nchw(xb).requires_grad_(True)
    wr = wq.clone().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    out = F.conv_transpose2d(xr, wr, br, stride=2)
    out.backward(nchw(dyv)[:, :, :2 * h, :2 * w].contiguous())
    assert relerr(nchw(dx), xr.grad) < 6e-3
    assert relerr(dw, wr.grad) < 2e-4
    assert relerr(db, br.grad) < 1e-4


@pytest.mark.parametrize("m,fin,fout,split", [(3000, 238, 1650, 0), (1000, 1650, 1650, 0), (777, 3300, 1650, 1650),
                                              (4096, 64, 128, 0)])
def test_linear(m, fin, fout, split):
    """nn.Linear over M pixels (SpectralUNET blocks): fwd, dgrad, wgrad with padded feature strides."""
    pad = lambda f: (f + 63) // 64 * 64
    if split:
        xa, xb_ = rnd(m, split, seed=11), rnd(m, split, seed=12)
        buf = torch.zeros((1, 1, m, 2 * pad(split)), dtype=torch.bfloat16, device=DEV)
        buf[0, 0, :, :split] = xa.to(torch.bfloat16)
        buf[0, 0, :, pad(split):pad(split) + split] = xb_.to(torch.bfloat16)
        xfull = torch.cat([buf[0, 0, :, :split], buf[0, 0, :, pad(split):pad(split) + split]], 1).float()
    else:
        buf = torch.zeros((1, 1, m, pad(fin)), dtype=torch.bfloat16, device=DEV)
        buf[0, 0, :, :fin] = rnd(m, fin, seed=11).to(torch.bfloat16)
        xfull = buf[0, 0, :, :fin].float()
    wt = rnd(fout, fin, scale=1 / math.sqrt(fin), seed=13)
    spec = ops.WeightSpec("linear", fout, fin, split=split)
    y = torch.full((1, 1, m, pad(fout)), float("nan"), dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros((fout, 2), dtype=torch.float64, device=DEV)
    ops.igemm_fwd(buf, spec.pack_fwd(wt, dtype=BF), fout, 1, y, pad(fout), stats=stats)
    torch.cuda.synchronize()
    wq = wt.to(torch.bfloat16).float()
    ref = xfull @ wq.t()
    assert relerr(y[0, 0, :, :fout].float(), ref) < 6e-3
    assert (y[0, 0, :, fout:] == 0).all()          # pad features are written as exact zeros
    assert torch.allclose(stats[:, 0], y[0, 0, :, :fout].double().sum(0), rtol=1e-6, atol=1e-3)
    dy = torch.zeros((1, 1, m, pad(fout)), dtype=torch.bfloat16, device=DEV)
    dy[0, 0, :, :fout] = rnd(m, fout, seed=14).to(torch.bfloat16)
    dwp = spec.grad_buffer(DEV)
    ops.igemm_wgrad(buf, dy, 0, fout, dwp, dy_c=fout)
    dw = torch.empty_like(wt)
    spec.unpack_grad(dwp, dw)
    torch.cuda.synchronize()
    assert relerr(dw, dy[0, 0, :, :fout].float().t() @ xfull) < 2e-4
    if not split:
        dx = torch.full((1, 1, m, pad(fin)), float("nan"), dtype=torch.bfloat16, device=DEV)
        ops.igemm_fwd(dy, spec.pack_dgrad(wt, dtype=BF), fin, 1, dx, pad(fin), x_c=fout)
        torch.cuda.synchronize()
        assert relerr(dx[0, 0, :, :fin].float(), dy[0, 0, :, :fout].float() @ wq) < 6e-3


@pytest.mark.parametrize("cout,cin", [(64, 238), (128, 64), (96, 40), (1024, 512)])
def test_tiled_conv3x3_pack_matches_generic(cout, cin):
    w = rnd(cout, cin, 3, 3, seed=50)
    spec = ops.WeightSpec("conv3x3", cout, cin)
    f = torch.zeros((cout, 9 * ops.kpad(cin)), dtype=FH, device=DEV)
    d = torch.zeros((cin, 9 * ops.kpad(cout)), dtype=FH, device=DEV)
    spec.pack_both(w.reshape(-1), f, d)
    assert torch.equal(f, spec.pack_fwd(w.reshape(-1), dtype=FH))
    assert torch.equal(d, spec.pack_dgrad(w.reshape(-1), dtype=FH))
    g = torch.randn((cout, 9 * ops.kpad(cin)), device=DEV)
    a, b = torch.empty_like(w), torch.empty_like(w)
    g2 = g.clone()
    ops.unpack_conv3x3(g, cout, cin, a)
    ops.unpack(g2, b.view(-1), **spec.fwd)
    assert torch.equal(a, b)
    assert not g.view(cout, 9, -1)[:, :, :cin].any()      # the packed buffer is reset behind the read


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 32, 24, 64, 64), (1, 40, 50, 128, 64), (2, 16, 33, 64, 128),
                                            (1, 33, 70, 256, 128)])
def test_dgrad_fused_bn_backward_reduction(n, h, w, cin, cout):
    """igemm_fwd(bw=...) on a 3x3 dgrad launch == hpri_bn_relu_bwd_reduce on the dy it produced."""
    assert ops.conv3x3_halo_ok(h, w, cin)
    dyn = nhwc(rnd(n, cout, h, w, seed=21), dt=FH)                  # gradient entering the dgrad
    wt = rnd(cout, cin, 3, 3, scale=1 / math.sqrt(cout * 9), seed=22)
    wp = ops.WeightSpec("conv3x3", cout, cin).pack_dgrad(wt, dtype=FH)
    raw = nhwc(rnd(n, cin, h, w, seed=23), dt=FH)                   # raw conv output of the layer below
    scale = torch.rand(ops.kpad(cin), device=DEV) + 0.5
    shift = torch.randn(ops.kpad(cin), device=DEV) * 0.3
    mean = torch.randn(ops.kpad(cin), device=DEV) * 0.1
    inv = torch.rand(ops.kpad(cin), device=DEV) + 0.5
    sums = torch.full((cin, 3), 7.0, dtype=torch.float64, device=DEV)
    dx = torch.empty((n, h, w, cin), dtype=FH, device=DEV)
    ops.igemm_fwd(dyn, wp, cin, 9, dx, cin, bw=(raw, scale, shift, mean, inv, sums))
    dx_ref = torch.empty_like(dx)
    ops.igemm_fwd(dyn, wp, cin, 9, dx_ref, cin)
    assert torch.equal(dx, dx_ref)                                   # the stored gradient is unchanged
    ref = torch.zeros((cin, 3), dtype=torch.float64, device=DEV)
    from hyperpri_b200 import _lib
    import ctypes as C
    xv, dv = ops.view(raw), ops.view(dx)
    ops.check(_lib.lib().hpri_bn_relu_bwd_reduce(C.byref(xv), ops._ptr(scale), ops._ptr(shift), ops._ptr(mean),
                                                 ops._ptr(inv), C.byref(dv), None, None, None, ops._ptr(ref),
                                                 ops._stream()), "reduce")
    torch.cuda.synchronize()
    tol = 1e-4 * ref[:, :2].abs().max().item() + 1e-6
    assert (sums[:, :2] - ref[:, :2]).abs().max().item() < tol, (sums[:3], ref[:3])


def test_table_driven_pack_and_unpack_match_per_layer():
    """One launch over a device job table == the per-layer kernels, for ragged layer sizes."""
    shapes = [(64, 238), (128, 64), (96, 40), (256, 128)]
    ws = [rnd(co, ci, 3, 3, seed=60 + i) for i, (co, ci) in enumerate(shapes)]
    jobs, ref = [], []
    for (co, ci), w in zip(shapes, ws):
        spec = ops.WeightSpec("conv3x3", co, ci)
        f = torch.zeros((co, 9 * ops.kpad(ci)), dtype=FH, device=DEV)
        d = torch.zeros((ci, 9 * ops.kpad(co)), dtype=FH, device=DEV)
        g = torch.randn((co, 9 * ops.kpad(ci)), device=DEV)
        g.view(co, 9, -1)[:, :, ci:] = 0                      # padding columns of a packed gradient are always zero
        gd = torch.empty_like(w)
        jobs.append(dict(w=w.reshape(-1), fwd=f, dgrad=d, gpacked=g, gdst=gd, cout=co, cin=ci))
        gref = torch.empty_like(w)
        ops.unpack(g.clone(), gref.view(-1), **spec.fwd)
        ref.append((spec.pack_fwd(w.reshape(-1), dtype=FH), spec.pack_dgrad(w.reshape(-1), dtype=FH), gref))
    table = ops.Conv3x3JobTable(jobs, DEV)
    ops.pack_conv3x3_batch(table)
    ops.unpack_conv3x3_batch(table)
    torch.cuda.synchronize()
    for j, (f, d, g) in zip(jobs, ref):
        assert torch.equal(j["fwd"], f) and torch.equal(j["dgrad"], d) and torch.equal(j["gdst"], g)
        assert not j["gpacked"].any()                           # zeroed behind the read


@pytest.mark.parametrize("cin,cout", [(128, 64), (1024, 512), (64, 192)])
def test_tiled_convT_pack_matches_generic(cin, cout):
    w = rnd(cin, cout, 2, 2, seed=70)
    spec = ops.WeightSpec("convT2x2", cout, cin)
    got = spec.pack_fwd(w.reshape(-1), dtype=FH)                # tiled transpose
    want = ops.pack(w.reshape(-1), dtype=FH, **spec.fwd)        # generic index-map kernel
    assert torch.equal(got, want)
    g = torch.randn((4 * cout, ops.kpad(cin)), device=DEV)
    a, b = torch.empty_like(w), torch.empty_like(w)
    ops.unpack(g.clone(), b.view(-1), **spec.fwd)
    spec.unpack_grad(g, a)
    assert torch.equal(a, b) and not g.any()


def test_scale_check_unscales_and_flags_overflow():
    x = torch.randn(100003, device=DEV)
    want = x * 0.125
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    xs = x.clone()
    ops.scale_check(xs, 0.125, flag)
    assert torch.equal(xs, want) and flag.item() == 0
    xs[77777] = float("inf")
    ops.scale_check(xs, 1.0, flag)
    assert flag.item() == 1


def test_device_pr_curve_matches_torch_restatement():
    """hpri_pr_hist + DevicePRCurve == metrics.binned_pr_curve / confusion_counts on the concatenated predictions."""
    from hyperpri_b200 import metrics as M
    g = torch.Generator(device="cpu").manual_seed(11)
    curve = M.DevicePRCurve(torch.device(DEV), 500)
    all_l, all_m = [], []
    for n in (70001, 123457, 5):
        lg = (torch.randn(n, generator=g) * 3).to(DEV)
        mk = (torch.rand(n, generator=g) > 0.9).float().to(DEV)
        lg[:3] = torch.tensor([0.0, 40.0, -40.0], device=DEV)[: min(3, n)]     # p = 0.5, 1.0, ~0
        curve.update(lg, mk)
        all_l.append(lg); all_m.append(mk)
    lg, mk = torch.cat(all_l), torch.cat(all_m)
    probs = torch.sigmoid(lg)
    p0, r0, t0 = M.binned_pr_curve(probs, mk, 500)
    p1, r1, t1 = curve.compute()
    assert torch.equal(t0, t1) and torch.equal(p0, p1) and torch.equal(r0, r1)
    ref_loss = torch.nn.functional.binary_cross_entropy_with_logits(lg, mk)
    assert abs(curve.bce_loss().item() - ref_loss.item()) < 1e-5
    for thr in (0.0, 0.31, 0.5, 0.77, 1.0):
        t = torch.round(torch.tensor(thr), decimals=2).to(DEV)
        want = [float(v) for v in M.confusion_counts(probs > t, mk)]
        got = [float(v) for v in curve.counts_at(thr)]
        assert got == want, (thr, got, want)


def test_mixed_format_rejected_and_convert():
    x = nhwc(rnd(1, 64, 8, 8, seed=40), dt=FH)
    dy = nhwc(rnd(1, 64, 8, 8, seed=41), dt=BF)
    spec = ops.WeightSpec("conv3x3", 64, 64)
    with pytest.raises(RuntimeError, match="HPRI_ERR_ARG"):
        ops.igemm_wgrad(x, dy, 1, 64, spec.grad_buffer(DEV))
    xb = torch.empty_like(x, dtype=BF)
    ops.convert16(x, xb)
    assert torch.equal(xb, x.to(BF))


def test_ingest_exact():
    """Band slice / crop / flip / layout are pure indexing: exact up to the bf16 cast (dataset.py:266-270)."""
    src = torch.rand((2, 299, 20, 37), device=DEV)
    out = ops.hsi_ingest(src, 25, 263, c_pad=240, dtype=FH)
    assert torch.equal(out[..., :238], src[:, 25:263].permute(0, 2, 3, 1).to(FH))
    out = ops.hsi_ingest(src, 25, 263, c_pad=240, dtype=BF)
    ref = src[:, 25:263].permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(out[..., :238], ref) and (out[..., 238:] == 0).all()
    out = ops.hsi_ingest(src, 25, 263, crop=(3, 5, 12, 30), flip_w=True, c_pad=240, dtype=BF)
    ref = src[:, 25:263, 3:15, 5:35].flip(-1).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(out[..., :238], ref)
    mean = torch.rand(238, device=DEV)
    std = torch.rand(238, device=DEV) + 0.5
    out = ops.hsi_ingest(src * 255, 25, 263, flip_h=True, scale=1 / 255, band_mean=mean, band_std=std, c_pad=240,
                         dtype=BF)
    ref = ((src[:, 25:263] * 255 * (1 / 255) - mean[:, None, None]) / std[:, None, None]).flip(-2)
    assert (out[..., :238].float() - ref.permute(0, 2, 3, 1)).abs().max() < 2e-2
    assert ops.absmax(src * 255).item() == (src * 255).abs().max().item()
    # a cube the host already holds in fp16 (converted before the PCIe copy) gives bit-identical network input
    out16 = ops.hsi_ingest(src.to(FH), 25, 263, crop=(3, 5, 12, 30), flip_w=True, c_pad=240, dtype=FH)
    out32 = ops.hsi_ingest(src, 25, 263, crop=(3, 5, 12, 30), flip_w=True, c_pad=240, dtype=FH)
    assert torch.equal(out16, out32)
    rgb = torch.rand((2, 3, 9, 70), device=DEV)
    out = ops.hsi_ingest(rgb, 0, 3, c_pad=8, dtype=BF)
    assert torch.equal(out[..., :3], rgb.permute(0, 2, 3, 1).to(torch.bfloat16)) and (out[..., 3:] == 0).all()


@pytest.mark.parametrize("n,h,w,c,pool", [(2, 12, 17, 64, True), (1, 7, 9, 128, False), (2, 6, 10, 1024, True),
                                          (1, 1, 500, 1650, False)])
def test_bn_relu_fwd_bwd(n, h, w, c, pool):
    cp = (c + 63) // 64 * 64
    raw = torch.zeros((n, h, w, cp), dtype=torch.bfloat16, device=DEV)
    raw[..., :c] = nhwc(rnd(n, c, h, w, seed=20))
    gamma = (torch.rand(c, device=DEV) + 0.5)
    beta = (torch.rand(c, device=DEV) - 0.5) * 0.4
    cbias = torch.rand(c, device=DEV)
    rmean = torch.zeros(c, device=DEV)
    rvar = torch.ones(c, device=DEV)
    nbt = torch.zeros((), dtype=torch.long, device=DEV)
    xs = raw[..., :c].double().reshape(-1, c)
    stats = torch.stack([xs.sum(0), (xs * xs).sum(0)], 1).contiguous()
    scale = torch.zeros(cp, device=DEV); shift = torch.zeros(cp, device=DEV)
    smean = torch.zeros(cp, device=DEV); sinv = torch.zeros(cp, device=DEV)
    cnt = n * h * w
    ops.bn_finalize(stats, cnt, gamma, beta, cbias, rmean, rvar, nbt, True, scale, shift, smean, sinv, c)
    y = torch.full((n, h, w, cp), float("nan"), dtype=torch.bfloat16, device=DEV)
    pooled = torch.full((n, h // 2, w // 2, cp), float("nan"), dtype=torch.bfloat16, device=DEV) if pool else None
    ops.bn_relu_apply(raw, scale, shift, y, pooled, c=c)
    torch.cuda.synchronize()
    xr = nchw(raw, c).clone().requires_grad_(True)
    g = gamma.clone().requires_grad_(True)
    b = beta.clone().requires_grad_(True)
    bnr = torch.ones(c, device=DEV)
    bnm = torch.zeros(c, device=DEV)
    yr = F.relu(F.batch_norm(xr + cbias[None, :, None, None], bnm, bnr, g, b, True, 0.1, 1e-5))
    assert (nchw(y, c) - yr).abs().max() < 2e-2 * max(1.0, yr.abs().max().item())
    assert torch.allclose(rmean, bnm, rtol=1e-4, atol=1e-5) and torch.allclose(rvar, bnr, rtol=1e-4, atol=1e-5)
    assert nbt.item() == 1 and (stats == 0).all()
    if pool:
        pr = F.max_pool2d(nchw(y, c), 2)
        assert torch.equal(nchw(pooled, c), pr)
    # backward: dy direct (+ pooled gradient routed through the arg-max)
    dy = torch.zeros((n, h, w, cp), dtype=torch.bfloat16, device=DEV)
    dy[..., :c] = nhwc(rnd(n, c, h, w, seed=21))
    dpool = None
    if pool:
        dpool = torch.zeros((n, h // 2, w // 2, cp), dtype=torch.bfloat16, device=DEV)
        dpool[..., :c] = nhwc(rnd(n, c, h // 2, w // 2, seed=22))
    dx = torch.full((n, h, w, cp), float("nan"), dtype=torch.bfloat16, device=DEV)
    sums = torch.zeros((c, 3), dtype=torch.float64, device=DEV)
    dgamma = torch.zeros(c, device=DEV); dbeta = torch.zeros(c, device=DEV)
    ops.bn_relu_bwd(raw, scale, shift, smean, sinv, gamma, dx, sums, cnt, dy=dy, dpool=dpool, dgamma=dgamma,
                    dbeta=dbeta, c=c)
    torch.cuda.synchronize()
    loss = (yr * nchw(dy, c)).sum()
    if pool:
        loss = loss + (F.max_pool2d(yr, 2) * nchw(dpool, c)).sum()
    loss.backward()
    assert relerr(nchw(dx, c), xr.grad) < 1.5e-2
    assert relerr(dgamma, g.grad) < 2e-3 and relerr(dbeta, b.grad) < 2e-3


def test_head_and_bce():
    n, h, w, c = 2, 13, 21, 64
    raw = nhwc(rnd(n, c, h, w, seed=30))
    scale = torch.rand(c, device=DEV) + 0.5
    shift = torch.rand(c, device=DEV) - 0.5
    hw = rnd(c, scale=0.2, seed=31)
    hb = rnd(1, seed=32)
    logits = torch.empty((n, 1, h, w), dtype=torch.float32, device=DEV)
    ops.head_fwd(raw, scale, shift, hw, hb, logits)
    act = F.relu(nchw(raw) * scale[None, :, None, None] + shift[None, :, None, None])
    ref = (act * hw[None, :, None, None]).sum(1, keepdim=True) + hb
    assert (logits - ref).abs().max() < 1e-4 * max(1.0, ref.abs().max().item())
    target = (torch.rand((n, 1, h, w), device=DEV) > 0.7).float()
    loss_sum = torch.zeros((), dtype=torch.float64, device=DEV)
    dlogit = torch.empty_like(logits)
    counts = torch.zeros(4, dtype=torch.int64, device=DEV)
    ops.bce_fwd_bwd(logits, target, loss_sum, dlogit, counts)
    torch.cuda.synchronize()
    lr = logits.clone().requires_grad_(True)
    lref = F.binary_cross_entropy_with_logits(lr, target)
    lref.backward()
    assert abs(loss_sum.item() / logits.numel() - lref.item()) < 1e-6
    assert (dlogit - lr.grad).abs().max() < 1e-8
    seg = torch.sigmoid(logits) > 0.5
    tp = (seg & (target > 0.5)).sum().item()
    assert counts[0].item() == tp and counts.sum().item() == logits.numel()
    # head backward through bn_relu_bwd (dy = dlogit * w) incl. d(head weight)
    gamma = torch.ones(c, device=DEV)
    smean = torch.zeros(c, device=DEV)
    sinv = scale.clone()          # scale = gamma*invstd with gamma = 1, mean = 0 -> shift = beta
    dx = torch.empty_like(raw)
    sums = torch.zeros((c, 3), dtype=torch.float64, device=DEV)
    dhw = torch.zeros(c, device=DEV)
    ops.bn_relu_bwd(raw, scale, shift, smean, sinv, gamma, dx, sums, n * h * w, head_w=hw, dlogit=dlogit,
                    dhead_w=dhw)
    torch.cuda.synchronize()
    ref_dhw = (act * dlogit).sum((0, 2, 3))
    assert relerr(dhw, ref_dhw) < 1e-3


def test_mul16_strided_views():
    """hpri_mul16 on channel-sliced NHWC views (the two halves of a concat buffer) against torch."""
    torch.manual_seed(0)
    cat = (torch.randn((2, 9, 13, 128), device="cuda") * 2).half()
    out = torch.zeros((2, 9, 13, 128), device="cuda", dtype=torch.float16)
    ops.mul16(cat[..., :64], cat[..., 64:], out[..., 64:])
    want = (cat[..., :64].float() * cat[..., 64:].float()).half()
    assert torch.equal(out[..., 64:], want) and float(out[..., :64].abs().max()) == 0.0


@pytest.mark.parametrize("n,h,w,c,ph,pw", [(2, 5, 7, 64, 0, 0), (1, 9, 6, 24, 1, 1), (2, 1, 3, 8, 0, 1), (1, 38, 60, 128, 0, 1)])
def test_upsample2_bilinear_align_corners_fwd_bwd(n, h, w, c, ph, pw):
    """hpri_upsample2_fwd / _bwd against torch's Upsample(align_corners=True) + zero pad and its autograd, writing
    into / reading from the second half of a concat-shaped buffer."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    x = torch.randn((n, h, w, c), device="cuda").half()
    H2, W2 = 2 * h + ph, 2 * w + pw
    cat = torch.full((n, H2, W2, 2 * c), 7.0, device="cuda", dtype=torch.float16)
    ops.upsample2_fwd(x, cat[..., c:])
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.pad(F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True), [0, pw, 0, ph])
    got = cat[..., c:].float().permute(0, 3, 1, 2)
    assert float((got - yr.detach()).abs().max()) <= 2e-3 * float(yr.detach().abs().max()) + 1e-3
    assert float((cat[..., :c] - 7.0).abs().max()) == 0.0
    if ph:
        assert float(cat[:, 2 * h:, :, c:].abs().max()) == 0.0
    if pw:
        assert float(cat[:, :, 2 * w:, c:].abs().max()) == 0.0
    g = torch.randn((n, H2, W2, 2 * c), device="cuda").half()
    dx = torch.empty((n, h, w, c), device="cuda", dtype=torch.float16)
    ops.upsample2_bwd(g[..., c:], dx)
    yr.backward(g[..., c:].float().permute(0, 3, 1, 2))
    want = xr.grad.permute(0, 2, 3, 1)
    assert float((dx.float() - want).abs().max()) <= 3e-3 * float(want.abs().max()) + 1e-3


@pytest.mark.parametrize("c,dt", [(64, FH), (256, FH), (128, BF)])
def test_bn_contiguous_kernels_match_the_strided_ones(c, dt):
    """Dense tensors take the block-contiguous kernels (bn_*_contig_k); the same values embedded in a wider buffer
    (pixel stride != C) take the strided flat kernels.  Same arithmetic: forward bit-equal, backward equal up to the
    order of the fp64 reductions."""
    n, h, w = 2, 11, 19
    cnt = n * h * w
    raw_d = nhwc(rnd(n, c, h, w, seed=80), dt=dt)
    dy_d = nhwc(rnd(n, c, h, w, seed=81), dt=dt)
    wide = lambda t: torch.cat([t, torch.zeros_like(t)], -1)[..., :c]          # view with pix_stride = 2c
    raw_s, dy_s = wide(raw_d), wide(dy_d)
    assert raw_s.stride(2) == 2 * c and raw_d.stride(2) == c
    scale = torch.rand(c, device=DEV) + 0.5
    shift = torch.randn(c, device=DEV) * 0.3
    mean = torch.randn(c, device=DEV) * 0.1
    inv = torch.rand(c, device=DEV) + 0.5
    gamma = torch.rand(c, device=DEV) + 0.5
    y_d = torch.empty_like(raw_d)
    y_s = wide(torch.empty_like(raw_d))
    ops.bn_relu_apply(raw_d, scale, shift, y_d)
    ops.bn_relu_apply(raw_s, scale, shift, y_s)
    assert torch.equal(y_d, y_s)
    out = {}
    for tag, raw, dy in (("d", raw_d, dy_d), ("s", raw_s, dy_s)):
        dx = torch.empty_like(raw_d) if tag == "d" else wide(torch.empty_like(raw_d))
        sums = torch.zeros((c, 3), dtype=torch.float64, device=DEV)
        dg, db = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
        ops.bn_relu_bwd(raw, scale, shift, mean, inv, gamma, dx, sums, cnt, dy=dy, dgamma=dg, dbeta=db)
        out[tag] = (dx.float().clone(), sums.clone(), dg, db)
    torch.cuda.synchronize()
    # fp32 per-thread partials are summed in a different order by the two mappings before the fp64 accumulation
    assert torch.allclose(out["d"][1], out["s"][1], rtol=1e-4, atol=1e-4 * out["s"][1].abs().max().item())
    assert relerr(out["d"][0], out["s"][0]) < (2e-3 if dt == FH else 1.6e-2)      # at most an output ulp
    for i in (2, 3):                                   # dgamma / dbeta follow the sums
        assert torch.allclose(out["d"][i], out["s"][i], rtol=1e-4, atol=1e-4 * out["s"][i].abs().max().item())


def test_param_grads_are_unscaled_accumulated_and_flagged():
    """hpri_bn_relu_bwd_apply writes d{gamma,beta} = out_beta * old + out_scale * sum and raises the overflow flag on a
    non-finite value; hpri_colsum / hpri_sum_f32 / the unpack kernels take the same 1 / loss-scale factor."""
    n, h, w, c = 1, 9, 10, 64
    raw, dy = nhwc(rnd(n, c, h, w, seed=90), dt=FH), nhwc(rnd(n, c, h, w, seed=91), dt=FH)
    scale = torch.rand(c, device=DEV) + 0.5
    shift = torch.randn(c, device=DEV) * 0.3
    mean, inv, gamma = torch.zeros(c, device=DEV), torch.ones(c, device=DEV), torch.ones(c, device=DEV)
    dx = torch.empty_like(raw)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)

    def run(out_scale, out_beta, dg, db, dyv=dy):
        sums = torch.zeros((c, 3), dtype=torch.float64, device=DEV)
        ops.bn_relu_bwd(raw, scale, shift, mean, inv, gamma, dx, sums, n * h * w, dy=dyv, dgamma=dg, dbeta=db,
                        out_scale=out_scale, out_beta=out_beta, flag=flag)
    g1, b1 = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    run(1.0, 0.0, g1, b1)
    g2, b2 = torch.full((c,), 3.0, device=DEV), torch.full((c,), -2.0, device=DEV)
    run(0.25, 1.0, g2, b2)
    assert torch.allclose(g2, 3.0 + 0.25 * g1, rtol=1e-6, atol=1e-6) and torch.allclose(b2, -2.0 + 0.25 * b1, rtol=1e-6, atol=1e-6)
    assert flag.item() == 0
    bad = dy.clone()
    bad[0, 0, 0, 5] = float("inf")
    run(1.0, 0.0, g1, b1, bad)
    assert flag.item() == 1
    x = nhwc(rnd(2, 128, 6, 7, seed=92), dt=FH)
    o = torch.full((128,), 2.0, device=DEV)
    ops.colsum(x, o, beta=0.5, scale=0.125)
    assert torch.allclose(o, 1.0 + 0.125 * x.float().sum((0, 1, 2)), rtol=1e-5, atol=1e-5)
    v = torch.randn(10007, device=DEV)
    s = torch.empty(1, device=DEV)
    ops.sum_f32(v, s, scale=0.5)
    assert abs(s.item() - 0.5 * v.double().sum().item()) < 1e-3


@pytest.mark.parametrize("n,h,w,c,strided", [(2, 37, 51, 64, False), (1, 20, 33, 512, True), (2, 9, 13, 1024, False),
                                             (1, 64, 70, 128, True)])
def test_deterministic_column_and_scalar_sums(n, h, w, c, strided):
    """hpri_set_deterministic: hpri_colsum with one CTA per eight channels, hpri_sum_f32 on one CTA -- same values as the
    multi-CTA launches up to fp32 summation order, bit-identical from run to run; also on the channel-sliced, cropped
    views ConvTranspose's bias gradient is taken from."""
    full = nhwc(rnd(n, 2 * c if strided else c, h + 2, w + 1, seed=7), dt=FH)
    x = full[:, :h, :w, c:] if strided else full[:, :h, :w]
    ref = x.float().sum((0, 1, 2))
    v = torch.randn(123457, device=DEV)
    outs = []
    try:
        for det in (False, True, True):
            ops.set_deterministic(det)
            o = torch.full((c,), 3.0, device=DEV)
            ops.colsum(x, o, beta=1.0, scale=0.5)
            s1 = torch.empty(1, device=DEV)
            ops.sum_f32(v, s1, scale=2.0)
            torch.cuda.synchronize()
            outs.append((o, s1))
    finally:
        ops.set_deterministic(False)
    for o, s1 in outs:
        assert torch.allclose(o, 3.0 + 0.5 * ref, rtol=1e-4, atol=1e-3)
        assert abs(s1.item() - 2.0 * v.double().sum().item()) < 2e-2
    assert torch.equal(outs[1][0], outs[2][0]) and torch.equal(outs[1][1], outs[2][1])


def test_batch_tables_with_convT_jobs_scale_and_overflow_flag():
    """One table-driven launch packs 3x3 and ConvTranspose2d layers together (both operands each) and one unpacks /
    unscales their packed gradients; == the per-layer kernels.  A non-finite packed value raises the flag."""
    w3 = rnd(96, 40, 3, 3, seed=100)
    wt = rnd(128, 64, 2, 2, seed=101)                         # ConvTranspose2d weight [cin][cout][2][2]
    s3, st = ops.WeightSpec("conv3x3", 96, 40), ops.WeightSpec("convT2x2", 64, 128)
    f3 = torch.zeros((96, 9 * ops.kpad(40)), dtype=FH, device=DEV)
    d3 = torch.zeros((40, 9 * ops.kpad(96)), dtype=FH, device=DEV)
    ft = torch.zeros((4 * 64, ops.kpad(128)), dtype=FH, device=DEV)
    dtt = torch.zeros((128, 4 * ops.kpad(64)), dtype=FH, device=DEV)
    g3 = torch.randn((96, 9 * ops.kpad(40)), device=DEV); g3.view(96, 9, -1)[:, :, 40:] = 0
    gt = torch.randn((4 * 64, ops.kpad(128)), device=DEV)
    o3, ot = torch.empty_like(w3), torch.empty_like(wt)
    r3, rt = torch.empty_like(w3), torch.empty_like(wt)
    ops.unpack(g3.clone(), r3.view(-1), **s3.fwd)
    ops.unpack(gt.clone(), rt.view(-1), **st.fwd)
    jobs = [dict(w=w3.reshape(-1), fwd=f3, dgrad=d3, gpacked=g3, gdst=o3, cout=96, cin=40),
            dict(w=wt.reshape(-1), fwd=ft, dgrad=dtt, gpacked=gt, gdst=ot, cout=64, cin=128, kind=1)]
    table = ops.Conv3x3JobTable(jobs, DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.pack_conv3x3_batch(table)
    ops.unpack_conv3x3_batch(table, 0.5, flag)
    torch.cuda.synchronize()
    assert torch.equal(f3, s3.pack_fwd(w3.reshape(-1), dtype=FH)) and torch.equal(d3, s3.pack_dgrad(w3.reshape(-1), dtype=FH))
    assert torch.equal(ft, st.pack_fwd(wt.reshape(-1), dtype=FH)) and torch.equal(dtt, st.pack_dgrad(wt.reshape(-1), dtype=FH))
    assert torch.equal(o3, 0.5 * r3) and torch.equal(ot, 0.5 * rt) and flag.item() == 0
    assert not g3.any() and not gt.any()                      # zeroed behind the read
    gt[7, 3] = float("nan")
    ops.unpack_conv3x3_batch(table, 1.0, flag)
    assert flag.item() == 1


def test_fused_adam_skips_the_step_while_the_overflow_flag_is_raised():
    from hyperpri_b200.optim import FusedAdam
    p = torch.randn(5000, device=DEV).requires_grad_(True)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    opt = FusedAdam([p], lr=1e-2, found_inf=flag)
    before = p.detach().clone()
    p.grad = torch.full_like(p, float("inf"))
    flag.fill_(1)
    opt.step()
    assert torch.equal(p.detach(), before) and not opt.state[p]["exp_avg"].any()
    flag.zero_()
    p.grad = torch.randn_like(p)
    opt.step()
    assert not torch.equal(p.detach(), before) and torch.isfinite(p).all()
    assert opt.skipped_steps() == 1
