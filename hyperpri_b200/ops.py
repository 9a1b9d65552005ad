"""Tensor-level wrappers over the C ABI (include/hyperpri_b200.h).

Every function takes CUDA torch tensors, passes raw pointers / strides / the current stream to
the native library and returns immediately (stream-ordered).  torch is used for device memory
and streams only; no arithmetic on the hot path is done by torch here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import View, check

bf16 = torch.bfloat16
f16 = torch.float16
ACT = f16        # forward activations and forward weight operands (11-bit mantissa; BN keeps them O(1))
GRAD = f16       # gradients and dgrad weight operands; the loss gradient is scaled by a power of two (engine)
_DT = {bf16: 0, f16: 1}


PROFILE = None          # set to a list to record (kernel family, start event, end event) per native call
TENSOR_KERNELS = {"igemm_fwd", "convT_fwd", "convT_dgrad", "igemm_wgrad"}


def _timed(fn):
    name = fn.__name__

    def wrapper(*a, **k):
        if PROFILE is None:
            return fn(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        desc = " ".join("x".join(map(str, t.shape)) for t in list(a) + list(k.values())
                        if isinstance(t, torch.Tensor) and t.dim() >= 2)
        PROFILE.append((name, e0, e1, desc))
        return r

    wrapper.__name__, wrapper.__doc__ = name, fn.__doc__
    return wrapper


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def view(t: Optional[torch.Tensor], c: Optional[int] = None) -> Optional[View]:
    """hpri_view_t of an NHWC bf16 tensor (any strided slice whose channel stride is 1).
    `c` overrides the logical channel count (e.g. 1650 of a 1664-wide buffer)."""
    if t is None:
        return None
    assert t.dtype in _DT and t.dim() == 4 and t.is_cuda, (t.dtype, t.shape)
    assert t.stride(3) == 1
    n, h, w, cc = t.shape
    return View(t.data_ptr(), n, h, w, cc if c is None else c, t.stride(2), t.stride(1), t.stride(0), _DT[t.dtype])


def _vp(v: Optional[View]):
    return None if v is None else C.byref(v)


def kpad(c: int) -> int:
    return (c + 63) // 64 * 64


# ----------------------------------------------------------------------------- weights
@_timed
def pack(src: torch.Tensor, G, R, T, Cc, sg, sr, st, sc, flip=False, out: Optional[torch.Tensor] = None,
         src_offset: int = 0, dtype=None):
    """Generic fp32 parameter -> 16-bit operand [(G*R), T*kpad(C)] (see hpri_pack_weights)."""
    assert src.dtype == torch.float32 and src.is_cuda and src.is_contiguous()
    kc = kpad(Cc)
    if out is None:
        out = torch.empty((G * R, T * kc), dtype=dtype or ACT, device=src.device)
    assert out.is_contiguous() and out.shape == (G * R, T * kc)
    check(_lib.lib().hpri_pack_weights(C.c_void_p(src.data_ptr() + 4 * src_offset), _ptr(out), _DT[out.dtype], G, R, T,
                                       Cc, kc, sg, sr, st, sc, int(flip), _stream()), "hpri_pack_weights")
    return out


@_timed
def unpack(packed: torch.Tensor, dst: torch.Tensor, G, R, T, Cc, sg, sr, st, sc, flip=False, beta=0.0,
           zero_src=False, scale=1.0, flag=None):
    """Packed fp32 gradient -> torch layout: dst = beta * dst + scale * packed.  zero_src: reset `packed` to zero behind
    the read (the split-K weight-gradient kernel accumulates into it, so this replaces a memset before the next step).
    flag (int32 tensor): raised on a non-finite value."""
    assert packed.dtype == torch.float32 and dst.dtype == torch.float32 and dst.is_contiguous()
    check(_lib.lib().hpri_unpack_grads(_ptr(packed), _ptr(dst), G, R, T, Cc, kpad(Cc), sg, sr, st, sc, int(flip),
                                       float(beta), int(zero_src), float(scale), _ptr(flag), _stream()), "hpri_unpack_grads")
    return dst


@_timed
def pack_conv3x3(w, cout, cin, fwd=None, dgrad=None):
    """W[co][ci][3][3] fp32 -> fwd operand [co, 9*kpad(ci)] and/or dgrad operand [ci, 9*kpad(co)] (zero-initialised
    once by the caller; only valid entries are written)."""
    check(_lib.lib().hpri_pack_conv3x3(_ptr(w), cout, cin, _ptr(fwd), _DT[fwd.dtype] if fwd is not None else 0,
                                       _ptr(dgrad), _DT[dgrad.dtype] if dgrad is not None else 0, _stream()),
          "hpri_pack_conv3x3")


@_timed
def unpack_conv3x3(packed, cout, cin, dst):
    """Packed conv3x3 gradient -> W layout; always zeroes `packed` behind the read."""
    check(_lib.lib().hpri_unpack_conv3x3(_ptr(packed), cout, cin, _ptr(dst), _stream()), "hpri_unpack_conv3x3")


class Conv3x3JobTable:
    """Device table of hpri_conv3x3_job_t for the table-driven pack / unpack launches.  jobs: dicts with keys
    w, fwd, dgrad (or None), gpacked, gdst, cout, cin and optionally kind (0 conv3x3, 1 ConvTranspose2d k2 s2)."""

    def __init__(self, jobs, device):
        arr = (_lib.Conv3x3Job * len(jobs))()
        tiles = 0
        for i, j in enumerate(jobs):
            arr[i].w = j["w"].data_ptr()
            arr[i].dst_fwd = j["fwd"].data_ptr()
            arr[i].dst_dgrad = 0 if j.get("dgrad") is None else j["dgrad"].data_ptr()
            arr[i].grad_packed = j["gpacked"].data_ptr()
            arr[i].grad_dst = j["gdst"].data_ptr()
            arr[i].cout, arr[i].cin = j["cout"], j["cin"]
            arr[i].fwd_dtype = _DT[j["fwd"].dtype]
            arr[i].dgrad_dtype = 0 if j.get("dgrad") is None else _DT[j["dgrad"].dtype]
            arr[i].kind = int(j.get("kind", 0))
            arr[i].tile0 = tiles
            tiles += ((j["cin"] + 31) // 32) * ((j["cout"] + 31) // 32)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.dev = raw.to(device)
        self.n, self.tiles = len(jobs), tiles
        self.key = tuple(j["w"].data_ptr() for j in jobs)


@_timed
def pack_conv3x3_batch(table: Conv3x3JobTable):
    check(_lib.lib().hpri_pack_conv3x3_batch(_ptr(table.dev), table.n, table.tiles, _stream()), "hpri_pack_conv3x3_batch")


@_timed
def unpack_conv3x3_batch(table: Conv3x3JobTable, scale=1.0, flag=None):
    """gdst = scale * gpacked for every job of the table (packed buffers zeroed behind the read); flag: raised on a
    non-finite value."""
    check(_lib.lib().hpri_unpack_conv3x3_batch(_ptr(table.dev), table.n, table.tiles, float(scale), _ptr(flag), _stream()),
          "hpri_unpack_conv3x3_batch")


@_timed
def pack_convT(w, cin, cout, fwd):
    """W[ci][co][2][2] fp32 -> forward operand [4*cout, kpad(cin)] (zero-initialised once by the caller)."""
    check(_lib.lib().hpri_pack_convT2x2(_ptr(w), cin, cout, _ptr(fwd), _DT[fwd.dtype], _stream()), "hpri_pack_convT2x2")


@_timed
def unpack_convT(packed, cin, cout, dst):
    check(_lib.lib().hpri_unpack_convT2x2(_ptr(packed), cin, cout, _ptr(dst), _stream()), "hpri_unpack_convT2x2")


class WeightSpec:
    """Index maps between a torch-layout fp32 parameter and the packed GEMM operands."""

    def __init__(self, kind: str, cout: int, cin: int, split: int = 0):
        self.kind, self.cout, self.cin, self.split = kind, cout, cin, split
        if kind == "conv3x3":          # W[co, ci, 3, 3]
            self.fwd = dict(G=1, R=cout, T=9, Cc=cin, sg=0, sr=cin * 9, st=1, sc=9)
            self.dgr = dict(G=1, R=cin, T=9, Cc=cout, sg=0, sr=9, st=1, sc=cin * 9, flip=True)
            self.rows, self.taps = cout, 9
        elif kind == "linear":         # W[out, in]; split>0: input is cat of two `split`-wide halves,
            if split:                  # each padded to kpad(split) in the activation buffer
                assert cin == 2 * split
                self.fwd = dict(G=1, R=cout, T=2, Cc=split, sg=0, sr=cin, st=split, sc=1)
                self.dgr = None        # dgrad of a concat input is done per half
            else:
                self.fwd = dict(G=1, R=cout, T=1, Cc=cin, sg=0, sr=cin, st=0, sc=1)
                self.dgr = dict(G=1, R=cin, T=1, Cc=cout, sg=0, sr=1, st=0, sc=cin)
            self.rows, self.taps = cout, 1
        elif kind == "convT2x2":       # W[ci, co, 2, 2]; fwd rows (a*2+b)*co, K = ci
            self.fwd = dict(G=4, R=cout, T=1, Cc=cin, sg=1, sr=4, st=0, sc=cout * 4)
            self.dgr = dict(G=1, R=cin, T=4, Cc=cout, sg=0, sr=cout * 4, st=1, sc=4)
            self.rows, self.taps = 4 * cout, 1
        else:
            raise ValueError(kind)

    def pack_both(self, w, fwd, dgrad):
        """Refresh both operands of a conv3x3 in one tiled pass (buffers from zeros_like allocations)."""
        assert self.kind == "conv3x3"
        pack_conv3x3(w, self.cout, self.cin, fwd, dgrad)

    def pack_fwd(self, w, out=None, dtype=None):
        if self.kind == "convT2x2" and self.cin % 64 == 0:      # no padding columns: the tiled transpose covers all
            if out is None:
                out = torch.empty((4 * self.cout, kpad(self.cin)), dtype=dtype or ACT, device=w.device)
            pack_convT(w, self.cin, self.cout, out)
            return out
        return pack(w, out=out, dtype=dtype or ACT, **self.fwd)

    def pack_dgrad(self, w, out=None, dtype=None):
        return pack(w, out=out, dtype=dtype or GRAD, **self.dgr)

    def grad_buffer(self, device):
        f = self.fwd
        return torch.zeros((f["G"] * f["R"], f["T"] * kpad(f["Cc"])), dtype=torch.float32, device=device)

    def unpack_grad(self, packed, dst, beta=0.0, scale=1.0, flag=None):
        """Unpack AND reset `packed` to zero (ready for the next accumulation)."""
        if self.kind == "conv3x3" and beta == 0.0 and scale == 1.0 and flag is None:
            return unpack_conv3x3(packed, self.cout, self.cin, dst)
        if self.kind == "convT2x2" and beta == 0.0 and scale == 1.0 and flag is None:
            return unpack_convT(packed, self.cin, self.cout, dst)
        f = dict(self.fwd)
        return unpack(packed, dst, beta=beta, zero_src=True, scale=scale, flag=flag, **f)


# ----------------------------------------------------------------------------- contractions
@_timed
def igemm_fwd(x, wpack, rows, taps, y, n_store, bias=None, stats=None, block_n=0, x_c=None, y_c=None,
              accumulate=False, fin=None, bw=None, zero_sums=True):
    """fin: a BnFin from bn_fin(...) -> the launch also finalises train-mode BatchNorm (scale / shift / saved and
    running statistics) in its last CTA; no separate bn_finalize launch is needed.
    bw: (raw_below, scale, shift, save_mean, save_invstd, sums) of the BatchNorm+ReLU layer whose output gradient this
    3x3 dgrad launch produces -> its backward reduction (pass 1) is accumulated in the epilogue (sums zeroed here
    unless the caller has already cleared them: zero_sums=False)."""
    xv, yv = view(x, x_c), view(y, y_c)
    bws = None
    if bw is not None:
        raw, sc, sh, mu, iv, sums = bw
        if zero_sums:
            sums.zero_()
        rv = view(raw, y_c)
        bws = _lib.BnBwd(C.pointer(rv), sc.data_ptr(), sh.data_ptr(), mu.data_ptr(), iv.data_ptr(), sums.data_ptr())
    check(_lib.lib().hpri_igemm_fwd(_vp(xv), _ptr(wpack), _DT[wpack.dtype], rows, wpack.shape[1], taps, _vp(yv), n_store, _ptr(bias),
                                    _ptr(stats), int(accumulate), block_n, None if fin is None else C.byref(fin),
                                    None if bws is None else C.byref(bws), _stream()), "hpri_igemm_fwd")


def conv3x3_halo_ok(h, w, rows) -> bool:
    """True when a 3x3 launch of this shape runs on the halo-reuse kernel (needed by igemm_fwd(bw=...))."""
    return bool(_lib.lib().hpri_conv3x3_halo_ok(int(h), int(w), int(rows)))


DETERMINISTIC = __import__("os").environ.get("HPRI_DETERMINISTIC", "0") in ("1", "fwd")     # forward statistics
DETERMINISTIC_BWD = __import__("os").environ.get("HPRI_DETERMINISTIC", "0") == "1"           # and the backward pass


def set_deterministic(on: bool, backward: Optional[bool] = None):
    """Run-to-run reproducible training steps (the reference's Trainer(deterministic='warn'), PLTrainer.py:430,439,447).
    Forward: every CTA's BatchNorm partial sums go to its own slot (warps added in a fixed order) and the last CTA adds
    the slots in CTA order, instead of fp64 atomics in arrival order -- bit-identical logits.  Backward: the
    BatchNorm-backward reduction runs as its own kernels (fixed-order per-CTA partials whose fp32-valued addends are
    combined exactly in fp64) instead of in the dgrad epilogue, weight gradients are not split over pixel ranges
    (splits=1: one CTA sums each element), ConvTranspose / OutConv bias sums run on one CTA.  Slower (the full-resolution
    weight gradients lose most of their parallelism); same values up to summation order.
    backward=False keeps the production backward pass (HPRI_DETERMINISTIC=fwd): what the parity tests use, whose
    assertions are on the logits but which should exercise the kernels training runs."""
    global DETERMINISTIC, DETERMINISTIC_BWD
    DETERMINISTIC = bool(on)
    DETERMINISTIC_BWD = DETERMINISTIC if backward is None else bool(backward) and DETERMINISTIC
    check(_lib.lib().hpri_set_deterministic(int(DETERMINISTIC_BWD)), "hpri_set_deterministic")


def bn_fin(count, gamma, beta, conv_bias, rmean, rvar, nbt, scale, shift, smean, sinv, counter, momentum=0.1, eps=1e-5,
           partials=None):
    """hpri_bn_fin_t for igemm_fwd(fin=...): the arguments of bn_finalize(training=True) plus a zero-initialised
    uint32 ticket counter owned by the layer.  partials: fp32 scratch (>= 148 * C * 2) selecting the deterministic
    statistics path."""
    f = _lib.BnFin()
    f.partials = 0 if partials is None else partials.data_ptr()
    f.gamma, f.beta, f.conv_bias = gamma.data_ptr(), beta.data_ptr(), 0 if conv_bias is None else conv_bias.data_ptr()
    f.running_mean, f.running_var, f.num_batches_tracked = rmean.data_ptr(), rvar.data_ptr(), nbt.data_ptr()
    f.scale, f.shift, f.save_mean, f.save_invstd = scale.data_ptr(), shift.data_ptr(), smean.data_ptr(), sinv.data_ptr()
    f.counter, f.count, f.momentum, f.eps = counter.data_ptr(), int(count), float(momentum), float(eps)
    return f


def set_conv_algo(algo: int):
    """-1 heuristic, 0 generic per-tap kernel, 1 halo-reuse kernel on single CTAs, 2 halo-reuse kernel on CTA pairs
    (tests / benchmarks)."""
    check(_lib.lib().hpri_set_conv_algo(int(algo)), "hpri_set_conv_algo")


def set_wgrad_algo(algo: int):
    """-1 heuristic (halo-reuse weight-gradient kernel where it applies), 0 generic per-tap kernel only."""
    check(_lib.lib().hpri_set_wgrad_algo(int(algo)), "hpri_set_wgrad_algo")


def set_reverse_elementwise(on: bool):
    """Traversal order of the HBM-bound BatchNorm kernels: last chunk first, or ascending (default; no measured gain)."""
    check(_lib.lib().hpri_set_reverse_elementwise(int(bool(on))), "hpri_set_reverse_elementwise")


def set_halo_a_stages(stages: int):
    """Depth of the 3x3 halo kernel's halo-block ring: 2 (default) or 3."""
    check(_lib.lib().hpri_set_halo_a_stages(int(stages)), "hpri_set_halo_a_stages")


def set_sm_reserve(sms: int):
    """Keep `sms` SMs out of the persistent tensor-core grids (room for concurrently running NCCL CTAs)."""
    check(_lib.lib().hpri_set_sm_reserve(int(sms)), "hpri_set_sm_reserve")


@_timed
def convT_fwd(x, wpack, cout, y, bias=None, block_n=0):
    xv, yv = view(x), view(y)
    check(_lib.lib().hpri_convT2x2_fwd(_vp(xv), _ptr(wpack), _DT[wpack.dtype], cout, wpack.shape[1], _vp(yv), _ptr(bias), block_n,
                                       _stream()), "hpri_convT2x2_fwd")


@_timed
def convT_dgrad(dy, wpack, cin, dx, block_n=0):
    dv, xv = view(dy), view(dx)
    check(_lib.lib().hpri_convT2x2_dgrad(_vp(dv), _ptr(wpack), _DT[wpack.dtype], cin, wpack.shape[1], _vp(xv), block_n, _stream()),
          "hpri_convT2x2_dgrad")


@_timed
def igemm_wgrad(x, dy, mode, n_total, dw, block_n=0, splits=0, x_c=None, dy_c=None):
    xv, dv = view(x, x_c), view(dy, dy_c)
    assert dw.dtype == torch.float32 and dw.is_contiguous()
    check(_lib.lib().hpri_igemm_wgrad(_vp(xv), _vp(dv), mode, n_total, _ptr(dw), dw.shape[1], block_n, splits,
                                      _stream()), "hpri_igemm_wgrad")


# ----------------------------------------------------------------------------- ingest
@_timed
def hsi_ingest(src, lo, hi, crop=None, flip_h=False, flip_w=False, scale=1.0, band_mean=None, band_std=None,
               c_pad=None, out=None, dtype=None):
    """src fp32 (or host-pre-converted fp16) [N, bands, H, W] -> NHWC 16-bit [N, h, w, c_pad]."""
    assert src.dtype in (torch.float32, torch.float16) and src.is_contiguous() and src.dim() == 4
    n, bt, H, W = src.shape
    i0, j0, h, w = crop if crop is not None else (0, 0, H, W)
    nb = hi - lo
    c_pad = c_pad or (nb + 7) // 8 * 8
    if out is None:
        out = torch.empty((n, h, w, c_pad), dtype=dtype or ACT, device=src.device)
    fn = _lib.lib().hpri_hsi_ingest if src.dtype == torch.float32 else _lib.lib().hpri_hsi_ingest_f16
    check(fn(_ptr(src), n, bt, H, W, lo, hi, i0, j0, h, w, int(flip_h), int(flip_w),
                                     float(scale), _ptr(band_mean), _ptr(band_std), _ptr(out), _DT[out.dtype], c_pad,
                                     _stream()),
          "hpri_hsi_ingest")
    return out


@_timed
def convert16(x, y, c=None):
    xv, yv = view(x, c), view(y, c)
    check(_lib.lib().hpri_convert16(_vp(xv), _vp(yv), _stream()), "hpri_convert16")
    return y


@_timed
def mul16(a, b, y):
    """y = a * b over NHWC 16-bit views (Up with use_attention=True, model_parts.py:84-85)."""
    av, bv, yv = view(a), view(b), view(y)
    check(_lib.lib().hpri_mul16(_vp(av), _vp(bv), _vp(yv), _stream()), "hpri_mul16")
    return y


@_timed
def upsample2_fwd(x, y):
    """Bilinear x2, align_corners=True, into the top-left of y; the rest of y is zeroed (Up's padding)."""
    xv, yv = view(x), view(y)
    check(_lib.lib().hpri_upsample2_fwd(_vp(xv), _vp(yv), _stream()), "hpri_upsample2_fwd")
    return y


@_timed
def upsample2_bwd(dy, dx):
    dv, xv = view(dy), view(dx)
    check(_lib.lib().hpri_upsample2_bwd(_vp(dv), _vp(xv), _stream()), "hpri_upsample2_bwd")
    return dx


def absmax(src):
    out = torch.empty(1, dtype=torch.float32, device=src.device)
    check(_lib.lib().hpri_absmax(_ptr(src), src.numel(), _ptr(out), _stream()), "hpri_absmax")
    return out


# ----------------------------------------------------------------------------- BN / ReLU / pool
@_timed
def bn_finalize(stats, count, gamma, beta, conv_bias, rmean, rvar, nbt, training, scale, shift, smean, sinv,
                C_, momentum=0.1, eps=1e-5):
    check(_lib.lib().hpri_bn_finalize(_ptr(stats), count, _ptr(gamma), _ptr(beta), _ptr(conv_bias), _ptr(rmean),
                                      _ptr(rvar), _ptr(nbt), momentum, eps, int(training), _ptr(scale), _ptr(shift),
                                      _ptr(smean), _ptr(sinv), C_, _stream()), "hpri_bn_finalize")


@_timed
def bn_relu_apply(x, scale, shift, y, pooled=None, c=None):
    xv, yv, pv = view(x, c), view(y, c), view(pooled, c)
    check(_lib.lib().hpri_bn_relu_apply(_vp(xv), _ptr(scale), _ptr(shift), _vp(yv), _vp(pv), _stream()),
          "hpri_bn_relu_apply")


@_timed
def bn_relu_bwd(x, scale, shift, smean, sinv, gamma, dx, sums, count, dy=None, dpool=None, head_w=None,
                dlogit=None, dgamma=None, dbeta=None, dhead_w=None, c=None, reduced=False, out_scale=1.0, out_beta=0.0,
                flag=None):
    """reduced=True: `sums` was already accumulated by the dgrad launch that produced dy (igemm_fwd(bw=...)).
    dgamma / dbeta / dhead_w = out_beta * (old) + out_scale * (reduced sums); flag: raised on a non-finite value."""
    xv, dyv, dpv, dxv = view(x, c), view(dy, c), view(dpool, c), view(dx, c)
    L = _lib.lib()
    if not reduced:
        check(L.hpri_bn_relu_bwd_reduce(_vp(xv), _ptr(scale), _ptr(shift), _ptr(smean), _ptr(sinv), _vp(dyv), _vp(dpv),
                                        _ptr(head_w), _ptr(dlogit), _ptr(sums), _stream()), "hpri_bn_relu_bwd_reduce")
    check(L.hpri_bn_relu_bwd_apply(_vp(xv), _ptr(scale), _ptr(shift), _ptr(smean), _ptr(sinv), _ptr(gamma), _vp(dyv),
                                   _vp(dpv), _ptr(head_w), _ptr(dlogit), _ptr(sums), count, _vp(dxv), _ptr(dgamma),
                                   _ptr(dbeta), _ptr(dhead_w), float(out_scale), float(out_beta), _ptr(flag), _stream()),
          "hpri_bn_relu_bwd_apply")


@_timed
def bn_relu_bwd_reduce(x, scale, shift, smean, sinv, sums, dy=None, dpool=None, head_w=None, dlogit=None, c=None):
    """Pass 1 alone (sums[c] = {sum dz, sum dz * xhat, sum dlogit * act}, zeroed first): the pixel-parallel SpectralUNET
    all-reduces it over the ranks before bn_relu_bwd(..., reduced=True)."""
    xv, dyv, dpv = view(x, c), view(dy, c), view(dpool, c)
    check(_lib.lib().hpri_bn_relu_bwd_reduce(_vp(xv), _ptr(scale), _ptr(shift), _ptr(smean), _ptr(sinv), _vp(dyv), _vp(dpv),
                                             _ptr(head_w), _ptr(dlogit), _ptr(sums), _stream()), "hpri_bn_relu_bwd_reduce")


# ----------------------------------------------------------------------------- head / loss
@_timed
def head_fwd(x, scale, shift, w, b, logits, c=None):
    xv = view(x, c)
    check(_lib.lib().hpri_head_fwd(_vp(xv), _ptr(scale), _ptr(shift), _ptr(w), _ptr(b), _ptr(logits), _stream()),
          "hpri_head_fwd")


@_timed
def bce_fwd_bwd(logits, target, loss_sum, dlogit=None, counts=None, grad_scale=1.0, thr=0.5):
    assert logits.dtype == torch.float32 and target.dtype == torch.float32
    assert logits.is_contiguous() and target.is_contiguous()
    check(_lib.lib().hpri_bce_fwd_bwd(_ptr(logits), _ptr(target), logits.numel(), grad_scale, thr, _ptr(loss_sum),
                                      _ptr(dlogit), _ptr(counts), _stream()), "hpri_bce_fwd_bwd")


@_timed
def colsum(x, out, beta=0.0, c=None, scale=1.0):
    """out[c] = beta * out[c] + scale * sum over the pixels of x."""
    xv = view(x, c)
    check(_lib.lib().hpri_colsum(_vp(xv), _ptr(out), float(beta), float(scale), _stream()), "hpri_colsum")


@_timed
def scale_check(x, scale, flag):
    """x *= scale (fp32, in place); flag (int32 tensor) |= 1 on a non-finite result."""
    check(_lib.lib().hpri_scale_check(_ptr(x), x.numel(), float(scale), _ptr(flag), _stream()), "hpri_scale_check")


@_timed
def sum_f32(x, out, scale=1.0):
    check(_lib.lib().hpri_sum_f32(_ptr(x), x.numel(), _ptr(out), float(scale), _stream()), "hpri_sum_f32")


# ----------------------------------------------------------------------------- validation histograms
@_timed
def pr_hist(logits, target, thr, cut, hist_pos, hist_neg, cut_pos, cut_neg, bce_sum):
    """Accumulate the binned-PR-curve histograms of one batch (see hpri_pr_hist)."""
    assert logits.dtype == torch.float32 and target.dtype == torch.float32 and logits.is_contiguous()
    assert target.is_contiguous() and logits.numel() == target.numel()
    check(_lib.lib().hpri_pr_hist(_ptr(logits), _ptr(target), logits.numel(), _ptr(thr), thr.numel(), _ptr(cut),
                                  cut.numel(), _ptr(hist_pos), _ptr(hist_neg), _ptr(cut_pos), _ptr(cut_neg),
                                  _ptr(bce_sum), _stream()), "hpri_pr_hist")
