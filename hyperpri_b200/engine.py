"""Hand-scheduled forward / backward plans over the native kernels.

``UNetEngine`` runs UNet and CubeNET (reference models.py:23-68, 148-247 with
model_parts.py:14-99, including the `use_attention`, `bilinear` and `first_depth` variants through a
per-level channel plan); ``SpectralEngine`` runs SpectralUNET (models.py:71-145).  An engine
owns the NHWC 16-bit activation workspace for one input shape, packed 16-bit copies of the fp32
master parameters (refreshed when a parameter's version changes) and fp32 gradient buffers.
The nn.Modules in ``hyperpri_b200.src.Experiments.models`` call it through one
autograd.Function, so ``loss.backward()`` fills ``.grad`` on the module's own Parameters.

Dataflow decisions:
  * conv -> (raw fp16 + per-channel sum/sumsq in the GEMM epilogue, finalised by the launch's last CTA) ->
    bn_relu_apply writes the activation straight into its consumer's buffer (the first half of
    the decoder's concat buffer for skips, plus the 2x2-pooled tensor in the same pass);
  * ConvTranspose2d writes the second half of the concat buffer (pixel-shuffle epilogue); the
    zero-pad column/row of odd sizes is zeroed once at workspace creation;
  * the train-mode conv bias is cancelled by BatchNorm: it is skipped in the GEMM and re-added to
    running_mean; its gradient is identically zero and returned as zeros;
  * backward recomputes ReLU masks / pool arg-max from the saved raw conv outputs;
  * everything stored between kernels is fp16 (activations, weight operands, loss-scaled gradients), every
    accumulation fp32 / fp64; parameter gradients land unscaled in one flat fp32 arena ordered by backward completion
    (the data-parallel buckets), whose views are handed to the Parameters as their .grad;
  * weight gradients run on a second stream, the bucketed all-reduce hook behind them on that stream; the next batch's
    ingest may run on a third (set_next_input).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch

from . import ops
from .ops import ACT, GRAD, WeightSpec, kpad


def _z(shape, dev, dtype=ACT):
    return torch.zeros(shape, dtype=dtype, device=dev)


def _e(shape, dev, dtype=ACT):
    return torch.empty(shape, dtype=dtype, device=dev)


class _PackedParam:
    """16-bit GEMM operand(s) of one fp32 parameter, re-packed when the parameter changes."""

    def __init__(self, spec: WeightSpec, need_dgrad: bool):
        self.spec, self.need_dgrad = spec, need_dgrad
        self.key = None
        self.fwd = self.dgr = None
        self.gpack = None

    def alloc3x3(self, device):
        if self.fwd is None:                   # zero once: the tiled pack never writes padding
            f, g = self.spec.fwd, self.spec.dgr
            self.fwd = torch.zeros((f["R"], 9 * kpad(f["Cc"])), dtype=ACT, device=device)
            if self.need_dgrad:
                self.dgr = torch.zeros((g["R"], 9 * kpad(g["Cc"])), dtype=GRAD, device=device)

    def alloc_convT(self, device):
        if self.fwd is None:                   # zero once: the tiled pack never writes padding
            sp = self.spec
            self.fwd = torch.zeros((4 * sp.cout, kpad(sp.cin)), dtype=ACT, device=device)
            self.dgr = torch.zeros((sp.cin, 4 * kpad(sp.cout)), dtype=GRAD, device=device)

    def stale(self, w: torch.Tensor) -> bool:
        return (w.data_ptr(), w._version) != self.key

    def refresh(self, w: torch.Tensor):
        key = (w.data_ptr(), w._version)
        if key != self.key:
            wc = w.detach().reshape(-1)
            if self.spec.kind == "conv3x3":
                self.alloc3x3(w.device)
                self.spec.pack_both(wc, self.fwd, self.dgr)
            else:
                self.fwd = self.spec.pack_fwd(wc, out=self.fwd)
                if self.need_dgrad:
                    self.dgr = self.spec.pack_dgrad(wc, out=self.dgr)
            self.key = key


class _CBR:
    """conv3x3 (or Linear) -> BatchNorm -> ReLU unit: parameter names + per-layer BN state."""

    def __init__(self, conv: str, bn: str, cin: int, cout: int, dev, kind="conv3x3", need_dgrad=True, split=0):
        self.conv, self.bn, self.cin, self.cout = conv, bn, cin, cout
        self.pp = _PackedParam(WeightSpec(kind, cout, cin, split=split), need_dgrad and not split)
        cp = kpad(cout)
        self.stats = _z((cout, 2), dev, torch.float64)
        self.sums = _z((cout, 3), dev, torch.float64)
        self.scale, self.shift = _z((cp,), dev, torch.float32), _z((cp,), dev, torch.float32)
        self.smean, self.sinv = _z((cp,), dev, torch.float32), _z((cp,), dev, torch.float32)
        self.ticket = _z((1,), dev, torch.int32)        # last-CTA-done counter of the fused BatchNorm finalisation
        self.gw = self.pp.spec.grad_buffer(dev)


class _EngineBase:
    """Shared by both engines: loss scaling of the fp16 gradient path and arena finalisation."""
    bucket_hook = None

    def _init_scaling(self, device):
        # [0]: overflow flag of the current step's gradients (raised by the unpack / BatchNorm-backward kernels on a
        # non-finite value, read on the device by FusedAdam, cleared at the start of the next backward)
        if getattr(self, "overflow", None) is None:
            self.overflow = torch.zeros(1, dtype=torch.int32, device=device)
        self._S = 1.0
        self._gw_dirty = False
        self._init_side_stream(device)

    def _gw_buffers(self):
        raise NotImplementedError

    # ------------------------------------------------------------------ weight gradients on a second stream
    # A weight gradient only feeds the optimizer, so nothing later in backward waits for it.  Launched on a second
    # (lower-priority) stream it fills the SMs the persistent dgrad kernels leave idle in their last wave and runs
    # under the HBM-bound BatchNorm-backward kernels of the following layers (tensor pipe vs DRAM: different limits).
    # HPRI_WGRAD_STREAM=0 keeps everything on one stream.
    def _init_side_stream(self, device):
        self._side = None
        self._side_events = []
        self._side_used = 0
        if torch.device(device).type == "cuda" and os.environ.get("HPRI_WGRAD_STREAM", "1") != "0":
            self._side = torch.cuda.Stream(device=device, priority=0)

    def set_overlap(self, enabled: bool):
        """Turn the side stream off / back on (per-kernel timing needs serialised launches: an event pair around a
        kernel that shares the SMs with another stream's kernel also counts the time it spent waiting for them)."""
        self._no_prefetch = not enabled
        if not enabled:
            if self._side is not None:
                self._side_saved, self._side = self._side, None
        elif getattr(self, "_side_saved", None) is not None:
            self._side, self._side_saved = self._side_saved, None

    @staticmethod
    def _wgrad_splits():
        """0: the launcher picks the pixel-range split (partial products meet in fp32 red.add, in arrival order);
        deterministic mode: 1, every element of a weight gradient is summed by one CTA in a fixed order."""
        return 1 if ops.DETERMINISTIC_BWD else 0

    def _partials(self, cout):
        """Scratch of the deterministic-statistics mode (ops.set_deterministic / HPRI_DETERMINISTIC=1), else None."""
        if not ops.DETERMINISTIC:
            return None
        need = 148 * 2 * int(cout)
        buf = getattr(self, "_det_partials", None)
        if buf is None or buf.numel() < need:
            buf = torch.zeros(max(need, 148 * 2 * 2048), dtype=torch.float32, device=self.dev)
            self._det_partials = buf
        return buf

    def _event(self):
        if self._side_used == len(self._side_events):
            self._side_events.append(torch.cuda.Event())
        ev = self._side_events[self._side_used]
        self._side_used += 1
        return ev

    def _fork(self):
        """Event marking the current stream's work so far, for a later _on_side(fn, after=...)."""
        if self._side is None:
            return None
        ev = self._event()
        ev.record()
        return ev

    def _on_side(self, fn, after=None):
        """Run fn's launches on the side stream, ordered after `after` (default: everything launched so far on the
        current stream; False: only behind what the side stream already holds)."""
        if self._side is None:
            fn()
            return
        if after is not False:
            self._side.wait_event(after if after is not None else self._fork())
        with torch.cuda.stream(self._side):
            fn()

    def _side_mark(self):
        """Event marking the side stream's work so far (None without a side stream)."""
        if self._side is None:
            return None
        ev = self._event()
        ev.record(self._side)
        return ev

    def _join(self, ev):
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    # ------------------------------------------------------------------ ingest of the NEXT batch under this batch's backward
    # The HSI ingest (band slice + NCHW fp32/fp16 -> NHWC fp16, 1.7 GB of HBM traffic at 2 x 238 x 608 x 968) is
    # bandwidth-bound and has no dependence on the step before it, while backward is dominated by tensor-bound launches:
    # a caller that knows its next input registers it (set_next_input) and backward() ingests it into the second x
    # buffer on a third stream; the next forward(x) recognises the tensor and starts from the ready buffer.
    def set_next_input(self, x, ready_event=None):
        """x: the tensor the NEXT forward() will be given (None: nothing to prefetch); ready_event: CUDA event after which
        its contents are valid (e.g. the host->device copy of a DevicePrefetcher slot)."""
        self._next_input = None if x is None else (x, ready_event)

    def _launch_prefetch(self):
        nxt, self._next_input = getattr(self, "_next_input", None), None
        self._prefetched = None
        if nxt is None or not hasattr(self, "_ingest_into") or getattr(self, "_no_prefetch", False) or \
                os.environ.get("HPRI_INGEST_PREFETCH", "1") == "0":
            return
        x, ready = nxt
        if x is None or not x.is_cuda or self.ws is None or not self._same_input_shape(x):
            return
        if getattr(self, "_ingest_stream", None) is None:
            self._ingest_stream = torch.cuda.Stream(device=self.dev)
            self._ingest_event = torch.cuda.Event()
        ws = self.ws
        if "x_alt" not in ws:
            ws["x_alt"] = torch.empty_like(ws["x"])
        st = self._ingest_stream
        st.wait_stream(torch.cuda.current_stream())       # x_alt's last reader (the previous step's first-layer wgrad) is done
        if ready is not None:
            st.wait_event(ready)
        with torch.cuda.stream(st):
            self._ingest_into(x, ws["x_alt"])
            self._ingest_event.record(st)
        x.record_stream(st)
        self._prefetched = (x.data_ptr(), x._version, tuple(x.shape), x.dtype)

    def _take_prefetched(self, x) -> bool:
        """True when x is the registered next input: its ingested form becomes ws["x"]."""
        pf, self._prefetched = getattr(self, "_prefetched", None), None
        if pf is None or pf != (x.data_ptr(), x._version, tuple(x.shape), x.dtype) or "x_alt" not in (self.ws or {}):
            return False
        torch.cuda.current_stream().wait_event(self._ingest_event)
        self.ws["x"], self.ws["x_alt"] = self.ws["x_alt"], self.ws["x"]
        return True

    def invalidate_packed(self):
        """Force a re-pack of every 16-bit weight operand at the next forward (what an optimizer step causes)."""
        for pp in self._packed_params():
            pp.key = None
        self._wo_key = None

    def _begin_backward(self):
        """The packed weight-gradient buffers are accumulated into by split-K launches and reset to zero by the
        unpack kernels, so they are clean at the start of every backward unless the previous one was interrupted."""
        if self._gw_dirty:
            for g in self._gw_buffers():
                g.zero_()
        self._gw_dirty = True
        self._side_used = 0
        self._clear_overflow()

    def _end_backward(self):
        self._gw_dirty = False

    def _clear_overflow(self):
        self.overflow.zero_()

    def overflowed(self) -> bool:
        """True when the last backward produced a non-finite gradient (synchronises; FusedAdam needs no such read:
        it skips the step on the device).  A trainer may call this occasionally and lower `loss_scale_shift`."""
        return bool(self.overflow.item())

    loss_scale_shift = 0        # lowers the static loss scale by 2^shift (a trainer's reaction to overflowed())

    def loss_scale(self) -> float:
        """Power of two S with |S * dlogit| <= 2^-4 for a mean-reduced BCE (|dlogit| <= 1/numel): keeps the
        fp16 gradient tensors in the normal range with 2^20 headroom before overflow."""
        numel = self._logit_numel()
        return float(2.0 ** (math.ceil(math.log2(numel)) - 4 - self.loss_scale_shift))

    def _logit_numel(self):
        return self.ws["logits"].numel()

    def _scaled_dlogit(self, dlogit, prescaled):
        """prescaled: dlogit comes from loss_and_dlogit and carries loss_scale() (mean-reduced BCE: the bound on
        |dlogit| is exact).  Otherwise it is the gradient of an arbitrary criterion: the power-of-two scale is chosen
        from its measured magnitude (one host read of max|dlogit| -- only on this path, which the reference's
        configured BCEWithLogitsLoss does not take)."""
        dlogit = dlogit.contiguous().float()
        if prescaled:
            self._S = self.loss_scale()
            return dlogit
        amax = float(dlogit.abs().amax())
        if not math.isfinite(amax) or amax <= 0.0:
            self._S = self.loss_scale()
        else:
            self._S = float(2.0 ** max(-40, min(60, math.floor(math.log2(2.0 ** -4 / amax)) - self.loss_scale_shift)))
        return torch.mul(dlogit, self._S, out=self._dlogit_buffers()[1])

    def finalize_grads(self):
        """Kept for callers of the round-1 API: gradients now leave every kernel already unscaled (the unpack,
        BatchNorm-backward and bias-sum kernels take 1 / loss scale), so there is nothing left to do here."""
        return None

    def _dlogit_buffers(self):
        """(dlogit written by the fused BCE kernel, scratch of the same shape)."""
        return self.ws["dlogit"], self.ws["dlogit_s"]

    def loss_and_dlogit(self, logits, mask, grad_scale=1.0, thr=0.5):
        """Fused BCE forward + gradient.  The returned dlogit carries grad_scale * loss_scale():
        pass it to backward(..., prescaled=True)."""
        ws = self.ws
        dl, _ = self._dlogit_buffers()
        ops.bce_fwd_bwd(logits, mask.contiguous().float(), ws["loss_sum"], dl, ws["counts"],
                        grad_scale=grad_scale * self.loss_scale(), thr=thr)
        return ws["loss_sum"], dl, ws["counts"]

    def scaled_stored_dlogit(self, g):
        """The dlogit loss_and_dlogit stored, times the scalar device tensor g (the gradient flowing into the loss)."""
        dl, scratch = self._dlogit_buffers()
        return torch.mul(dl, g.to(dl.dtype), out=scratch)


class UNetEngine(_EngineBase):
    CH = [64, 128, 256, 512, 1024]

    def __init__(self, params: Dict[str, torch.Tensor], first: str, in_ch: int, device, attention: bool = False,
                 first_depth: int = 64, bilinear: bool = False):
        """params: name -> Parameter/buffer of the owning module (reference state-dict names).
        first: 'unet' (DoubleConv(in_ch, 64)) or 'cube' (Conv3d over in_ch bands, then inc2).
        first_depth: CubeNET's number of first-layer feature maps (models.py:168-178); when it is not 64 the last
        decoder block is `upsample4` / `upconv4` over cat([x1 (first_depth), up (64)]) (models.py:193-199, 229-240).
        bilinear: Up uses nn.Upsample(x2, bilinear, align_corners) instead of the ConvTranspose and DoubleConvs with
        mid channels (model_parts.py:56-61; models.py:33,43-49: down4 and the decoder outputs are halved)."""
        self.P, self.first, self.in_ch, self.dev = params, first, in_ch, device
        F = int(first_depth) if first == "cube" else 64
        if F % 8 or F <= 0:
            raise NotImplementedError("CubeNET first_depth must be a positive multiple of 8 (16-byte NHWC channel runs)")
        self.bilinear = bool(bilinear)
        if self.bilinear and F != 64:
            raise ValueError("bilinear=True with first_depth != 64 does not run in the reference either "
                             "(upconv4 expects 128 + first_depth channels, models.py:195-196)")
        # channel plan per level: CE encoder block output (= skip), U upsampled operand of the concat, M / D outputs of
        # the decoder block's first / second conv
        self.CE = [F] + self.CH[1:4] + [self.CH[4] // (2 if self.bilinear else 1)]
        self.U = list(self.CH[:4])
        self.M = list(self.CH[:4])
        self.D = [64, 64, 128, 256] if self.bilinear else list(self.CH[:4])
        self.up_name = {l: f"up{4 - l}.up" for l in range(4)}
        self.dec_name = {l: f"up{4 - l}.conv.double_conv" for l in range(4)}
        if F != 64:
            self.up_name[0], self.dec_name[0] = "upsample4", "upconv4.double_conv"
        # per decoder level: the reference's first_depth != 64 branch concatenates at the last block whatever
        # use_attention says (models.py:229-240)
        self.attl = [bool(attention) and not (l == 0 and F != 64) for l in range(4)]
        # use_attention=True (model_parts.py:84-85): the decoder block convolves skip * up (C channels) instead of
        # cat([skip, up]) (2C); the product and its two backward products are hpri_mul16 launches
        self.att = bool(attention)
        d = device
        CE, U, M, D = self.CE, self.U, self.M, self.D
        if first == "unet":
            self.cin_pad = (in_ch + 7) // 8 * 8
            enc0 = [_CBR("inc.double_conv.0", "inc.double_conv.1", self.cin_pad, 64, d, need_dgrad=False),
                    _CBR("inc.double_conv.3", "inc.double_conv.4", 64, 64, d)]
            enc0[0].true_cin = in_ch
        else:
            self.cin_pad = (in_ch + 7) // 8 * 8
            enc0 = [_CBR("first_conv", "inc.1", self.cin_pad, F, d, need_dgrad=False),
                    _CBR("inc2.0", "inc2.1", F, F, d)]
            enc0[0].true_cin = in_ch
        self.enc: List[List[_CBR]] = [enc0]
        for i in range(1, 5):
            p = f"down{i}.maxpool_conv.1.double_conv"
            self.enc.append([_CBR(p + ".0", p + ".1", CE[i - 1], CE[i], d), _CBR(p + ".3", p + ".4", CE[i], CE[i], d)])
        self.dec: Dict[int, List[_CBR]] = {}
        self.up: Dict[int, _PackedParam] = {}
        self.up_gw: Dict[int, torch.Tensor] = {}
        for i in range(1, 5):
            lvl = 4 - i
            p = self.dec_name[lvl]
            self.dec[lvl] = [_CBR(p + ".0", p + ".1", U[lvl] if self.attl[lvl] else CE[lvl] + U[lvl], M[lvl], d),
                             _CBR(p + ".3", p + ".4", M[lvl], D[lvl], d)]
            if not self.bilinear:
                self.up[lvl] = _PackedParam(WeightSpec("convT2x2", U[lvl], D[lvl + 1] if lvl < 3 else CE[4]), True)
                self.up_gw[lvl] = self.up[lvl].spec.grad_buffer(d)
        # BN-backward sums of every layer in one buffer: the ones accumulated by dgrad epilogues are cleared by ONE fill
        layers = [L for grp in list(self.enc) + list(self.dec.values()) for L in grp]
        self._bw_sums = _z((sum(L.cout for L in layers) + 1, 3), d, torch.float64)
        self.overflow = self._bw_sums[-1].view(torch.int32)[:1]       # cleared by the same fill as the sums
        off = 0
        for L in layers:
            L.sums = self._bw_sums[off:off + L.cout]
            off += L.cout
        # the first conv's real input-channel count differs from the padded view: pack from the true layout
        self.enc[0][0].pp = _PackedParam(_first_spec(in_ch, self.cin_pad, F), False)
        self.enc[0][0].gw = self.enc[0][0].pp.spec.grad_buffer(d)
        self.ws = None
        self.ws_key = None
        self.training_fwd = False
        self._tables = {}
        self._build_grad_arena()
        self.bucket_hook = None      # callable(flat_slice) invoked as each gradient bucket is complete
        self._init_scaling(device)

    def _build_grad_arena(self):
        """One flat fp32 buffer for every parameter gradient, ordered by backward completion
        (head, decoder top-down, encoder bottom-up) so that contiguous buckets can be all-reduced
        while the rest of backward is still running."""
        order, buckets = [], []

        def cbr_names(L):
            return [L.bn + ".weight", L.bn + ".bias", L.conv + ".weight", L.conv + ".bias"]

        start = 0
        order += ["outc.conv.bias", "outc.conv.weight"]
        for l in (0, 1, 2, 3):
            a, b = self.dec[l]
            order += cbr_names(b) + cbr_names(a)
            if not self.bilinear:
                order += [self.up_name[l] + ".weight", self.up_name[l] + ".bias"]
            buckets.append(len(order))
        for l in (4, 3, 2, 1, 0):
            a, b = self.enc[l]
            order += cbr_names(b) + cbr_names(a)
            buckets.append(len(order))
        offs, total = {}, 0
        for nm in order:
            offs[nm] = total
            total += self.P[nm].numel()
        self.arena = torch.zeros(total, dtype=torch.float32, device=self.dev)
        self.grads = {nm: self.arena[offs[nm]: offs[nm] + self.P[nm].numel()].view(self.P[nm].shape) for nm in order}
        ends = [offs[order[i]] if i < len(order) else total for i in buckets]
        self.bucket_bounds = list(zip([0] + ends[:-1], ends))
        del start

    def _bucket_done(self, idx):
        """Every launch of gradient bucket idx has been issued.  With an all-reduce hook the bucket's unpack (packed fp32
        -> arena, unscaled) and the hook run ON THE SIDE STREAM, stream-ordered behind the bucket's weight gradients: the
        main stream (dgrads, BatchNorm backward) never waits for them -- the BatchNorm parameter gradients of the bucket
        are covered by the events the weight-gradient launches already waited for.  Without a hook ONE unpack launch
        follows the join at the end of backward."""
        if self.bucket_hook is not None:
            self._on_side(lambda idx=idx: self._unpack_and_hook(idx), after=False)
        elif idx == 8:
            self._join(self._side_mark())
            self._unpack_bucket(idx)

    def _unpack_and_hook(self, idx):
        self._unpack_bucket(idx)
        a, b = self.bucket_bounds[idx]
        if b > a:
            self.bucket_hook(self.arena[a:b])

    # ------------------------------------------------------------------ workspace
    def _workspace(self, n, h, w):
        key = (n, h, w)
        if self.ws_key == key:
            return self.ws
        d, CE, U, M, D = self.dev, self.CE, self.U, self.M, self.D
        H, W = [h], [w]
        for _ in range(4):
            H.append(H[-1] // 2)
            W.append(W[-1] // 2)
        if H[4] < 1 or W[4] < 1:
            raise ValueError(f"input {h}x{w} too small for four 2x2 poolings")
        ws = {"H": H, "W": W, "n": n}
        ws["x"] = _e((n, h, w, self.cin_pad), d)
        for l in range(5):
            ws[f"enc_raw_a{l}"] = _e((n, H[l], W[l], CE[l]), d)
            ws[f"enc_act_a{l}"] = _e((n, H[l], W[l], CE[l]), d)
            ws[f"enc_raw_b{l}"] = _e((n, H[l], W[l], CE[l]), d)
            if l < 4:
                ws[f"cat{l}"] = _z((n, H[l], W[l], CE[l] + U[l]), d)      # [skip | upsampled], pad stays zero
                ws[f"gcat{l}"] = _z((n, H[l], W[l], CE[l] + U[l]), d, GRAD)
                if self.attl[l]:   # cat / gcat hold [skip | up] and [d skip | d up]; the conv works on the products
                    ws[f"mul{l}"] = _e((n, H[l], W[l], U[l]), d)
                    ws[f"gmul{l}"] = _e((n, H[l], W[l], U[l]), d, GRAD)
                ws[f"pool{l + 1}"] = _e((n, H[l + 1], W[l + 1], CE[l]), d)
                ws[f"gpool{l + 1}"] = _e((n, H[l + 1], W[l + 1], CE[l]), d, GRAD)
                ws[f"dec_raw_a{l}"] = _e((n, H[l], W[l], M[l]), d)
                ws[f"dec_act_a{l}"] = _e((n, H[l], W[l], M[l]), d)
                ws[f"dec_raw_b{l}"] = _e((n, H[l], W[l], D[l]), d)
                if l > 0:
                    ws[f"dec_act_b{l}"] = _e((n, H[l], W[l], D[l]), d)
            else:
                ws["act_b4"] = _e((n, H[4], W[4], CE[4]), d)
            # gradient buffers of the level, shared between its decoder and encoder blocks (and between the roles below)
            # whenever the channel counts agree -- always, except first_depth != 64 (level 0) and bilinear=True
            pool = {}

            def gbuf(c, tag=""):
                if (c, tag) not in pool:
                    pool[(c, tag)] = _e((n, H[l], W[l], c), d, GRAD)
                return pool[(c, tag)]
            cd, cm = (D[l], M[l]) if l < 4 else (CE[4], CE[4])
            side = self._side is not None
            # grad wrt a raw conv output: one per conv of the level when the weight gradients run on the side stream
            # (the one of an earlier conv may still be reading its buffer when a later conv's is written), else shared
            ws[f"R{l}"] = gbuf(cd, "r")                           # second conv of the decoder block
            ws[f"Ra{l}"] = gbuf(cm, "ra" if side else "r")        # first conv of the decoder block
            ws[f"Re{l}"] = gbuf(CE[l], "re" if side else "r")     # second / first conv of the encoder block
            ws[f"Rae{l}"] = gbuf(CE[l], "rae" if side else "r")
            ws[f"A{l}"] = gbuf(cm, "act")                         # grad wrt the first conv's activation
            ws[f"G{l}"] = gbuf(cd, "act")                         # grad wrt the block's output (from the level above)
            ws[f"Ae{l}"] = gbuf(CE[l], "act")
        if GRAD != ACT:
            ws["cvt"] = _e((n * h * w * max(self.cin_pad, CE[0] + U[0]),), d, GRAD)
        ws["logits"] = _e((n, 1, h, w), d, torch.float32)
        ws["dlogit"] = _e((n, 1, h, w), d, torch.float32)
        ws["dlogit_s"] = _e((n, 1, h, w), d, torch.float32)
        ws["loss_sum"] = _z((), d, torch.float64)
        ws["counts"] = _z((4,), d, torch.int64)
        self.ws, self.ws_key = ws, key
        return ws

    # ------------------------------------------------------------------ helpers
    def _cbr_fwd(self, L: _CBR, x, raw, act, training, pooled=None, apply=True):
        P = self.P
        L.pp.refresh(P[L.conv + ".weight"])
        n, h, w, _ = raw.shape
        if training:     # statistics in the GEMM epilogue, finalised by the launch's last CTA
            fin = ops.bn_fin(n * h * w, P[L.bn + ".weight"], P[L.bn + ".bias"], P[L.conv + ".bias"],
                             P[L.bn + ".running_mean"], P[L.bn + ".running_var"], P[L.bn + ".num_batches_tracked"],
                             L.scale, L.shift, L.smean, L.sinv, L.ticket, partials=self._partials(L.cout))
            ops.igemm_fwd(x, L.pp.fwd, L.cout, 9, raw, L.cout, stats=L.stats, fin=fin)
        else:
            ops.igemm_fwd(x, L.pp.fwd, L.cout, 9, raw, L.cout)
            ops.bn_finalize(L.stats, n * h * w, P[L.bn + ".weight"], P[L.bn + ".bias"], P[L.conv + ".bias"],
                            P[L.bn + ".running_mean"], P[L.bn + ".running_var"], P[L.bn + ".num_batches_tracked"],
                            False, L.scale, L.shift, L.smean, L.sinv, L.cout)
        if apply:
            ops.bn_relu_apply(raw, L.scale, L.shift, act, pooled)

    def _grad(self, name, like=None):
        return self.grads[name]

    def _as_grad_dtype(self, x):
        """wgrad needs x and dy in one format (tcgen05 kind::f16 faults on mixed f16 x bf16 operands).
        With fp16 gradients this is the identity; a bf16-gradient build converts through a scratch buffer."""
        if x.dtype == GRAD:
            return x
        n, h, w, c = x.shape
        buf = self.ws["cvt"][: n * h * w * c].view(n, h, w, c)
        return ops.convert16(x, buf)

    def _cbr_bwd(self, L: _CBR, x_in, raw, R, count, dy=None, dpool=None, head_w=None, dlogit=None, dhead_w=None,
                 dx_out=None, below=None, reduced=False):
        """BN+ReLU backward -> R, wgrad, optional dgrad into dx_out.
        below = (layer, raw) of the BatchNorm+ReLU layer whose output gradient dx_out IS (no other contribution): its
        backward reduction is fused into this dgrad launch; that layer is then called with reduced=True."""
        P = self.P
        ops.bn_relu_bwd(raw, L.scale, L.shift, L.smean, L.sinv, P[L.bn + ".weight"], R, L.sums, count, dy=dy,
                        dpool=dpool, head_w=head_w, dlogit=dlogit, dgamma=self._grad(L.bn + ".weight", L.scale[:L.cout]),
                        dbeta=self._grad(L.bn + ".bias", L.scale[:L.cout]), dhead_w=dhead_w, reduced=reduced,
                        out_scale=1.0 / self._S, flag=self.overflow)
        xg = self._as_grad_dtype(x_in)
        ready = self._fork()                 # R is complete here
        # the dgrad is on the critical path of backward: it is launched first so that it, not the weight gradient, gets
        # the SMs; the weight gradient then fills in behind it
        if dx_out is not None:
            bw = None
            if below is not None:
                Lb, raw_b = below
                bw = (raw_b, Lb.scale, Lb.shift, Lb.smean, Lb.sinv, Lb.sums)
            ops.igemm_fwd(R, L.pp.dgr, L.cin, 9, dx_out, L.cin, bw=bw, zero_sums=False)
        # the packed gradient is unpacked into the arena per bucket (_unpack_bucket); the conv bias gradient is
        # identically zero under train-mode BN
        self._on_side(lambda: ops.igemm_wgrad(xg, R, 1, L.cout, L.gw, splits=self._wgrad_splits()), after=ready)

    def _fusable(self, l, enc=False):
        """The a-layer of level l gets its whole output gradient from the b-layer's dgrad launch; the reduction can
        ride in that launch when it runs on the halo kernel."""
        if ops.DETERMINISTIC_BWD:   # the fused epilogue adds its warps' partial sums in arrival order
            return False
        return ops.conv3x3_halo_ok(self.ws["H"][l], self.ws["W"][l], (self.CE if enc else self.M)[l])

    # ------------------------------------------------------------------ forward
    def _ingest_into(self, x: torch.Tensor, dst: torch.Tensor):
        if x.dim() == 5:                       # CubeNET: N x 1 x D x H x W  (reshape is a no-copy squeeze)
            x = x.reshape(x.shape[0], x.shape[2], x.shape[3], x.shape[4])
        x = x.contiguous()
        if x.dtype != torch.float16:          # fp16 cubes (converted by the data loader before H2D) are ingested as is
            x = x.float()
        ops.hsi_ingest(x, 0, x.shape[1], c_pad=self.cin_pad, out=dst)

    def ingest(self, x: torch.Tensor, ws):
        self._ingest_into(x, ws["x"])

    def _same_input_shape(self, x):
        return self.ws_key == (x.shape[0], x.shape[-2], x.shape[-1])

    def forward(self, x: torch.Tensor, training: bool) -> torch.Tensor:
        n = x.shape[0]
        h, w = x.shape[-2], x.shape[-1]
        same_ws = self.ws_key == (n, h, w)
        ws = self._workspace(n, h, w)
        if not (same_ws and self._take_prefetched(x)):
            self.ingest(x, ws)
        return self.forward_ingested(ws, training)

    # ------------------------------------------------------------------ table-driven weight pack / gradient unpack
    def _conv_layers(self):
        return [L for grp in list(self.enc) + [self.dec[l] for l in (3, 2, 1, 0)] for L in grp]

    def _jobs(self, layers, ups=()):
        """Job records of the table-driven pack / unpack launches: 3x3 layers and (kind 1) the ConvTranspose2d of the
        decoder levels in `ups`."""
        jobs = []
        for L in layers:
            w = self.P[L.conv + ".weight"]
            L.pp.alloc3x3(w.device)
            jobs.append(dict(w=w, fwd=L.pp.fwd, dgrad=L.pp.dgr, gpacked=L.gw, gdst=self.grads[L.conv + ".weight"],
                             cout=L.pp.spec.cout, cin=L.pp.spec.cin))
        for l in ups:
            up, wn = self.up[l], self.up_name[l] + ".weight"
            w = self.P[wn]
            up.alloc_convT(w.device)
            jobs.append(dict(w=w, fwd=up.fwd, dgrad=up.dgr, gpacked=self.up_gw[l], gdst=self.grads[wn],
                             cout=up.spec.cout, cin=up.spec.cin, kind=1))
        return jobs

    def _table(self, name, layers, ups=()):
        """Cached device job table for a fixed list of layers (rebuilt if a parameter's storage moved)."""
        key = tuple(self.P[L.conv + ".weight"].data_ptr() for L in layers) + \
            tuple(self.P[self.up_name[l] + ".weight"].data_ptr() for l in ups)
        t = self._tables.get(name)
        if t is None or t.key != key:
            t = ops.Conv3x3JobTable(self._jobs(layers, ups), self.dev)
            t.key = key
            self._tables[name] = t
        return t

    def _up_levels(self):
        return tuple(sorted(self.up.keys()))

    def _refresh_packed(self):
        """Re-pack stale 16-bit weight operands: one table-driven launch when every 3x3 layer is stale (the normal
        case after an optimizer step), per layer otherwise."""
        layers = self._conv_layers()
        P = self.P
        ups = self._up_levels()
        if all(L.pp.stale(P[L.conv + ".weight"]) for L in layers) and \
                all(self.up[l].stale(P[self.up_name[l] + ".weight"]) for l in ups):
            ops.pack_conv3x3_batch(self._table("all", layers, ups))
            for L in layers:
                w = P[L.conv + ".weight"]
                L.pp.key = (w.data_ptr(), w._version)
            for l in ups:
                w = P[self.up_name[l] + ".weight"]
                self.up[l].key = (w.data_ptr(), w._version)

    def _unpack_bucket(self, idx):
        """Gradients of the 3x3 layers of bucket idx: packed fp32 -> arena.  Without an all-reduce hook the nine
        buckets are unpacked by ONE launch at the end of backward."""
        inv = 1.0 / self._S
        if self.bucket_hook is not None:
            l = idx if idx < 4 else 8 - idx
            grp = self.dec[l] if idx < 4 else self.enc[l]
            ups = (l,) if (idx < 4 and l in self.up) else ()
            ops.unpack_conv3x3_batch(self._table(f"bucket{idx}", [grp[1], grp[0]], ups), inv, self.overflow)
        elif idx == 8:
            ops.unpack_conv3x3_batch(self._table("all", self._conv_layers(), self._up_levels()), inv, self.overflow)

    def forward_ingested(self, ws, training: bool) -> torch.Tensor:
        P, CE, U = self.P, self.CE, self.U
        self.training_fwd = training
        self._refresh_packed()
        cur = ws["x"]
        for l in range(5):
            a, b = self.enc[l]
            self._cbr_fwd(a, cur, ws[f"enc_raw_a{l}"], ws[f"enc_act_a{l}"], training)
            if l < 4:
                self._cbr_fwd(b, ws[f"enc_act_a{l}"], ws[f"enc_raw_b{l}"], ws[f"cat{l}"][..., :CE[l]], training,
                              pooled=ws[f"pool{l + 1}"])
                cur = ws[f"pool{l + 1}"]
            else:
                self._cbr_fwd(b, ws["enc_act_a4"], ws["enc_raw_b4"], ws["act_b4"], training)
                cur = ws["act_b4"]
        for l in (3, 2, 1, 0):
            if self.bilinear:
                ops.upsample2_fwd(cur, ws[f"cat{l}"][..., CE[l]:])
            else:
                up = self.up[l]
                up.refresh(P[self.up_name[l] + ".weight"])
                ops.convT_fwd(cur, up.fwd, U[l], ws[f"cat{l}"][..., CE[l]:], bias=P[self.up_name[l] + ".bias"])
            a, b = self.dec[l]
            if self.attl[l]:
                ops.mul16(ws[f"cat{l}"][..., :CE[l]], ws[f"cat{l}"][..., CE[l]:], ws[f"mul{l}"])
            self._cbr_fwd(a, ws[f"mul{l}"] if self.attl[l] else ws[f"cat{l}"], ws[f"dec_raw_a{l}"], ws[f"dec_act_a{l}"], training)
            if l > 0:
                self._cbr_fwd(b, ws[f"dec_act_a{l}"], ws[f"dec_raw_b{l}"], ws[f"dec_act_b{l}"], training)
                cur = ws[f"dec_act_b{l}"]
            else:
                self._cbr_fwd(b, ws["dec_act_a0"], ws["dec_raw_b0"], None, training, apply=False)
        last = self.dec[0][1]
        ops.head_fwd(ws["dec_raw_b0"], last.scale, last.shift, P["outc.conv.weight"].detach().reshape(-1),
                     P["outc.conv.bias"], ws["logits"])
        return ws["logits"]

    # ------------------------------------------------------------------ backward
    def backward(self, dlogit: torch.Tensor, prescaled: bool = False) -> Dict[str, torch.Tensor]:
        """dlogit: gradient of the loss wrt the logits.  prescaled=True: it already carries loss_scale()
        (the fused BCE kernel applies it for free).  Gradients come out of the kernels scaled; they are
        unscaled in the arena by finalize_grads() (called here, or by the all-reduce hook owner)."""
        if not self.training_fwd:
            raise NotImplementedError("backward through eval-mode BatchNorm is not on the hot path")
        ws, P, CE, U = self.ws, self.P, self.CE, self.U
        n, H, W = ws["n"], ws["H"], ws["W"]
        self._begin_backward()
        self._launch_prefetch()            # the next batch's ingest runs under this backward (set_next_input)
        self._bw_sums.zero_()              # one fill for the BN-backward sums the dgrad epilogues accumulate into
        dlogit = self._scaled_dlogit(dlogit, prescaled)
        cnt = [n * H[l] * W[l] for l in range(5)]
        head_w = P["outc.conv.weight"].detach().reshape(-1)
        ops.sum_f32(dlogit, self._grad("outc.conv.bias", P["outc.conv.bias"]), scale=1.0 / self._S)
        for l in (0, 1, 2, 3):
            a, b = self.dec[l]
            fuse = self._fusable(l)
            below = (a, ws[f"dec_raw_a{l}"]) if fuse else None
            if l == 0:
                dhw = self._grad("outc.conv.weight", P["outc.conv.weight"]).view(-1)
                self._cbr_bwd(b, ws["dec_act_a0"], ws["dec_raw_b0"], ws["R0"], cnt[0], head_w=head_w, dlogit=dlogit,
                              dhead_w=dhw, dx_out=ws["A0"], below=below)
            else:
                self._cbr_bwd(b, ws[f"dec_act_a{l}"], ws[f"dec_raw_b{l}"], ws[f"R{l}"], cnt[l], dy=ws[f"G{l}"],
                              dx_out=ws[f"A{l}"], below=below)
            if self.attl[l]:
                self._cbr_bwd(a, ws[f"mul{l}"], ws[f"dec_raw_a{l}"], ws[f"Ra{l}"], cnt[l], dy=ws[f"A{l}"],
                              dx_out=ws[f"gmul{l}"], reduced=fuse)
                # d skip = g * up, d up = g * skip, written where the concat path keeps them
                ops.mul16(ws[f"gmul{l}"], ws[f"cat{l}"][..., CE[l]:], ws[f"gcat{l}"][..., :CE[l]])
                ops.mul16(ws[f"gmul{l}"], ws[f"cat{l}"][..., :CE[l]], ws[f"gcat{l}"][..., CE[l]:])
            else:
                self._cbr_bwd(a, ws[f"cat{l}"], ws[f"dec_raw_a{l}"], ws[f"Ra{l}"], cnt[l], dy=ws[f"A{l}"],
                              dx_out=ws[f"gcat{l}"], reduced=fuse)
            if self.bilinear:              # nn.Upsample backward: gather form over the (padded) destination gradient
                ops.upsample2_bwd(ws[f"gcat{l}"][..., CE[l]:], ws[f"G{l + 1}"])
                self._bucket_done(l)
                continue
            # ConvTranspose2d backward: its output is the second half of cat[l] over the 2h x 2w region
            up = self.up[l]
            dy_up = ws[f"gcat{l}"][:, :2 * H[l + 1], :2 * W[l + 1], CE[l]:]
            x_up = ws[f"dec_act_b{l + 1}"] if l < 3 else ws["act_b4"]
            ops.convT_dgrad(dy_up, up.dgr, up.spec.cin, ws[f"G{l + 1}"])
            gw = self.up_gw[l]
            wn = self.up_name[l] + ".weight"

            def up_wgrad(x_up=self._as_grad_dtype(x_up), dy_up=dy_up, gw=gw, l=l):
                # the packed gradient is unpacked with the bucket's 3x3 layers (_unpack_bucket)
                ops.igemm_wgrad(x_up, dy_up, 2, 4 * U[l], gw, splits=self._wgrad_splits())
                ops.colsum(dy_up, self._grad(self.up_name[l] + ".bias", P[self.up_name[l] + ".bias"]), scale=1.0 / self._S)
            self._on_side(up_wgrad)
            self._bucket_done(l)
        # encoder, deepest first
        for l in (4, 3, 2, 1, 0):
            a, b = self.enc[l]
            fuse = self._fusable(l, enc=True)
            below = (a, ws[f"enc_raw_a{l}"]) if fuse else None
            if l == 4:
                self._cbr_bwd(b, ws["enc_act_a4"], ws["enc_raw_b4"], ws["Re4"], cnt[4], dy=ws["G4"], dx_out=ws["Ae4"],
                              below=below)
            else:
                self._cbr_bwd(b, ws[f"enc_act_a{l}"], ws[f"enc_raw_b{l}"], ws[f"Re{l}"], cnt[l],
                              dy=ws[f"gcat{l}"][..., :CE[l]], dpool=ws[f"gpool{l + 1}"], dx_out=ws[f"Ae{l}"], below=below)
            x_in = ws[f"pool{l}"] if l > 0 else ws["x"]
            self._cbr_bwd(a, x_in, ws[f"enc_raw_a{l}"], ws[f"Rae{l}"], cnt[l], dy=ws[f"Ae{l}"],
                          dx_out=ws[f"gpool{l}"] if l > 0 else None, reduced=fuse)
            self._bucket_done(4 + (4 - l))
        if self.first == "cube":               # module registered twice (models.py:169-171): same tensor
            self.grads["inc.0.weight"] = self.grads["first_conv.weight"]
            self.grads["inc.0.bias"] = self.grads["first_conv.bias"]
        self._end_backward()
        return self.grads

    def _gw_buffers(self):
        return [L.gw for grp in list(self.enc) + list(self.dec.values()) for L in grp] + list(self.up_gw.values())

    def _clear_overflow(self):
        pass                                # backward()'s single fill of _bw_sums clears the flag too

    def _packed_params(self):
        return [L.pp for grp in list(self.enc) + list(self.dec.values()) for L in grp] + list(self.up.values())


def _first_spec(true_cin: int, cin_pad: int, cout: int = 64) -> WeightSpec:
    """First conv: the parameter has `true_cin` input channels, the activation view `cin_pad`."""
    s = WeightSpec("conv3x3", cout, true_cin)
    assert kpad(true_cin) == kpad(cin_pad)
    return s


# =======================================================================================
class SpectralEngine(_EngineBase):
    """SpectralUNET: nine Linear->BatchNorm1d->ReLU blocks on an (R*C) x D pixel matrix per image
    (models.py:105-115,132-144), concat by writing into halves of shared buffers."""
    BLOCKS = ["tail", "down1", "down2", "down3", "down4", "up1", "up2", "up3", "up4"]

    def __init__(self, params: Dict[str, torch.Tensor], hsi_depth: int, feats: int, device, bnorm: bool = True):
        """bnorm=False (models.py:72,105-110): every block is Linear -> ReLU.  The same kernels run it: the GEMM stores
        W x, the "BatchNorm apply" pass runs with scale = 1 and shift = the Linear bias, the backward apply pass with
        gamma = invstd = 1, mean = 0 and zero reduction sums is exactly the ReLU backward, and the (now non-zero) bias
        gradient is a column sum of its output."""
        self.P, self.D, self.F, self.dev = params, hsi_depth, feats, device
        self.bnorm = bool(bnorm)
        # N-tile of the forward / dgrad GEMMs: 256-column tcgen05.mma tiles are tensor-pipe bound (128 cycles per MMA
        # against 96 of operand fetch) while 128-column ones sit on the shared-memory operand bandwidth; 1650 features
        # are 6.45 such tiles (7 with the ragged last one), measured 15 % faster than 13 tiles of 128
        self.bn_tile = 256
        # weight gradients: 128-column tiles re-read X and dY from L2 at 128 B / cycle / SM (ncu: 9.1 GB of DRAM reads for
        # 2.8 GB of operands, tensor pipe 56 %); 256-column tiles need 94 B / cycle
        self.wgrad_tile = int(os.environ.get("HPRI_SPECTRAL_WGRAD_TILE", "256"))
        self.Fp = kpad(feats)
        self.Dp = (hsi_depth + 7) // 8 * 8
        d = device
        self.L: Dict[str, _CBR] = {}
        for nm in self.BLOCKS:
            cin = hsi_depth if nm == "tail" else (2 * feats if nm in ("up2", "up3", "up4") else feats)
            split = feats if nm in ("up2", "up3", "up4") else 0
            L = _CBR(nm + ".0", nm + ".1", cin, feats, d, kind="linear", need_dgrad=(nm != "tail"), split=split)
            if split:
                L.dgr_cat = _z((2 * self.Fp, kpad(feats)), d, GRAD)     # rows = cat-buffer feature index
            self.L[nm] = L
        self.ws = None
        self.ws_key = None
        self.training_fwd = False
        self.pp = None                      # hyperpri_b200.parallel.PixelParallel: this rank holds a row strip of every image
        self._init_scaling(device)
        per_block = (".0.weight", ".0.bias", ".1.weight", ".1.bias") if self.bnorm else (".0.weight", ".0.bias")
        names = [nm + k for nm in self.BLOCKS for k in per_block]
        names += ["outc.weight", "outc.bias"]
        if not self.bnorm:
            self._ones = torch.ones(self.Fp, dtype=torch.float32, device=d)
            self._zeros = _z((self.Fp,), d, torch.float32)
            for L in self.L.values():
                L.shift_nb = _z((self.Fp,), d, torch.float32)        # the Linear bias, padded: the "shift" of the apply pass
                L.bias_key = None
        offs, total = {}, 0
        for k in names:
            offs[k] = total
            total += params[k].numel()
        self.arena = torch.zeros(total, dtype=torch.float32, device=d)
        self.grads = {k: self.arena[offs[k]: offs[k] + params[k].numel()].view(params[k].shape) for k in names}
        self.bucket_bounds = [(0, total)]
        self.w_outc = _z((2 * self.Fp,), d, torch.float32)

    def _gw_buffers(self):
        return [L.gw for L in self.L.values()]

    def _packed_params(self):
        return [L.pp for L in self.L.values()]

    def set_pixel_parallel(self, pp):
        """Shard every image's pixels over pp's ranks (see PixelParallel); None restores the single-GPU plan."""
        if pp is not None and pp.world > 1 and not self.bnorm:
            raise NotImplementedError("pixel-parallel SpectralUNET is built for bnorm=True (the configured model)")
        self.pp = pp if (pp is not None and pp.world > 1) else None
        self.ws = self.ws_key = None

    def _dlogit_buffers(self):
        if self.pp is not None:
            return self.ws["dlogit_full"], self.ws["dlogit_full_s"]
        return self.ws["dlogit"], self.ws["dlogit_s"]

    def _logit_numel(self):
        return self.ws["n"] * self.ws["rows_full"] * self.ws["c"]

    def _workspace(self, n, r, c, R=None):
        """r: rows of this rank's strip; R: rows of the whole image (= r without pixel parallelism)."""
        R = r if R is None else R
        key = (n, r, c, R)
        if self.ws_key == key:
            return self.ws
        d, m, Fp = self.dev, r * c, self.Fp
        ws = {"n": n, "m": m, "r": r, "c": c, "rows_full": R, "m_glob": R * c, "img": []}
        if self.pp is not None:
            ws["logits_full"] = _e((n, 1, R, c), d, torch.float32)
            ws["dlogit_full"] = _e((n, 1, R, c), d, torch.float32)
            ws["dlogit_full_s"] = _e((n, 1, R, c), d, torch.float32)
        ws["x"] = _e((n, r, c, self.Dp), d)
        for _ in range(n):
            im = {nm: _e((1, 1, m, Fp), d) for nm in ("raw_" + b for b in self.BLOCKS)}
            for k in (1, 2, 3, 4):
                im[f"cat{k}"] = _z((1, 1, m, 2 * Fp), d)
            im["x4"] = _z((1, 1, m, Fp), d)
            im["bn"] = {b: [_z((Fp,), d, torch.float32) for _ in range(4)] for b in self.BLOCKS}
            ws["img"].append(im)
        for k in (2, 3, 4):
            ws[f"gcat{k}"] = _z((1, 1, m, 2 * Fp), d, GRAD)
        # gradient wrt a raw GEMM output; three rotate when the weight gradients run on the side stream (the one of
        # layer k may still be reading its buffer while layers k+1, k+2 are written)
        ws["R"] = [_z((1, 1, m, Fp), d, GRAD) for _ in range(3 if self._side is not None else 1)]
        if GRAD != ACT:
            ws["cvt"] = _z((1, 1, m, max(2 * Fp, self.Dp)), d, GRAD)
        ws["dlogit_s"] = _e((n, 1, r, c), d, torch.float32)
        ws["g4"] = _z((1, 1, m, Fp), d, GRAD)
        ws["gt"] = _z((1, 1, m, Fp), d, GRAD)
        ws["logits"] = _e((n, 1, r, c), d, torch.float32)
        ws["dlogit"] = _e((n, 1, r, c), d, torch.float32)
        ws["loss_sum"] = _z((), d, torch.float64)
        ws["counts"] = _z((4,), d, torch.int64)
        self.ws, self.ws_key = ws, key
        return ws

    def _refresh(self):
        P, F, Fp = self.P, self.F, self.Fp
        for nm, L in self.L.items():
            w = P[nm + ".0.weight"]
            key = (w.data_ptr(), w._version)
            if L.pp.key != key:
                L.pp.refresh(w)
                if L.pp.spec.split:
                    wc = w.detach().reshape(-1)
                    for half in (0, 1):   # dgrad operand rows follow the cat buffer: [0,F) and [Fp, Fp+F)
                        ops.pack(wc, G=1, R=F, T=1, Cc=F, sg=0, sr=1, st=0, sc=2 * F, src_offset=half * F,
                                 out=L.dgr_cat[half * Fp: half * Fp + F])
        if not self.bnorm:
            for nm, L in self.L.items():
                b = P[nm + ".0.bias"]
                key = (b.data_ptr(), b._version)
                if L.bias_key != key:
                    L.shift_nb[:F].copy_(b.detach())
                    L.bias_key = key
        wo_p = P["outc.weight"]
        wo_key = (wo_p.data_ptr(), wo_p._version)
        if getattr(self, "_wo_key", None) != wo_key:        # padded copy of the head weights, refreshed when they change
            wo = wo_p.detach().reshape(-1)
            self.w_outc.zero_()
            self.w_outc[:F].copy_(wo[:F])
            self.w_outc[Fp:Fp + F].copy_(wo[F:])
            self._wo_key = wo_key

    def _grad(self, name, like=None):
        return self.grads[name]

    def _io(self, im, nm):
        """(input view, logical in-features, output activation view) of block nm."""
        Fp = self.Fp
        src = {"down1": im["cat1"][..., :Fp], "down2": im["cat2"][..., :Fp], "down3": im["cat3"][..., :Fp],
               "down4": im["cat4"][..., :Fp], "up1": im["x4"], "up2": im["cat4"], "up3": im["cat3"],
               "up4": im["cat2"]}
        dst = {"tail": im["cat1"][..., :Fp], "down1": im["cat2"][..., :Fp], "down2": im["cat3"][..., :Fp],
               "down3": im["cat4"][..., :Fp], "down4": im["x4"], "up1": im["cat4"][..., Fp:],
               "up2": im["cat3"][..., Fp:], "up3": im["cat2"][..., Fp:], "up4": im["cat1"][..., Fp:]}
        return src.get(nm), dst[nm]

    def forward(self, x: torch.Tensor, training: bool) -> torch.Tensor:
        n, dch, r, c = x.shape
        x = x.contiguous()
        x = x if x.dtype == torch.float16 else x.float()
        if self.pp is not None:             # every rank receives the whole batch and ingests its row strip of each image
            r0, r1 = self.pp.rows(r)
            ws = self._workspace(n, r1 - r0, c, R=r)
            ws["r0"] = r0
            ops.hsi_ingest(x, 0, dch, crop=(r0, 0, r1 - r0, c), c_pad=self.Dp, out=ws["x"])
        else:
            ws = self._workspace(n, r, c)
            ops.hsi_ingest(x, 0, dch, c_pad=self.Dp, out=ws["x"])
        return self.forward_ingested(ws, training)

    def forward_ingested(self, ws, training: bool) -> torch.Tensor:
        P, F, Fp, m = self.P, self.F, self.Fp, ws["m"]
        pp, m_glob = self.pp, ws["m_glob"]
        self.training_fwd = training
        self._refresh()
        for i in range(ws["n"]):
            im = ws["img"][i]
            xin = ws["x"][i].reshape(1, 1, m, self.Dp)
            for nm in self.BLOCKS:
                L = self.L[nm]
                src, dst = self._io(im, nm)
                if nm == "tail":
                    src = xin
                raw = im["raw_" + nm]
                scale, shift, smean, sinv = im["bn"][nm]
                if not self.bnorm:
                    ops.igemm_fwd(src, L.pp.fwd, F, 1, raw, Fp, x_c=(self.D if nm == "tail" else None),
                                  block_n=self.bn_tile)
                    ops.bn_relu_apply(raw, self._ones, L.shift_nb, dst, None, c=F)
                    continue
                if training and pp is not None:
                    # this rank's partial (sum, sum of squares) -> exact per-image statistics over all strips
                    ops.igemm_fwd(src, L.pp.fwd, F, 1, raw, Fp, stats=L.stats, x_c=(self.D if nm == "tail" else None),
                                  block_n=self.bn_tile)
                    pp.all_reduce_(L.stats)
                    ops.bn_finalize(L.stats, m_glob, P[nm + ".1.weight"], P[nm + ".1.bias"], P[nm + ".0.bias"],
                                    P[nm + ".1.running_mean"], P[nm + ".1.running_var"],
                                    P[nm + ".1.num_batches_tracked"], True, scale, shift, smean, sinv, F)
                elif training:
                    fin = ops.bn_fin(m, P[nm + ".1.weight"], P[nm + ".1.bias"], P[nm + ".0.bias"],
                                     P[nm + ".1.running_mean"], P[nm + ".1.running_var"],
                                     P[nm + ".1.num_batches_tracked"], scale, shift, smean, sinv, L.ticket,
                                     partials=self._partials(F))
                    ops.igemm_fwd(src, L.pp.fwd, F, 1, raw, Fp, stats=L.stats, x_c=(self.D if nm == "tail" else None),
                                  block_n=self.bn_tile, fin=fin)
                else:
                    ops.igemm_fwd(src, L.pp.fwd, F, 1, raw, Fp, x_c=(self.D if nm == "tail" else None),
                                  block_n=self.bn_tile)
                    ops.bn_finalize(L.stats, m, P[nm + ".1.weight"], P[nm + ".1.bias"], P[nm + ".0.bias"],
                                    P[nm + ".1.running_mean"], P[nm + ".1.running_var"],
                                    P[nm + ".1.num_batches_tracked"], False, scale, shift, smean, sinv, F)
                ops.bn_relu_apply(raw, scale, shift, dst, None, c=F)
            ops.head_fwd(im["cat1"], None, None, self.w_outc, P["outc.bias"], ws["logits"][i])
        if pp is not None:
            return pp.gather_rows(ws["logits"], ws["rows_full"], out=ws["logits_full"])
        return ws["logits"]

    def backward(self, dlogit: torch.Tensor, prescaled: bool = False) -> Dict[str, torch.Tensor]:
        if not self.training_fwd:
            raise NotImplementedError("backward through eval-mode BatchNorm is not on the hot path")
        ws, P, F, Fp, m = self.ws, self.P, self.F, self.Fp, self.ws["m"]
        self._begin_backward()
        dlogit = self._scaled_dlogit(dlogit, prescaled)
        inv = 1.0 / self._S
        pp, m_glob = self.pp, ws["m_glob"]
        # pixel parallel: weight and OutConv-bias gradients are partial sums over this rank's pixels (summed over ranks by
        # the arena all-reduce below); gradients made from the all-reduced BatchNorm sums are the same on every rank, so
        # they are divided by the world size first
        inv_bn = inv / pp.world if pp is not None else inv
        if pp is not None:
            r0 = ws["r0"]
            dlogit = torch.mul(dlogit[:, :, r0:r0 + ws["r"]], 1.0, out=ws["dlogit_s"])     # this rank's rows, contiguous
        ops.sum_f32(dlogit, self._grad("outc.bias", P["outc.bias"]), scale=inv)
        go = self._grad("outc.weight", P["outc.weight"]).view(-1)      # [x0 features | up4 features] (models.py:143)
        r_turn, r_busy = 0, {}
        for i in range(ws["n"]):
            im = ws["img"][i]
            dl = dlogit[i].reshape(-1)
            xin = ws["x"][i].reshape(1, 1, m, self.Dp)
            acc_beta = 0.0 if i == 0 else 1.0          # parameter gradients accumulate over the per-image launches
            # (block, dy source, dgrad destination, accumulate?)
            plan = [("up4", None, ws["gcat2"], False), ("up3", ws["gcat2"][..., Fp:], ws["gcat3"], False),
                    ("up2", ws["gcat3"][..., Fp:], ws["gcat4"], False), ("up1", ws["gcat4"][..., Fp:], ws["g4"], False),
                    ("down4", ws["g4"], ws["gcat4"][..., :Fp], True), ("down3", ws["gcat4"][..., :Fp], ws["gcat3"][..., :Fp], True),
                    ("down2", ws["gcat3"][..., :Fp], ws["gcat2"][..., :Fp], True), ("down1", ws["gcat2"][..., :Fp], ws["gt"], False),
                    ("tail", ws["gt"], None, False)]
            for nm, dy, dx_dst, acc in plan:
                L = self.L[nm]
                src, _ = self._io(im, nm)
                if nm == "tail":
                    src = xin
                scale, shift, smean, sinv = im["bn"][nm]
                head = nm in ("up4", "tail")
                hw = dhw = None
                if head:                               # the OutConv reads cat(x0, up4 output): its two halves
                    off = Fp if nm == "up4" else 0
                    hw = self.w_outc[off:off + Fp]
                    dhw = go[F:] if nm == "up4" else go[:F]
                k = r_turn % len(ws["R"])
                r_turn += 1
                R = ws["R"][k]
                self._join(r_busy.get(k))          # the weight gradient that last read this buffer has finished
                if not self.bnorm:
                    if head:                       # d(outc weight) half = sum_p dlogit * act: third column of the reduction
                        ops.bn_relu_bwd_reduce(im["raw_" + nm], self._ones, L.shift_nb, self._zeros, self._ones, L.sums,
                                               dy=dy, head_w=hw, dlogit=dl, c=F)
                        part = L.sums[:F, 2].float() * inv
                        dhw.copy_(part) if i == 0 else dhw.add_(part)
                        L.sums.zero_()             # the apply pass must see zero sums: plain ReLU backward
                    ops.bn_relu_bwd(im["raw_" + nm], self._ones, L.shift_nb, self._zeros, self._ones, self._ones, R, L.sums,
                                    m_glob, dy=dy, head_w=hw, dlogit=dl if head else None, c=F, reduced=True)
                    ops.colsum(R, self._grad(nm + ".0.bias", P[nm + ".0.bias"]), beta=acc_beta, scale=inv, c=F)
                elif pp is not None:
                    ops.bn_relu_bwd_reduce(im["raw_" + nm], scale, shift, smean, sinv, L.sums, dy=dy, head_w=hw,
                                           dlogit=dl if head else None, c=F)
                    pp.all_reduce_(L.sums)
                if self.bnorm:
                    ops.bn_relu_bwd(im["raw_" + nm], scale, shift, smean, sinv, P[nm + ".1.weight"], R, L.sums, m_glob, dy=dy,
                                    head_w=hw, dlogit=dl if head else None,
                                    dgamma=self._grad(nm + ".1.weight", P[nm + ".1.weight"]),
                                    dbeta=self._grad(nm + ".1.bias", P[nm + ".1.bias"]), dhead_w=dhw, c=F,
                                    out_scale=inv_bn, out_beta=acc_beta, flag=self.overflow, reduced=pp is not None)
                xg = src if src.dtype == GRAD else ops.convert16(src, ws["cvt"][..., :src.shape[-1]])
                self._on_side(lambda xg=xg, R=R, L=L, nm=nm: ops.igemm_wgrad(
                    xg, R, 0, F, L.gw, x_c=(self.D if nm == "tail" else None), dy_c=F, block_n=self.wgrad_tile,
                    splits=self._wgrad_splits()))
                r_busy[k] = self._side_mark()
                if dx_dst is not None:
                    if L.pp.spec.split:
                        ops.igemm_fwd(R, L.dgr_cat, 2 * Fp, 1, dx_dst, 2 * Fp, x_c=F, accumulate=acc)
                    else:
                        ops.igemm_fwd(R, L.pp.dgr, F, 1, dx_dst, Fp, x_c=F, accumulate=acc, block_n=self.bn_tile)
        self._join(self._side_mark())
        for nm, L in self.L.items():
            wn = nm + ".0.weight"
            L.pp.spec.unpack_grad(L.gw, self._grad(wn, P[wn]).view(-1), scale=inv, flag=self.overflow)
        self._end_backward()
        if pp is not None:
            pp.all_reduce_(self.arena)
        elif self.bucket_hook is not None:
            self.bucket_hook(self.arena)
        return self.grads
