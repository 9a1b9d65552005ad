"""UNet, SpectralUNET, CubeNET and the model factories with the reference's public surface
(reference src/Experiments/models.py:23-292): same constructor signatures, attributes,
``state_dict`` keys/shapes and default torch initialisation, so existing checkpoints load and
``optim.Adam(model.parameters())`` keeps working.  ``forward`` hands the whole network to the
B200 engine through one autograd.Function; there is no torch-op fallback.
"""
import torch
import torch.nn as nn

from .model_parts import DoubleConv, Down, Up, OutConv
from ... import engine as _engine
from ... import ops as _lib_ops


def set_parameter_requires_grad(model, feature_extraction):
    if feature_extraction:
        for param in model.parameters():
            param.requires_grad = False


def _deliver_grads(net, eng, run_backward):
    """Run the engine backward and hand its gradients to the module's Parameters WITHOUT copies: the engine keeps every
    parameter gradient in one flat fp32 arena, and a Parameter whose ``.grad`` is None (what ``zero_grad()`` leaves by
    default) simply gets the arena view as its ``.grad``.  Returned to autograd: None for those, the view for
    Parameters that already hold a gradient of their own (autograd then accumulates, as usual).  A ``.grad`` that still
    aliases the arena from the previous step (``zero_grad(set_to_none=False)``, or deliberate gradient accumulation) is
    kept exact too: the arena is snapshotted before the engine overwrites it and added back afterwards.
    Consequences of the zero-copy hand-over: ``p.grad`` is overwritten in place by the next backward, and
    ``register_post_accumulate_grad_hook`` hooks do not fire for gradients delivered this way."""
    params = net._hot_params()
    aliased = any(p.grad is not None and p.grad.data_ptr() == eng.grads[name].data_ptr() for name, p in params
                  if name in eng.grads)
    prev = eng.arena.clone() if aliased else None
    grads = run_backward()
    if eng.bucket_hook is not None:
        # data-parallel: buckets were handed to the all-reduce as they completed; finish() waits for them and unscales
        owner = getattr(eng.bucket_hook, "__self__", None)
        if owner is not None:
            owner.finish()
    if prev is not None:
        eng.arena.add_(prev)
    out = []
    for name, p in params:
        g = grads[name] if p.requires_grad else None
        if g is None:
            out.append(None)
        elif p.grad is None:
            p.grad = g
            out.append(None)
        elif p.grad.data_ptr() == g.data_ptr():
            out.append(None)
        else:
            out.append(g)
    return out


class _NetFn(torch.autograd.Function):
    """forward: engine forward (NHWC 16-bit workspace) -> fp32 logits; backward: engine backward -> gradients delivered
    by _deliver_grads."""

    @staticmethod
    def forward(ctx, net, x, *params):
        eng = net._get_engine(x.device)
        logits = eng.forward(x, net.training)
        ctx.net, ctx.dev = net, x.device
        return logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        net = ctx.net
        eng = net._get_engine(ctx.dev)
        return (None, None, *_deliver_grads(net, eng, lambda: eng.backward(dlogits)))


class _NetLossFn(torch.autograd.Function):
    """Network + mean-reduced BCEWithLogitsLoss in one node (the reference's step body, PLTrainer.py:83-91): the fused
    head / loss kernel produces the loss, the loss-scaled logit gradient and the TP/FP/FN/TN counts at `thr` in one
    pass over the logits; backward starts from that stored gradient.  Outputs: (loss, logits, counts); logits and
    counts are not differentiable (the reference only uses them detached)."""

    @staticmethod
    def forward(ctx, net, x, target, thr, grad_scale, *params):
        eng = net._get_engine(x.device)
        logits = eng.forward(x, True)
        loss_sum, _, counts = eng.loss_and_dlogit(logits, target, grad_scale=grad_scale, thr=thr)
        ctx.net, ctx.dev = net, x.device
        ctx.set_materialize_grads(False)
        loss = (loss_sum / logits.numel()).float()
        lg, cn = logits.clone(), counts.clone()
        ctx.mark_non_differentiable(lg, cn)
        return loss, lg, cn

    @staticmethod
    def backward(ctx, gloss, _glogits=None, _gcounts=None):
        net = ctx.net
        eng = net._get_engine(ctx.dev)
        if gloss is None:
            raise RuntimeError("the fused network + BCE node was differentiated without a loss gradient")
        dl = eng.scaled_stored_dlogit(gloss)       # 4.7 MB elementwise; the stored dlogit carries the loss scale
        grads = _deliver_grads(net, eng, lambda: eng.backward(dl, prescaled=True))
        return (None, None, None, None, None, *grads)


class _EngineNet(nn.Module):
    """Shared plumbing: lazily built engine keyed by device, parameter/buffer name table."""

    def _hot_params(self):
        return list(self.named_parameters())

    def _tensor_table(self):
        table = {}
        for k, v in self.named_parameters(remove_duplicate=False):
            table[k] = v
        for k, v in self.named_buffers(remove_duplicate=False):
            table[k] = v
        return table

    def _make_engine(self, device):
        raise NotImplementedError

    def _get_engine(self, device):
        eng = self.__dict__.get("_eng")
        if eng is None or eng.dev != device:
            if device.type != "cuda":
                raise RuntimeError("hyperpri_b200 models run on a CUDA (sm_100a) device only; got " + str(device))
            eng = self._make_engine(device)
            self.__dict__["_eng"] = eng
        eng.P = self._tensor_table()          # parameters may have been re-assigned (load_state_dict, .to())
        return eng

    def _run(self, x):
        if x.device != next(self.parameters()).device:
            raise RuntimeError("input and parameters are on different devices")
        params = [p for _, p in self._hot_params()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _NetFn.apply(self, x, *params)
        return self._get_engine(x.device).forward(x, self.training).clone()

    def bce_step(self, x, target, thr=0.5, grad_scale=1.0):
        """forward + mean BCEWithLogitsLoss (+ its gradient and the segmentation counts at `thr`) through the fused
        kernels: returns (loss, logits, counts[TP, FP, FN, TN]).  With autograd enabled in train mode, `loss.backward()`
        runs the engine backward from the stored logit gradient (scaled by `grad_scale`, e.g. 1 / world size)."""
        if x.device != next(self.parameters()).device:
            raise RuntimeError("input and parameters are on different devices")
        target = target.contiguous().float()
        params = [p for _, p in self._hot_params()]
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _NetLossFn.apply(self, x, target, float(thr), float(grad_scale), *params)
        eng = self._get_engine(x.device)
        logits = eng.forward(x, self.training)
        ws = eng.ws
        _lib_ops.bce_fwd_bwd(logits, target, ws["loss_sum"], None, ws["counts"], thr=float(thr))
        return (ws["loss_sum"] / logits.numel()).float(), logits.clone(), ws["counts"].clone()

    def set_next_input(self, x, ready_event=None):
        """Tell the engine which tensor the NEXT forward will receive (e.g. the batch a DevicePrefetcher has already
        staged): its ingest then runs under this step's backward instead of at the head of the next step."""
        dev = next(self.parameters()).device
        if dev.type == "cuda" and hasattr(self._get_engine(dev), "set_next_input"):
            self._get_engine(dev).set_next_input(x, ready_event)

    def _finish(self, logits):
        if getattr(self, "analyze", False):
            return (logits, logits, torch.sigmoid(logits))
        return logits


class UNet(_EngineNet):
    """models.py:23-68.  31,043,521 parameters for n_channels=3."""

    def __init__(self, n_channels, n_classes, bilinear=True, feature_extraction=False, use_attention=False,
                 analyze=False):
        super().__init__()
        self.n_channels, self.n_classes = n_channels, n_classes
        self.bilinear, self.use_attention, self.analyze = bilinear, use_attention, analyze
        if n_classes != 1:
            raise NotImplementedError("the B200 head kernel is built for n_classes=1 (every reference config)")
        w = [64 * 2 ** i for i in range(5)]
        factor = 2 if bilinear else 1                                   # models.py:33
        self.inc = DoubleConv(n_channels, w[0])
        self.down1, self.down2 = Down(w[0], w[1]), Down(w[1], w[2])
        self.down3, self.down4 = Down(w[2], w[3]), Down(w[3], w[4] // factor)
        self.up1 = Up(w[4], w[3], bilinear, use_attention=use_attention)
        self.up2 = Up(w[3], w[2], bilinear, use_attention=use_attention)
        self.up3 = Up(w[2], w[1], bilinear, use_attention=use_attention)
        self.up4 = Up(w[1], w[0] * factor, bilinear, use_attention=use_attention)
        self.outc = OutConv(w[0], n_classes)

    def _make_engine(self, device):
        return _engine.UNetEngine(self._tensor_table(), "unet", self.n_channels, device, attention=self.use_attention,
                                  bilinear=self.bilinear)

    def forward(self, x):
        return self._finish(self._run(x))


class CubeNET(_EngineNet):
    """models.py:148-247 (any first_depth that is a multiple of 8).  The Conv3d spanning all bands is executed as the 2-D
    3x3 conv over `hsi_depth` channels it equals; `first_conv` stays an nn.Conv3d so the
    (64,1,D,3,3) weight and the aliased `first_conv.*` / `inc.0.*` keys are preserved."""

    def __init__(self, hsi_depth, n_classes, first_depth=64, bilinear=True, use_attention=False, analyze=False):
        super().__init__()
        self.n_channels = 1
        self.depth, self.first_depth, self.n_classes = hsi_depth, first_depth, n_classes
        self.bilinear, self.use_attention, self.analyze = bilinear, use_attention, analyze
        if n_classes != 1:
            raise NotImplementedError("the B200 head kernel is built for n_classes=1 (every reference config)")
        self.first_conv = nn.Conv3d(1, first_depth, kernel_size=(hsi_depth, 3, 3), padding=(0, 1, 1))
        self.inc = nn.Sequential(self.first_conv, nn.BatchNorm3d(first_depth), nn.ReLU(inplace=True))
        self.inc2 = nn.Sequential(nn.Conv2d(first_depth, first_depth, kernel_size=3, padding=1),
                                  nn.BatchNorm2d(first_depth), nn.ReLU(inplace=True))
        c = 128
        self.down1, self.down2 = Down(first_depth, c), Down(c, 2 * c)
        factor = 2 if bilinear else 1                                   # models.py:166
        self.down3, self.down4 = Down(2 * c, 4 * c), Down(4 * c, 8 * c // factor)
        self.up1 = Up(8 * c, 4 * c, bilinear, use_attention=use_attention)
        self.up2 = Up(4 * c, 2 * c, bilinear, use_attention=use_attention)
        self.up3 = Up(2 * c, c, bilinear, use_attention=use_attention)
        if first_depth == 64:
            self.up4 = Up(c, 64 * factor, bilinear, use_attention=use_attention)
        else:       # models.py:193-199: the skip has first_depth channels, the up path 64; always concatenated (:229-240)
            if bilinear:
                raise ValueError("bilinear=True with first_depth != 64 does not run in the reference either: upconv4 expects "
                                 "128 + first_depth channels but receives 64 + first_depth (models.py:195-196, 229-240)")
            self.upsample4 = nn.ConvTranspose2d(c, 64, kernel_size=2, stride=2)
            self.upconv4 = DoubleConv(64 + first_depth, 64)
        self.outc = OutConv(64, n_classes)

    def _make_engine(self, device):
        return _engine.UNetEngine(self._tensor_table(), "cube", self.depth, device, attention=self.use_attention,
                                  first_depth=self.first_depth, bilinear=self.bilinear)

    def forward(self, x):
        """x: N x 1 x D x R x C (a depth mismatch is not raised by the reference either, models.py:211)."""
        return self._finish(self._run(x))


class SpectralUNET(_EngineNet):
    """models.py:71-145: per-pixel MLP U-Net, BatchNorm1d statistics per image."""

    def __init__(self, hsi_depth, n_classes, bn_feats=16, bnorm=True):
        super().__init__()
        self.hsi_depth = self.n_channels = hsi_depth
        self.n_classes = n_classes
        self.bnorm = bool(bnorm)
        if n_classes != 1:
            raise NotImplementedError("the B200 head kernel is built for n_classes=1 (every reference config)")
        f = bn_feats
        self.layer_feats = [f] * 5
        bn = self.bnorm
        self.tail = self._basic_module(hsi_depth, f, bn)
        self.down1, self.down2 = self._basic_module(f, f, bn), self._basic_module(f, f, bn)
        self.down3, self.down4 = self._basic_module(f, f, bn), self._basic_module(f, f, bn)
        self.up1 = self._basic_module(f, f, bn)
        self.up2, self.up3, self.up4 = (self._basic_module(2 * f, f, bn) for _ in range(3))
        self.outc = nn.Linear(2 * f, n_classes)

    def _basic_module(self, in_feats, out_feats, bn=True):
        if not bn:                                   # models.py:105-110
            return nn.Sequential(nn.Linear(in_feats, out_feats), nn.ReLU())
        return nn.Sequential(nn.Linear(in_feats, out_feats), nn.BatchNorm1d(out_feats), nn.ReLU())

    def _make_engine(self, device):
        eng = _engine.SpectralEngine(self._tensor_table(), self.hsi_depth, self.layer_feats[0], device, bnorm=self.bnorm)
        if self.__dict__.get("_pp") is not None:
            eng.set_pixel_parallel(self.__dict__["_pp"])
        return eng

    def enable_pixel_parallel(self, group=None):
        """The model-sharded option (`train_net(..., model_parallel=True)`; the reference: DeepSpeed ZeRO-2 over >= 2
        GPUs, PLTrainer.py:409-433): every rank of `group` keeps a row strip of every image and the per-image
        BatchNorm statistics are all-reduced per block (hyperpri_b200.parallel.PixelParallel).  Every rank must be
        given the same batch.  enable_pixel_parallel(False) returns to the single-GPU plan."""
        from ...parallel import PixelParallel
        pp = None if group is False else PixelParallel(group)
        self.__dict__["_pp"] = pp
        eng = self.__dict__.get("_eng")
        if eng is not None:
            eng.set_pixel_parallel(pp)
        return pp

    def forward(self, x):
        """x: N x D x R x C -> N x n_classes x R x C."""
        return self._run(x)


def initialize_model(model_name, num_classes, Network_parameters, analyze=False):
    """models.py:250-276."""
    if model_name == 'UNET':
        return UNet(Network_parameters['channels'], num_classes, bilinear=Network_parameters['bilinear'],
                    feature_extraction=Network_parameters['feature_extraction'],
                    use_attention=Network_parameters['use_attention'], analyze=analyze)
    if model_name == 'SpectralUNET':
        depth = Network_parameters['hsi_hi'] - Network_parameters['hsi_lo']
        return SpectralUNET(depth, num_classes, bn_feats=Network_parameters['spectral_bn_size'])
    if model_name == 'CubeNET':
        depth = Network_parameters['hsi_hi'] - Network_parameters['hsi_lo']
        return CubeNET(depth, num_classes, first_depth=Network_parameters['3d_featmaps'],
                       bilinear=Network_parameters['bilinear'], use_attention=Network_parameters['use_attention'],
                       analyze=analyze)
    raise RuntimeError('Invalid model')


def translate_load_dir(model_name, net_params):
    """models.py:279-292."""
    if model_name == 'SpectralUNET':
        return f"{model_name}_{net_params['spectral_bn_size']}"
    if model_name == 'CubeNET':
        return f"{model_name}_{net_params['3d_featmaps']}"
    return "UNET"
