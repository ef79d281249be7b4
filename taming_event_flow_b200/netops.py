"""Fused element-wise stages of the recurrent flow network's training step (``csrc/tef_net.cu``).

The convolutions stay on cuDNN (``north_star``); everything between them -- gate activations, state blending, bias and
bias-gradient sums, residual + ReLU, the flow heads' up-sampling and scaling -- is one CUDA kernel per direction instead of
the dozens of tiny ATen kernels autograd would issue.  Each operator is a ``torch.autograd.Function`` around the C ABI and
the ``aten.convolution`` / ``aten.convolution_backward`` primitives.  fp32, channels_last, CUDA only: ``usable()`` says
whether a tensor qualifies, and the modules of ``flownet.py`` fall back to their plain PyTorch formulation (the restatement
of upstream's module code, also the reference of ``tests/test_netops_gpu.py``) for anything else, e.g. under bf16 autocast.

Upstream: ``models/submodules.py:111-152`` (ConvGRU), ``:13-60,155-200`` (ConvLayer, ResidualBlock), ``models/model.py:65-85``
(flow heads), ``train_flow.py:106-108`` (flow scaling).
"""
import ctypes

import torch
from torch.autograd.function import once_differentiable

from . import _lib

CL = torch.channels_last
ACT = {None: 0, "none": 0, "relu": 1, "tanh": 2}


def usable(*tensors):
    """Fused kernels apply: CUDA, fp32, no autocast, channel counts that are multiples of four."""
    if torch.is_autocast_enabled("cuda"):
        return False
    for t in tensors:
        if t is None:
            continue
        if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.shape[1] % 4 == 0):
            return False
    return True


def usable_input(x):
    """A convolution input (any channel count: only cuDNN reads it)."""
    return x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and not torch.is_autocast_enabled("cuda")


def _cl(t):
    return t.contiguous(memory_format=CL)


def _rows(t):
    return t.shape[0] * t.shape[2] * t.shape[3]


def _conv(x, w, stride, padding):
    return _cl(torch.ops.aten.convolution(x, w, None, [stride, stride], [padding, padding], [1, 1], False, [0, 0], 1))


def _conv_bwd(g, x, w, stride, padding, need_x=True):
    gx, gw, _ = torch.ops.aten.convolution_backward(g, x, w, None, [stride, stride], [padding, padding], [1, 1], False, [0, 0], 1, [need_x, True, False])
    return gx, gw


def _conv_bwd_input(g, x, w, stride, padding):
    return torch.ops.aten.convolution_backward(g, x, w, None, [stride, stride], [padding, padding], [1, 1], False, [0, 0], 1, [True, False, False])[0]


def _conv_bwd_weight(g, x, w, stride, padding):
    return torch.ops.aten.convolution_backward(g, x, w, None, [stride, stride], [padding, padding], [1, 1], False, [0, 0], 1, [False, True, False])[1]


class WindowStacks:
    """Per-layer storage of one loss window for the DEFERRED weight gradient (back-propagation through time).

    A recurrent step re-uses every filter in each of the P passes of a loss window, and autograd computes one small weight
    gradient per pass and sums them (P launches of a skinny GEMM whose reduction dimension is one pass's B*H*W pixels, P - 1
    additions of the full filter bank).  The weight gradient is linear in (input, output gradient) pairs, so the fused
    operators keep the convolution inputs and the output gradients of all passes side by side -- `x[name]` and `g[name]` are
    [P*B, C, H, W] channels_last stacks, pass t in rows t*B..(t+1)*B, written in place by the kernels that produce them -- and
    the backward of the window's FIRST pass (the last to run) computes ONE weight gradient over all of them.
    Requires the loss to back-propagate through the window's first pass, which is what upstream's training loop does
    (``train_flow.py:106-137``: every pass feeds the loss); `begin(P)` opts a window in, a layer called more often than P times
    falls back to the per-pass gradient.  Passes whose backward never ran (their output did not reach the loss) contribute zero:
    their rows of the gradient stacks are cleared before the window's gradient is formed.  If backward ran for later passes but
    never for the first one, the window's weight gradient was never emitted; the next `begin` raises instead of training on.
    Memory: the convolution inputs of a window are alive until its backward pass under autograd anyway; the stacks add the output
    gradients of all passes (freed pass by pass otherwise) and stay allocated between windows -- about 1.5 GB for the 31 M-parameter
    network at B = 8, P = 10, 128 x 128, proportional to B * P * H * W."""

    def __init__(self):
        self.P, self.slot, self.used, self.key, self.x, self.g = 0, 0, 0, None, {}, {}
        self.ran = set()          # slots whose backward has run in this window

    def begin(self, P):
        if self.ran and 0 not in self.ran:
            ran, self.ran = sorted(self.ran), set()
            raise RuntimeError("deferred weight gradient lost: the previous loss window back-propagated through passes %s of a layer but not "
                               "through its first pass, whose backward emits the window's weight gradient (netops.WindowStacks); "
                               "open the window with begin_window(0) for losses that skip the first pass" % ran)
        self.P, self.slot, self.used = int(P), 0, 0
        self.ran = set()

    def window_rows(self, slot, B):
        """Called by the backward of pass `slot`: records it; for the first pass (the one that emits the window's weight gradient)
        clears the gradient rows of passes whose backward never ran and returns the number of rows to reduce over, else None."""
        self.ran.add(slot)
        if slot != 0:
            return None
        for t in range(self.used):
            if t not in self.ran:
                for g in self.g.values():
                    self.rows(g, t, B).zero_()
        return self.used * B

    def take_slot(self):
        """Slot of this forward call, or None when the window is not open / already full / gradients are off."""
        if self.P <= 0 or self.slot >= self.P or not torch.is_grad_enabled():
            return None
        t = self.slot
        self.slot += 1
        self.used = max(self.used, self.slot)
        return t

    def ensure(self, key, shapes_x, shapes_g, device):
        """(Re)allocate the stacks: shapes_* = {name: (B, C, H, W)} of one pass."""
        key = (self.P, key)
        if key != self.key:
            mk = lambda shp: torch.empty((self.P * shp[0],) + tuple(shp[1:]), dtype=torch.float32, device=device, memory_format=CL)
            self.x = {k: mk(v) for k, v in shapes_x.items()}
            self.g = {k: mk(v) for k, v in shapes_g.items()}
            self.key = key

    @staticmethod
    def rows(stack, t, B):
        return stack[t * B:(t + 1) * B]


class _FusedConvGRU(torch.autograd.Function):
    """new_state = ConvGRU(x, h) with the update and reset gates merged into one convolution (w_zr = [w_update; w_reset]).
    With `stacks` / `slot` (WindowStacks) the weight gradients of the window are computed once, by the first pass's backward."""

    @staticmethod
    def forward(ctx, x, h, w_zr, b_zr, w_c, b_c, stacks, slot):
        L = _lib.lib()
        st = _lib.stream()
        B, Cx, H, W = x.shape
        C = h.shape[1]
        pad = w_zr.shape[2] // 2
        M = B * H * W
        if slot is not None:
            stacks.ensure(("gru", B, Cx, C, H, W), {"xh": (B, Cx + C, H, W), "xrh": (B, Cx + C, H, W)}, {"zr": (B, 2 * C, H, W), "c": (B, C, H, W)}, x.device)
            xh, xrh = stacks.rows(stacks.x["xh"], slot, B), stacks.rows(stacks.x["xrh"], slot, B)
            torch.cat([x, h], 1, out=xh)
        else:
            xh = _cl(torch.cat([_cl(x), _cl(h)], 1))
            xrh = torch.empty_like(xh)
        zr = _conv(xh, w_zr, 1, pad)
        _lib.check(L.tef_gru_gates(_lib.ptr(zr), _lib.ptr(b_zr), _lib.ptr(xh), _lib.ptr(xrh), ctypes.c_long(M), Cx, C, st), "tef_gru_gates")
        c = _conv(xrh, w_c, 1, pad)
        out = torch.empty_like(c)
        _lib.check(L.tef_gru_output(_lib.ptr(c), _lib.ptr(b_c), _lib.ptr(xh), _lib.ptr(zr), _lib.ptr(out), ctypes.c_long(M), Cx, C, st), "tef_gru_output")
        # plain attributes, not save_for_backward: the stack slots are views of a buffer the other passes write to
        ctx.t = (xh, xrh, zr, c, w_zr, w_c)
        ctx.dims = (M, B, Cx, C, pad, b_zr is not None, b_c is not None, stacks, slot)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        if ctx.t is None:
            raise RuntimeError("a fused network operator can be back-propagated once (its activations are released by the first backward pass)")
        xh, xrh, zr, cand, w_zr, w_c = ctx.t
        ctx.t = None
        M, B, Cx, C, pad, has_bzr, has_bc, stacks, slot = ctx.dims
        L = _lib.lib()
        st = _lib.stream()
        gout = _cl(gout)
        if slot is not None:
            gc, gzr = stacks.rows(stacks.g["c"], slot, B), stacks.rows(stacks.g["zr"], slot, B)
        else:
            gc, gzr = torch.empty_like(cand), torch.empty_like(zr)
        gh = torch.empty_like(cand)
        gb = torch.zeros(3 * C, dtype=torch.float32, device=gout.device)
        gb_zr, gb_c = gb[:2 * C], gb[2 * C:]
        _lib.check(L.tef_gru_output_bwd(_lib.ptr(gout), _lib.ptr(cand), _lib.ptr(xh), _lib.ptr(zr), _lib.ptr(gc), _lib.ptr(gzr), _lib.ptr(gh),
                                        _lib.ptr(gb_c) if has_bc else None, _lib.ptr(gb_zr) if has_bzr else None, ctypes.c_long(M), Cx, C, st),
                   "tef_gru_output_bwd")
        gw_c = gw_zr = None
        if slot is None:
            gxrh, gw_c = _conv_bwd(gc, xrh, w_c, 1, pad)
        else:
            gxrh = _conv_bwd_input(gc, xrh, w_c, 1, pad)
        gxrh = _cl(gxrh)
        _lib.check(L.tef_gru_gates_bwd(_lib.ptr(gxrh), _lib.ptr(xh), _lib.ptr(zr), _lib.ptr(gzr), _lib.ptr(gh), _lib.ptr(gb_zr) if has_bzr else None,
                                       ctypes.c_long(M), Cx, C, st), "tef_gru_gates_bwd")
        if slot is None:
            gxh, gw_zr = _conv_bwd(gzr, xh, w_zr, 1, pad)
        else:
            gxh = _conv_bwd_input(gzr, xh, w_zr, 1, pad)
        gxh = _cl(gxh)
        gx = torch.empty((B, Cx, xh.shape[2], xh.shape[3]), dtype=torch.float32, device=gout.device, memory_format=CL)
        _lib.check(L.tef_gru_input_grads(_lib.ptr(gxrh), _lib.ptr(gxh), _lib.ptr(gx), _lib.ptr(gh), ctypes.c_long(M), Cx, C, st), "tef_gru_input_grads")
        n = stacks.window_rows(slot, B) if slot is not None else None
        if n is not None:                                       # the window's first pass: every pair of the window is in the stacks
            gw_c = _conv_bwd_weight(stacks.g["c"][:n], stacks.x["xrh"][:n], w_c, 1, pad)
            gw_zr = _conv_bwd_weight(stacks.g["zr"][:n], stacks.x["xh"][:n], w_zr, 1, pad)
        return gx, gh, gw_zr, (gb_zr if has_bzr else None), gw_c, (gb_c if has_bc else None), None, None


def conv_gru(x, h, w_zr, b_zr, w_c, b_c, stacks=None):
    """ConvGRU cell (upstream ``models/submodules.py:134-152``): z, r = sigmoid(conv_zr([x, h])); cand = tanh(conv_c([x, h * r]));
    returns h * (1 - z) + cand * z.  `w_zr` stacks the update gate's filters over the reset gate's.  `stacks`: the layer's
    WindowStacks when the weight gradients of a loss window are to be computed once (see there)."""
    _lib.require_cuda(x, h, w_zr, w_c)
    slot = stacks.take_slot() if stacks is not None else None
    return _FusedConvGRU.apply(x, h, w_zr, b_zr, w_c, b_c, stacks if slot is not None else None, slot)


class _ConvBiasAct(torch.autograd.Function):
    """act(conv(x, w) + b + residual); the backward pass forms the activation derivative and the bias gradient in one kernel.
    With `stacks` / `slot` the weight gradient of the loss window is computed once, by the first pass's backward (WindowStacks)."""

    @staticmethod
    def forward(ctx, x, w, b, residual, act, stride, padding, stacks, slot):
        L = _lib.lib()
        B = x.shape[0]
        if slot is not None:
            xs = conv_input_slot(stacks, slot, x.shape, w, stride, padding, x.device)
            if xs.data_ptr() != x.data_ptr() or xs.stride() != x.stride():
                xs.copy_(x)                       # producers that know their consumer write the slot themselves (decoder_up(out=...))
            x = xs
        else:
            x = _cl(x)
        y = _conv(x, w, stride, padding)
        res = _cl(residual) if residual is not None else None
        M, C = _rows(y), y.shape[1]
        _lib.check(L.tef_bias_act(_lib.ptr(y), _lib.ptr(b), _lib.ptr(res), act, ctypes.c_long(M), C, _lib.stream()), "tef_bias_act")
        ctx.t = (x, w, y)
        ctx.cfg = (act, stride, padding, b is not None, residual is not None, M, C, B, stacks, slot)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        if ctx.t is None:
            raise RuntimeError("a fused network operator can be back-propagated once (its activations are released by the first backward pass)")
        x, w, y = ctx.t
        ctx.t = None
        act, stride, padding, has_b, has_res, M, C, B, stacks, slot = ctx.cfg
        L = _lib.lib()
        gy = _cl(gy)
        if slot is not None:
            gpre = stacks.rows(stacks.g["y"], slot, B)
        else:
            gpre = torch.empty_like(gy) if act else gy
        gb = torch.zeros(C, dtype=torch.float32, device=gy.device) if has_b else None
        if act or has_b:
            _lib.check(L.tef_bias_act_bwd(_lib.ptr(gy), _lib.ptr(y), _lib.ptr(gpre), _lib.ptr(gb), act, ctypes.c_long(M), C, _lib.stream()), "tef_bias_act_bwd")
        elif slot is not None:
            gpre.copy_(gy)
        gx = gw = None
        if slot is None:
            gx, gw = _conv_bwd(gpre, x, w, stride, padding, need_x=ctx.needs_input_grad[0])
        else:
            if ctx.needs_input_grad[0]:
                gx = _conv_bwd_input(gpre, x, w, stride, padding)
            n = stacks.window_rows(slot, B)
            if n is not None:
                gw = _conv_bwd_weight(stacks.g["y"][:n], stacks.x["x"][:n], w, stride, padding)
        # the residual's gradient must not alias a stack slot the next window overwrites while autograd may still hold it
        return gx, gw, gb, ((gpre.clone() if slot is not None else gpre) if has_res else None), None, None, None, None, None


def conv_input_slot(stacks, slot, x_shape, w, stride, padding, device):
    """The [B, Cin, H, W] channels_last view a convolution's input of pass `slot` lives in (allocating the layer's stacks)."""
    B, Cin, H, W = x_shape
    k = w.shape[2]
    Ho, Wo = (H + 2 * padding - k) // stride + 1, (W + 2 * padding - k) // stride + 1
    stacks.ensure(("conv", B, Cin, H, W, w.shape[0], stride, padding), {"x": (B, Cin, H, W)}, {"y": (B, w.shape[0], Ho, Wo)}, device)
    return stacks.rows(stacks.x["x"], slot, B)


def conv_bias_act(x, w, b=None, residual=None, act="relu", stride=1, padding=None, stacks=None, slot=None):
    """act(conv2d(x, w, stride, padding) + b + residual), channels_last fp32 (ConvLayer / ResidualBlock of upstream's submodules).
    `stacks`: the layer's WindowStacks for one weight gradient per loss window; `slot`: a slot already taken from it (by a
    producer that wrote the input in place), else one is taken here."""
    _lib.require_cuda(x, w)
    if padding is None:
        padding = w.shape[2] // 2
    if stacks is not None and slot is None:
        slot = stacks.take_slot()
    return _ConvBiasAct.apply(x, w, b, residual, ACT[act], stride, padding, stacks if slot is not None else None, slot)


class _UpsampleScale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, H, W, scale):
        L = _lib.lib()
        B, C, h, w = pred.shape
        if C != 2 or pred.dtype != torch.float32:
            raise _lib.TefShapeError("upsample_scale expects a [B, 2, h, w] fp32 flow prediction, got %s %s" % (tuple(pred.shape), pred.dtype))
        out = torch.empty((B, 2, H, W), dtype=torch.float32, device=pred.device)
        strides = (ctypes.c_long * 4)(*pred.stride())
        _lib.check(L.tef_upsample_scale(_lib.ptr(pred), strides, h, w, _lib.ptr(out), B, H, W, ctypes.c_float(scale), _lib.stream()), "tef_upsample_scale")
        ctx.geom = (B, h, w, H, W, float(scale), tuple(pred.stride()))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        B, h, w, H, W, scale, strides = ctx.geom
        L = _lib.lib()
        g = g.contiguous()
        gpred = torch.empty_strided((B, 2, h, w), strides, dtype=torch.float32, device=g.device)
        cs = (ctypes.c_long * 4)(*strides)
        _lib.check(L.tef_upsample_scale_bwd(_lib.ptr(g), B, H, W, ctypes.c_float(scale), _lib.ptr(gpred), cs, h, w, _lib.stream()), "tef_upsample_scale_bwd")
        return gpred, None, None, None


def upsample_scale(pred, size, scale):
    """``F.interpolate(pred, size, mode="bilinear", align_corners=False) * scale`` for a 2-channel flow prediction, as one
    kernel each way; the result is the contiguous [B, 2, H, W] map ``update`` packs from (upstream ``models/model.py:65-85``
    and the ``flow_scaling`` of ``train_flow.py:106-108`` folded into `scale`)."""
    _lib.require_cuda(pred)
    if pred.dim() != 4:
        raise _lib.TefShapeError("upsample_scale expects [B, 2, h, w]")
    dense = pred.stride()
    if any(s <= 0 for s in dense):                       # expanded / negative strides: materialise
        pred = pred.contiguous()
    return _UpsampleScale.apply(pred, int(size[0]), int(size[1]), float(scale))


class _DecoderUp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, skip, pred, scale, out):
        L = _lib.lib()
        x = _cl(x)
        skip = _cl(skip) if skip is not None else None
        B, C, h, w = x.shape
        H, W = int(h * scale), int(w * scale)
        Cp = C + (2 if pred is not None else 0)
        out = out[0] if out is not None else None         # handed over in a list: a buffer to fill, not an input of the operator
        if out is None:
            out = torch.empty((B, Cp, H, W), dtype=torch.float32, device=x.device, memory_format=CL)
        elif tuple(out.shape) != (B, Cp, H, W) or not out.is_contiguous(memory_format=CL):
            raise _lib.TefShapeError("decoder_up: out must be a channels_last [%d, %d, %d, %d] tensor" % (B, Cp, H, W))
        ps = (ctypes.c_long * 4)(*pred.stride()) if pred is not None else None
        _lib.check(L.tef_decoder_up(_lib.ptr(x), _lib.ptr(skip), _lib.ptr(pred), ps, _lib.ptr(out), B, h, w, C, H, W, _lib.stream()), "tef_decoder_up")
        ctx.geom = (B, h, w, C, H, W, tuple(pred.stride()) if pred is not None else None, skip is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        B, h, w, C, H, W, pstr, has_skip = ctx.geom
        L = _lib.lib()
        g = _cl(g)
        gx = torch.empty((B, C, h, w), dtype=torch.float32, device=g.device, memory_format=CL)
        gpred = torch.empty_strided((B, 2, h, w), pstr, dtype=torch.float32, device=g.device) if pstr is not None else None
        ps = (ctypes.c_long * 4)(*pstr) if pstr is not None else None
        _lib.check(L.tef_decoder_up_bwd(_lib.ptr(g), _lib.ptr(gx), _lib.ptr(gpred), ps, B, h, w, C, H, W, _lib.stream()), "tef_decoder_up_bwd")
        return gx, (gx if has_skip else None), gpred, None, None


def decoder_up(x, skip=None, pred=None, scale=2, out=None):
    """``F.interpolate(cat([pred, x + skip], 1), scale_factor=scale, mode="bilinear", align_corners=False)`` in one kernel each way
    (the input of a decoder stage of upstream's recurrent U-Net, ``models/arch.py``): channels_last fp32, even channel count.
    `out`: write into this channels_last tensor (the consumer convolution's slot of a WindowStacks) instead of a new one."""
    _lib.require_cuda(x, skip, pred)
    if pred is not None and (pred.dim() != 4 or pred.shape[1] != 2 or any(st <= 0 for st in pred.stride())):
        raise _lib.TefShapeError("decoder_up expects a dense [B, 2, h, w] flow prediction")
    if not (1 <= scale <= 2.5):
        raise _lib.TefError("decoder_up supports scale factors 1 .. 2.5 (the adjoint kernel's window), got %r" % (scale,))
    return _DecoderUp.apply(x, skip, pred, scale, [out] if out is not None else None)
