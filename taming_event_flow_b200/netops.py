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


class _FusedConvGRU(torch.autograd.Function):
    """new_state = ConvGRU(x, h) with the update and reset gates merged into one convolution (w_zr = [w_update; w_reset])."""

    @staticmethod
    def forward(ctx, x, h, w_zr, b_zr, w_c, b_c):
        L = _lib.lib()
        st = _lib.stream()
        Cx, C = x.shape[1], h.shape[1]
        pad = w_zr.shape[2] // 2
        xh = _cl(torch.cat([_cl(x), _cl(h)], 1))
        M = _rows(xh)
        zr = _conv(xh, w_zr, 1, pad)
        xrh = torch.empty_like(xh)
        _lib.check(L.tef_gru_gates(_lib.ptr(zr), _lib.ptr(b_zr), _lib.ptr(xh), _lib.ptr(xrh), ctypes.c_long(M), Cx, C, st), "tef_gru_gates")
        c = _conv(xrh, w_c, 1, pad)
        out = torch.empty_like(c)
        _lib.check(L.tef_gru_output(_lib.ptr(c), _lib.ptr(b_c), _lib.ptr(xh), _lib.ptr(zr), _lib.ptr(out), ctypes.c_long(M), Cx, C, st), "tef_gru_output")
        ctx.save_for_backward(xh, xrh, zr, c, w_zr, w_c)
        ctx.dims = (M, Cx, C, pad, b_zr is not None, b_c is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        xh, xrh, zr, cand, w_zr, w_c = ctx.saved_tensors
        M, Cx, C, pad, has_bzr, has_bc = ctx.dims
        L = _lib.lib()
        st = _lib.stream()
        gout = _cl(gout)
        gc, gzr, gh = torch.empty_like(cand), torch.empty_like(zr), torch.empty_like(cand)
        gb = torch.zeros(3 * C, dtype=torch.float32, device=gout.device)
        gb_zr, gb_c = gb[:2 * C], gb[2 * C:]
        _lib.check(L.tef_gru_output_bwd(_lib.ptr(gout), _lib.ptr(cand), _lib.ptr(xh), _lib.ptr(zr), _lib.ptr(gc), _lib.ptr(gzr), _lib.ptr(gh),
                                        _lib.ptr(gb_c) if has_bc else None, _lib.ptr(gb_zr) if has_bzr else None, ctypes.c_long(M), Cx, C, st),
                   "tef_gru_output_bwd")
        gxrh, gw_c = _conv_bwd(gc, xrh, w_c, 1, pad)
        gxrh = _cl(gxrh)
        _lib.check(L.tef_gru_gates_bwd(_lib.ptr(gxrh), _lib.ptr(xh), _lib.ptr(zr), _lib.ptr(gzr), _lib.ptr(gh), _lib.ptr(gb_zr) if has_bzr else None,
                                       ctypes.c_long(M), Cx, C, st), "tef_gru_gates_bwd")
        gxh, gw_zr = _conv_bwd(gzr, xh, w_zr, 1, pad)
        gxh = _cl(gxh)
        gx = torch.empty((xh.shape[0], Cx, xh.shape[2], xh.shape[3]), dtype=torch.float32, device=gout.device, memory_format=CL)
        _lib.check(L.tef_gru_input_grads(_lib.ptr(gxrh), _lib.ptr(gxh), _lib.ptr(gx), _lib.ptr(gh), ctypes.c_long(M), Cx, C, st), "tef_gru_input_grads")
        return gx, gh, gw_zr, (gb_zr if has_bzr else None), gw_c, (gb_c if has_bc else None)


def conv_gru(x, h, w_zr, b_zr, w_c, b_c):
    """ConvGRU cell (upstream ``models/submodules.py:134-152``): z, r = sigmoid(conv_zr([x, h])); cand = tanh(conv_c([x, h * r]));
    returns h * (1 - z) + cand * z.  `w_zr` stacks the update gate's filters over the reset gate's."""
    _lib.require_cuda(x, h, w_zr, w_c)
    return _FusedConvGRU.apply(x, h, w_zr, b_zr, w_c, b_c)


class _ConvBiasAct(torch.autograd.Function):
    """act(conv(x, w) + b + residual); the backward pass forms the activation derivative and the bias gradient in one kernel."""

    @staticmethod
    def forward(ctx, x, w, b, residual, act, stride, padding):
        L = _lib.lib()
        x = _cl(x)
        y = _conv(x, w, stride, padding)
        res = _cl(residual) if residual is not None else None
        M, C = _rows(y), y.shape[1]
        _lib.check(L.tef_bias_act(_lib.ptr(y), _lib.ptr(b), _lib.ptr(res), act, ctypes.c_long(M), C, _lib.stream()), "tef_bias_act")
        ctx.save_for_backward(x, w, y)
        ctx.cfg = (act, stride, padding, b is not None, residual is not None, M, C)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        act, stride, padding, has_b, has_res, M, C = ctx.cfg
        L = _lib.lib()
        gy = _cl(gy)
        gpre = torch.empty_like(gy) if act else gy
        gb = torch.zeros(C, dtype=torch.float32, device=gy.device) if has_b else None
        if act or has_b:
            _lib.check(L.tef_bias_act_bwd(_lib.ptr(gy), _lib.ptr(y), _lib.ptr(gpre), _lib.ptr(gb), act, ctypes.c_long(M), C, _lib.stream()), "tef_bias_act_bwd")
        gx, gw = _conv_bwd(gpre, x, w, stride, padding, need_x=ctx.needs_input_grad[0])
        return gx, gw, gb, (gpre if has_res else None), None, None, None


def conv_bias_act(x, w, b=None, residual=None, act="relu", stride=1, padding=None):
    """act(conv2d(x, w, stride, padding) + b + residual), channels_last fp32 (ConvLayer / ResidualBlock of upstream's submodules)."""
    _lib.require_cuda(x, w)
    if padding is None:
        padding = w.shape[2] // 2
    return _ConvBiasAct.apply(x, w, b, residual, ACT[act], stride, padding)


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
