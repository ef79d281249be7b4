"""Event-warping primitives, drop-in for the reference's ``utils/iwe.py``.

Same names, arguments, shapes and return values as upstream (``utils/iwe.py:5-257``); every
function is a CUDA kernel launch through the C ABI (``include/tef_b200.h``) and the
differentiable ones are ``torch.autograd.Function`` s whose backward is a kernel as well.
Per-event arithmetic is bit-identical to the reference's eager CPU path; image
accumulation uses fp32 reductions, so only the summation order differs.
"""
import ctypes

import torch

from .._lib import TefShapeError, check, lib, ptr, require_cuda, stream

_f = ctypes.c_float
_l = ctypes.c_long


def _c(t):
    return t.contiguous().float()


# --------------------------------------------------------------------------------------------
class _Propagate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ts, loc, flow, tref):
        ts_c, loc_c, flow_c = _c(ts.expand(loc.shape[:-1] + (1,))), _c(loc), _c(flow.expand_as(loc))
        out = torch.empty_like(loc_c)
        n = loc_c.numel() // 2
        check(lib().tef_event_propagation(ptr(ts_c), ptr(loc_c), ptr(flow_c), _f(tref), ptr(out), _l(n), stream()), "tef_event_propagation")
        ctx.save_for_backward(ts_c, flow_c)
        ctx.tref, ctx.shapes = float(tref), (ts.shape, loc.shape, flow.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        ts_c, flow_c = ctx.saved_tensors
        g = _c(g)
        n = g.numel() // 2
        g_ts = torch.empty_like(ts_c) if ctx.needs_input_grad[0] else None
        g_loc = torch.empty_like(g) if ctx.needs_input_grad[1] else None
        g_flow = torch.empty_like(g) if ctx.needs_input_grad[2] else None
        check(lib().tef_event_propagation_bwd(ptr(g), ptr(ts_c), ptr(flow_c), _f(ctx.tref), ptr(g_ts), ptr(g_loc), ptr(g_flow), _l(n), stream()),
              "tef_event_propagation_bwd")
        s_ts, s_loc, s_flow = ctx.shapes
        if g_ts is not None:
            g_ts = g_ts.sum_to_size(s_ts)
        if g_flow is not None:
            g_flow = g_flow.sum_to_size(s_flow)
        return g_ts, g_loc, g_flow, None


def event_propagation(events_ts, events_idx, flow, tref):
    """Warp events with their flow to ``tref`` (upstream ``utils/iwe.py:5-14``).

    :param events_ts: [batch_size x N x 1] event timestamps
    :param events_idx: [batch_size x N x 2] event locations (y, x)
    :param flow: [batch_size x N x 2] per-event optical flow (y, x)
    :param tref: reference time toward which events are warped
    :return: warped event locations [batch_size x N x 2]
    """
    if not torch.is_tensor(events_ts):      # upstream also passes a plain number (loss/flow_val.py:57)
        events_ts = torch.full(events_idx.shape[:-1] + (1,), float(events_ts), dtype=torch.float32, device=events_idx.device)
    require_cuda(events_ts, events_idx, flow)
    return _Propagate.apply(events_ts, events_idx, flow, float(tref))


# --------------------------------------------------------------------------------------------
class _EventFlow(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mapx, mapy, loc):
        mx, my, lc = _c(mapx), _c(mapy), _c(loc)
        B, H, W = mx.shape
        N = lc.shape[1]
        out = torch.empty((B, N, 2), dtype=torch.float32, device=lc.device)
        check(lib().tef_get_event_flow(ptr(mx), ptr(my), ptr(lc), ptr(out), B, N, H, W, stream()), "tef_get_event_flow")
        ctx.save_for_backward(mx, my, lc)
        return out

    @staticmethod
    def backward(ctx, g):
        mx, my, lc = ctx.saved_tensors
        B, H, W = mx.shape
        N = lc.shape[1]
        g = _c(g)
        gmx = torch.zeros_like(mx) if ctx.needs_input_grad[0] else None
        gmy = torch.zeros_like(my) if ctx.needs_input_grad[1] else None
        gl = torch.empty_like(lc) if ctx.needs_input_grad[2] else None
        check(lib().tef_get_event_flow_bwd(ptr(g), ptr(mx), ptr(my), ptr(lc), ptr(gmx), ptr(gmy), ptr(gl), B, N, H, W, stream()),
              "tef_get_event_flow_bwd")
        return gmx, gmy, gl


def get_event_flow(flow_map_x, flow_map_y, event_loc):
    """Sample the flow maps at event locations, bilinear (upstream ``utils/iwe.py:17-40``).

    :param flow_map_x: [batch_size x H x W] horizontal flow map
    :param flow_map_y: [batch_size x H x W] vertical flow map
    :param event_loc: [batch_size x N x 2] event locations (y, x)
    :return: [batch_size x N x 2] per-event flow (y, x)
    """
    require_cuda(flow_map_x, flow_map_y, event_loc)
    return _EventFlow.apply(flow_map_x, flow_map_y, event_loc)


# --------------------------------------------------------------------------------------------
class _Purge(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loc, mask, H, W):
        lc, mk = _c(loc), _c(mask)
        ol, om = torch.empty_like(lc), torch.empty_like(mk)
        n = lc.numel() // 2
        check(lib().tef_purge_unfeasible(ptr(lc), ptr(mk), ptr(ol), ptr(om), _l(n), H, W, stream()), "tef_purge_unfeasible")
        ctx.save_for_backward(lc)
        ctx.res = (H, W)
        return ol, om

    @staticmethod
    def backward(ctx, g_loc, g_mask):
        (lc,) = ctx.saved_tensors
        n = lc.numel() // 2
        gl_in = _c(g_loc) if g_loc is not None else None
        gm_in = _c(g_mask) if g_mask is not None else None
        gl = torch.empty_like(lc) if (ctx.needs_input_grad[0] and gl_in is not None) else None
        gm = torch.empty_like(lc) if (ctx.needs_input_grad[1] and gm_in is not None) else None
        check(lib().tef_purge_unfeasible_bwd(ptr(lc), ptr(gl_in), ptr(gm_in), ptr(gl), ptr(gm), _l(n), ctx.res[0], ctx.res[1], stream()),
              "tef_purge_unfeasible_bwd")
        return gl, gm, None, None


def purge_unfeasible(event_loc, event_pol_mask, res):
    """Zero the location and polarity mask of events warped outside the image (upstream ``utils/iwe.py:43-60``)."""
    require_cuda(event_loc, event_pol_mask)
    if event_loc.dim() < 1 or event_loc.shape[-1] != 2:
        raise TefShapeError("event_loc must be [..., 2] (y, x), got %s" % (tuple(event_loc.shape),))
    cols = event_pol_mask.shape[-1] if event_pol_mask.dim() else 0
    if event_pol_mask.shape != event_loc.shape:       # e.g. a [B,N,1] mask (loss/flow_val.py:58): broadcast like upstream
        if event_pol_mask.dim() != event_loc.dim() or event_pol_mask.shape[:-1] != event_loc.shape[:-1] or cols not in (1, 2):
            raise TefShapeError("polarity mask %s does not match event locations %s ([..., 1] or [..., 2] expected)"
                                % (tuple(event_pol_mask.shape), tuple(event_loc.shape)))
        event_pol_mask = event_pol_mask.expand_as(event_loc)
    loc, mask = _Purge.apply(event_loc, event_pol_mask, int(res[0]), int(res[1]))
    return loc, (mask if cols == mask.shape[-1] else mask[..., 0:cols])


# --------------------------------------------------------------------------------------------
class _Interpolation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, warped, H, W, round_idx):
        wp = _c(warped)
        B, N = wp.shape[0], wp.shape[1]
        M = N if round_idx else 4 * N
        idx = torch.empty((B, M, 1), dtype=torch.float32, device=wp.device)
        wts = torch.empty((B, M, 1), dtype=torch.float32, device=wp.device)
        check(lib().tef_get_interpolation(ptr(wp), ptr(idx), ptr(wts), B, N, H, W, int(round_idx), stream()), "tef_get_interpolation")
        ctx.save_for_backward(wp)
        ctx.res, ctx.round_idx = (H, W), round_idx
        ctx.mark_non_differentiable(idx)
        return idx, wts

    @staticmethod
    def backward(ctx, g_idx, g_w):
        (wp,) = ctx.saved_tensors
        if ctx.round_idx or g_w is None:
            return torch.zeros_like(wp), None, None, None
        B, N = wp.shape[0], wp.shape[1]
        g = torch.empty_like(wp)
        check(lib().tef_get_interpolation_bwd(ptr(wp), ptr(_c(g_w)), ptr(g), B, N, ctx.res[0], ctx.res[1], stream()), "tef_get_interpolation_bwd")
        return g, None, None, None


def get_interpolation(warped_events, res, round_idx=False, zeros=None):
    """Scatter indices and bilinear (or rounding) weights of warped events (upstream ``utils/iwe.py:63-113``).

    Returns ``idx`` and ``weights`` of shape [batch_size x 4N x 1], corner-major (top-left, top-right,
    bottom-left, bottom-right), or [batch_size x N x 1] when ``round_idx``.  ``zeros`` is accepted for
    signature compatibility (upstream uses it as a scratch tensor) and ignored.
    """
    require_cuda(warped_events)
    return _Interpolation.apply(warped_events, int(res[0]), int(res[1]), bool(round_idx))


# --------------------------------------------------------------------------------------------
class _Interpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, weights, pol, zeros, H, W):
        ix, wt = _c(idx), _c(weights)
        B, M = ix.shape[0], ix.shape[1]
        pl = _c(pol.expand_as(wt)) if pol is not None else None
        if zeros is None:
            iwe = torch.zeros((B, H * W, 1), dtype=torch.float32, device=ix.device)
        else:
            iwe = zeros.detach().clone().float().contiguous().view(B, H * W, 1)
        check(lib().tef_interpolate(ptr(ix), ptr(wt), ptr(pl), ptr(iwe), B, _l(M), H, W, stream()), "tef_interpolate")
        ctx.save_for_backward(ix, wt, pl)
        ctx.res = (H, W)
        ctx.zshape = None if zeros is None else zeros.shape
        return iwe.view(B, 1, H, W)

    @staticmethod
    def backward(ctx, g):
        ix, wt, pl = ctx.saved_tensors
        B, M = ix.shape[0], ix.shape[1]
        H, W = ctx.res
        g = _c(g)
        g_w = torch.empty_like(wt) if ctx.needs_input_grad[1] else None
        g_p = torch.empty_like(wt) if (ctx.needs_input_grad[2] and pl is not None) else None
        if g_w is not None or g_p is not None:
            check(lib().tef_interpolate_bwd(ptr(ix), ptr(pl), ptr(wt), ptr(g), ptr(g_w), ptr(g_p), B, _l(M), H, W, stream()), "tef_interpolate_bwd")
        g_z = g.reshape(ctx.zshape) if (ctx.needs_input_grad[3] and ctx.zshape is not None) else None
        return None, g_w, g_p, g_z, None, None


def interpolate(idx, weights, res, polarity_mask=None, zeros=None):
    """Accumulate weighted events into an image of warped events (upstream ``utils/iwe.py:116-136``).

    :return: [batch_size x 1 x H x W] image; ``zeros`` (if given) is the start image and is not modified

    Indices outside ``[0, H*W)`` (which `get_interpolation` never produces; upstream's ``scatter_add_`` raises on them)
    are skipped.
    """
    require_cuda(idx, weights, polarity_mask, zeros)
    return _Interpolate.apply(idx, weights, polarity_mask, zeros, int(res[0]), int(res[1]))


# --------------------------------------------------------------------------------------------
def _check_deblur_shapes(fl, ev, H, W):
    if ev.dim() != 3 or ev.shape[2] != 4:
        raise TefShapeError("event_list must be [B,N,4], got %s" % (tuple(ev.shape),))
    if tuple(fl.shape) != (ev.shape[0], 2, H, W):
        raise TefShapeError("flow must be [%d,2,%d,%d], got %s" % (ev.shape[0], H, W, tuple(fl.shape)))


def _deblur_into(out_view, batch_stride, flow, event_list, res, round_idx, pol, pol_stride, round_flow):
    B, N = event_list.shape[0], event_list.shape[1]
    check(lib().tef_deblur_events(ptr(flow), ptr(event_list), ptr(pol) if pol is not None else None, _l(pol_stride), ptr(out_view), _l(batch_stride),
                                  B, N, int(res[0]), int(res[1]), int(bool(round_idx)), int(bool(round_flow)), stream()), "tef_deblur_events")


def deblur_events(flow, event_list, res, round_idx=True, polarity_mask=None, round_flow=True):
    """Image of events warped to t=1 with one flow map (upstream ``utils/iwe.py:139-224``), one fused kernel.

    Evaluation / visualisation path: runs without autograd, like its upstream call sites (``eval_flow.py:70,104``).
    """
    require_cuda(flow, event_list, polarity_mask)
    H, W = int(res[0]), int(res[1])
    fl, ev = _c(flow.detach()), _c(event_list.detach())
    _check_deblur_shapes(fl, ev, H, W)
    B = ev.shape[0]
    # upstream multiplies the weights by the polarity mask in both branches (:216-222)
    pol = _c(polarity_mask.detach()) if polarity_mask is not None else None
    if pol is not None and tuple(pol.shape) != (B, ev.shape[1], 1):
        # the kernel reads one mask value per event (stride 1): a [B,N,2] mask would be read with the wrong stride
        raise TefShapeError("polarity_mask must be [%d,%d,1] (one column of the loader's mask), got %s" % (B, ev.shape[1], tuple(pol.shape)))
    iwe = torch.empty((B, 1, H, W), dtype=torch.float32, device=ev.device)
    _deblur_into(iwe, H * W, fl, ev, res, round_idx, pol, 1, round_flow)
    return iwe


def compute_pol_iwe(flow, event_list, res, pol_mask, round_idx=True, round_flow=True):
    """Per-polarity image of warped events (upstream ``utils/iwe.py:227-257``): [batch_size x 2 x H x W]."""
    require_cuda(flow, event_list, pol_mask)
    H, W = int(res[0]), int(res[1])
    fl, ev, pm = _c(flow.detach()), _c(event_list.detach()), _c(pol_mask.detach())
    _check_deblur_shapes(fl, ev, H, W)
    B = ev.shape[0]
    if tuple(pm.shape) != (B, ev.shape[1], 2):
        raise TefShapeError("pol_mask must be [%d,%d,2], got %s" % (B, ev.shape[1], tuple(pm.shape)))
    iwe = torch.empty((B, 2, H, W), dtype=torch.float32, device=ev.device)
    for ch in range(2):
        _deblur_into(iwe[:, ch], 2 * H * W, fl, ev, res, round_idx, pm.view(-1)[ch:], 2, round_flow)
    return iwe
