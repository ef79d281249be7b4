"""Contrast-maximization losses, drop-in for the reference's ``loss/flow.py``.

Same classes, constructor arguments, ``update`` / ``reset`` / ``num_passes`` /
``__call__`` protocol and config keys as upstream (``loss/flow.py:14-746``), but the
whole loss -- event warping, IWE splatting, focus loss and the backward pass -- runs
in the fused CUDA kernels of ``csrc/tef_cm_*.cu`` instead of ~16 k eager tensor ops.

What `update` keeps per pass is the packed flow map (float2 interleaved) and a staged
copy of the event rows; nothing of size O(passes^2 * events) is ever materialised.
The returned loss is a 0-dim tensor wired into autograd: ``loss.backward()`` delivers
gradients to every tensor of every ``flow_list`` handed to `update`.
"""
import ctypes
import os

import torch

from .. import _lib
from .._lib import CmDesc, TefShapeError, check, lib, ptr, require_cuda, stream

_MODES = {"one": 1, "two": 2, "four": 4}


# TEF_FUSED_HIST=0: build the tile-sort histogram inside the forward call instead of inside update() (A/B switch)
_FUSED_HIST = os.environ.get("TEF_FUSED_HIST", "1") != "0"
# TEF_QUAD=1: quad-cell copies of the flow maps and gradient images (one 256-bit gather per 2x2 fetch).  Measured on B200
# (DESIGN.md decision 15): a third fewer load sectors, same kernel times, +0.13 ms for writing the copies -- off by default.
_QUAD = os.environ.get("TEF_QUAD", "0") == "1"


class _Workspace:
    """Grow-only device buffers reused from one loss window to the next, so that a training loop allocates
    nothing in steady state (the buffers are several hundred MB at DSEC resolution; going through the caching
    allocator every step fragments it and stalls on cudaMalloc/cudaFree)."""

    def __init__(self):
        self.bufs = {}
        self.views = {}        # (key, shape) -> view of the current buffer: steady-state windows ask for the same shapes

    def get(self, key, shape, dtype, device):
        hit = self.views.get((key, shape))
        if hit is not None and hit.dtype == dtype and hit.device == device:
            return hit
        n = 1
        for d in shape:
            n *= int(d)
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < n or buf.dtype != dtype or buf.device != device:
            buf = torch.empty((max(n, 1),), dtype=dtype, device=device)
            self.bufs[key] = buf
            self.views = {k: v for k, v in self.views.items() if k[0] != key}     # views of the old buffer are stale
        view = buf[:n].view(shape)
        if len(self.views) > 512:                   # ragged windows ask for ever-changing shapes: keep the cache bounded
            self.views.clear()
        self.views[(key, shape)] = view
        return view


class _Window:
    """Device state of one loss window (between two `reset` calls)."""

    def __init__(self, pool):
        self.pool = pool       # the module's list of idle workspaces
        self.ws = pool.pop() if pool else _Workspace()
        self.flows = []        # flows[t][f]: the caller's tensors (autograd leaves of the loss)
        self.packed = None     # [F,P,B,2,H+1,Wp,2] dual-phase rows
        self.packedq = None    # [F,P,B,4,H/2+1,W/2+1,8] quad cells (Iterative, non-deterministic) or None
        self.ev = ([], [])     # device addresses of the staged event rows per pass, (grad set, detached set)
        self.mk = ([], [])
        self.n = ([], [])
        self.sort = None       # (bins, sums, sorted_ev, sorted_mk), written by the forward kernels
        self.shape = None      # (F, B, H, W)
        self.img = None
        self.den = None
        self.consumed = False
        self.hist_valid = False    # sort_bins holds the fused histogram of every pass given to update() so far
        self.desc = None           # descriptor of the last forward call, reused by backward

    def release(self):
        """Hand the workspace back (after backward, or when the window dies un-differentiated)."""
        if self.ws is not None:
            self.pool.append(self.ws)
            self.ws = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class _CMLoss(torch.autograd.Function):
    """loss = CM(flow maps); backward = analytic gradient w.r.t. every flow map."""

    @staticmethod
    def forward(ctx, module, window, *flat_flows):
        loss = module._forward_kernels(window)
        ctx.module, ctx.window = module, window
        ctx.nflows = len(flat_flows)
        return loss

    @staticmethod
    def backward(ctx, gout):
        grads = ctx.module._backward_kernels(ctx.window, gout)     # [P,F,B,2,H,W]
        P, F = grads.shape[0], grads.shape[1]
        out = grads.view((P * F,) + grads.shape[2:]).unbind(0)     # one call instead of P*F Python-level slices
        extra = (len(ctx.window.flows) - P) * F                    # passes beyond the loss window get no gradient
        return (None, None) + out + (None,) * extra


class _Smoothness(torch.autograd.Function):
    """Per-sample value [B] of one smoothness prior over the first `n` passes of the window; backward = its gradient
    w.r.t. every flow map of those passes."""

    @staticmethod
    def forward(ctx, module, w, temporal, n, *flat_flows):
        F, B, H, W = w.shape
        P = module._max_passes()
        dev = w.packed.device
        L = lib()
        L.tef_flow_smoothing_scratch.restype = ctypes.c_long
        scratch = torch.empty((max(L.tef_flow_smoothing_scratch(B, H, W, n, F), 1),), dtype=torch.float32, device=dev)
        out = torch.empty((B,), dtype=torch.float32, device=dev)
        sums = None
        if temporal:
            sums = torch.empty((F, n - 1, B, 2), dtype=torch.float32, device=dev)
            check(L.tef_flow_temporal_smoothing(ptr(w.packed), B, H, W, P, F, n, ptr(scratch), ptr(sums), ptr(out), stream()), "tef_flow_temporal_smoothing")
        else:
            check(L.tef_flow_spatial_smoothing(ptr(w.packed), B, H, W, P, F, n, ptr(scratch), ptr(out), stream()), "tef_flow_spatial_smoothing")
        ctx.module, ctx.w, ctx.temporal, ctx.n, ctx.sums = module, w, temporal, n, sums
        return out

    @staticmethod
    def backward(ctx, gout):
        w, n = ctx.w, ctx.n
        F, B, H, W = w.shape
        P = ctx.module._max_passes()
        dev = w.packed.device
        Wp = (W + 3) & ~1
        gpacked = torch.empty((F * P * B * 2 * H * Wp * 2,), dtype=torch.float32, device=dev)
        grads = torch.empty((P, F, B, 2, H, W), dtype=torch.float32, device=dev)
        g = gout.detach().float().contiguous()
        L = lib()
        if ctx.temporal:
            check(L.tef_flow_temporal_smoothing_bwd(ptr(w.packed), ptr(ctx.sums), ptr(g), ptr(gpacked), B, H, W, P, F, n, stream()),
                  "tef_flow_temporal_smoothing_bwd")
        else:
            check(L.tef_flow_spatial_smoothing_bwd(ptr(w.packed), ptr(g), ptr(gpacked), B, H, W, P, F, n, stream()), "tef_flow_spatial_smoothing_bwd")
        check(L.tef_unpack_flow_grad(ptr(gpacked), ptr(grads), F, P, B, H, W, 0, None, stream()), "tef_unpack_flow_grad")
        out = [grads[t, f] if t < n else None for t in range(len(w.flows)) for f in range(F)]
        return (None, None, None, None) + tuple(out)


class BaseEventWarping(torch.nn.Module):
    """Base class of the CM losses (upstream ``loss/flow.py:14-213``)."""

    _linear = False

    def __init__(self, config, device, loss_scaling=True, border_compensation=True):
        super().__init__()
        self.device = device
        self.config = config
        self.loss_scaling = loss_scaling
        self.border_compensation = border_compensation
        self.res = config["loader"]["resolution"]
        self.batch_size = config["loader"]["batch_size"]
        self.flow_spat_smooth_weight = config["loss"]["flow_spat_smooth_weight"]
        self.flow_temp_smooth_weight = config["loss"]["flow_temp_smooth_weight"]
        self.deterministic = bool(config["loss"].get("deterministic", False))

        self._passes = 0
        self._num_flows = None
        self._pool = []
        self._win = _Window(self._pool)
        self._fn_update = lib().tef_update_pass

        # timescales for loss computation (loss/flow.py:42-44)
        self.passes_loss = [config["data"]["passes_loss"] // (2 ** s) for s in range(config["data"]["scales_loss"])]

    # ------------------------------------------------------------------ state
    @property
    def num_passes(self):
        return self._passes

    def reset_base(self):
        self._passes = 0
        self._win = None           # drop first: an un-differentiated window returns its workspace to the pool
        self._win = _Window(self._pool)   # a pending autograd graph keeps its own reference to the old window

    def reset(self):
        self.reset_base()

    def _max_passes(self):
        return max(self.passes_loss)

    # ----------------------------------------------------------------- update
    def update_base(self, flow_list):
        """Validate this pass' flow maps and make sure the window has its packed buffer (upstream ``update_base``,
        loss/flow.py:46-66).  The packing itself happens in `_update_pass`, fused with the event staging."""
        w = self._win
        if w.ws is None:
            raise RuntimeError("update() after backward(): call reset() first (upstream resets after every loss, train_flow.py:136-137)")
        if self._num_flows is None:
            self._num_flows = len(flow_list)
        F = self._num_flows
        if len(flow_list) != F:
            raise ValueError("flow_list has %d maps, expected %d" % (len(flow_list), F))
        if F > _lib.MAX_FLOWS:
            raise _lib.TefError("at most %d flow maps per pass" % _lib.MAX_FLOWS)
        f0 = flow_list[0]
        require_cuda(*flow_list)
        B, C, H, W = f0.shape if f0.dim() == 4 else (0, 0, 0, 0)
        if C != 2 or H != self.res[0] or W != self.res[1]:
            raise TefShapeError("flow maps must be [B,2,%d,%d], got %s" % (self.res[0], self.res[1], tuple(f0.shape)))
        for fl in flow_list[1:]:
            if fl.shape != f0.shape:       # every map is read as [B,2,H,W] with the batch size of the first one
                raise TefShapeError("every flow map of a pass must be %s, got %s" % (tuple(f0.shape), tuple(fl.shape)))
        if w.shape is not None and w.shape != (F, B, H, W):
            raise TefShapeError("flow maps changed shape inside a loss window: %s, then %s" % (w.shape, (F, B, H, W)))
        if w.packed is None:
            w.shape = (F, B, H, W)
            Wp = (W + 3) & ~1                      # dual-phase, zero-padded maps (csrc/tef_device.cuh)
            w.packed = w.ws.get("packed", (F, self._max_passes(), B, 2, H + 1, Wp, 2), torch.float32, f0.device)
            if _QUAD and not self._linear and not self.deterministic:
                w.packedq = w.ws.get("packedq", (F, self._max_passes(), B, 4, H // 2 + 1, W // 2 + 1, 8), torch.float32, f0.device)
        w.flows.append(list(flow_list))

    def _update_pass(self, flow_list, event_list, pol_mask, d_event_list, d_pol_mask):
        """One `tef_update_pass` call: pack the F flow maps of this pass, add the pass index to the caller's timestamps in
        place (loss/flow.py:457-458) and keep a staged copy of the event rows and masks of both sets (:459-473)."""
        w = self._win
        t = self._passes
        F, B, H, W = w.shape
        P = self._max_passes()
        u = _lib.UpdateDesc()
        u.F, u.t, u.P, u.B, u.H, u.W = F, min(t, P - 1), P, B, H, W
        keep = []                      # temporaries must outlive the (asynchronous) launches only in stream order
        for f, fl in enumerate(flow_list):
            if not (fl.is_cuda and fl.is_contiguous() and fl.dtype == torch.float32):
                require_cuda(fl)
                fl = fl.detach().contiguous().float()
                keep.append(fl)
            u.flow_maps[f] = fl.data_ptr()
        u.packed = w.packed.data_ptr()
        if w.packedq is not None:
            u.packedq = w.packedq.data_ptr()
        round_ts = self.config["loss"]["round_ts"]
        dev = w.packed.device
        for k, (ev, mk) in enumerate(((event_list, pol_mask), (d_event_list, d_pol_mask))):
            require_cuda(ev, mk)
            # the kernels read B * N rows of 4 (events) and 2 (mask) floats, with B taken from the flow maps: anything else
            # would be an out-of-bounds device read, where upstream fails with a torch shape error
            if ev.dim() != 3 or ev.shape[0] != B or ev.shape[2] != 4:
                raise TefShapeError("event list must be [%d,N,4] (batch size of the flow maps), got %s" % (B, tuple(ev.shape)))
            if mk.dim() != 3 or mk.shape[0] != B or mk.shape[1] != ev.shape[1] or mk.shape[2] != 2:
                raise TefShapeError("polarity mask must be [%d,%d,2] like its event list, got %s" % (B, ev.shape[1], tuple(mk.shape)))
            N = ev.shape[1]
            rows = B * N
            buf = w.ws.get(("stage", k, t), (rows, 6), torch.float32, dev)
            base = buf.data_ptr()
            w.ev[k].append(base)
            w.mk[k].append(base + rows * 16)
            w.n[k].append(N)
            u.rows[k] = rows
            if round_ts:
                # event_ts[...] = event_ts.min() + 0.5 (:461-463); min() of an empty tensor raises, like upstream
                ov = (ev[:, :, 0].min() + float(t) + 0.5).float().reshape(1)
                keep.append(ov)
                u.ts_override[k] = ov.data_ptr()
            if rows == 0:
                continue
            if ev.dtype != torch.float32:
                ev[:, :, 0:1] += t             # the in-place update of the caller's tensor, then a float copy for the kernels
                ev = ev.float()
                keep.append(ev)
                u.pass_index[k] = 0.0
            else:
                u.pass_index[k] = float(t)
            if mk.dtype != torch.float32:
                mk = mk.float()
                keep.append(mk)
            u.events[k], u.masks[k] = ev.data_ptr(), mk.data_ptr()
            if not (ev.is_contiguous() and mk.is_contiguous()):
                # e.g. the transposed views upstream's custom_collate returns: read through their strides (same launch)
                u.strided[k] = 1
                for j in range(3):
                    u.ev_strides[k][j], u.mk_strides[k][j] = ev.stride(j), mk.stride(j)
            u.ev_out[k], u.mk_out[k] = base, base + rows * 16
        if t >= P:
            # passes beyond the loss window are not part of the loss (upstream never reads them); only the in-place
            # timestamp update is observable, so stage into the scratch rows and skip the packing slot
            u.F = 0
        elif _FUSED_HIST and (t == 0 or w.hist_valid):
            # the staging kernel also counts the events into the tile-sort histogram (saves the forward a pass over them)
            tiles = ((W + 15) // 16) * ((H + 7) // 8)
            bins = w.ws.get("bins", (2 * P * B * tiles * 256 + 1,), torch.int32, dev)      # 16 x 8 pixels x 2 polarities per tile
            u.sort_bins, u.hist, u.zero_bins = bins.data_ptr(), 1, int(t == 0)
            w.hist_valid = True
        check(self._fn_update(ctypes.byref(u), stream()), "tef_update_pass")

    # ---------------------------------------------------------------- kernels
    def _desc(self, w):
        F, B, H, W = w.shape
        P = self._max_passes()
        d = CmDesc()
        d.B, d.H, d.W, d.P, d.F = B, H, W, P, F
        d.S = self.config["data"]["scales_loss"]
        d.mode = 0 if self._linear else _MODES[self.config["loss"]["iterative_mode"]]
        d.border_comp = int(bool(self.border_compensation))
        d.loss_scaling = int(bool(self.loss_scaling))
        d.deterministic = int(self.deterministic)
        for k in range(2):
            for t in range(P):
                d.ev[k][t] = w.ev[k][t]
                d.mk[k][t] = w.mk[k][t]
                d.n[k][t] = w.n[k][t]
        d.flow = w.packed.data_ptr()
        if w.packedq is not None:
            d.flowq = w.packedq.data_ptr()
        d.hist_done = int(w.hist_valid)
        return d

    def _forward_kernels(self, w):
        F, B, H, W = w.shape
        d = self._desc(w)
        sz = (ctypes.c_long * 13)()
        check(lib().tef_cm_sizes(ctypes.byref(d), int(self._linear), sz), "tef_cm_sizes")
        nslots, n_img, w.n_gflow, n_bins, n_sums, rows, rows_grad, n_pos, w.Wp, nchunks, n_gimg, w.n_gimgq, _ = (int(v) for v in sz)
        dev = w.packed.device
        i32, f32 = torch.int32, torch.float32
        if w.ws is None:
            raise RuntimeError("this loss window has already been back-propagated; call reset() and update() again")
        g = w.ws.get
        w.sort = (g("bins", (n_bins,), i32, dev), g("sums", (n_sums,), i32, dev), g("sorted_ev", (max(rows, 1), 8), f32, dev),
                  g("posbuf", (max(n_pos, 1),), f32, dev), g("alive", (max(F * rows_grad, 1),), i32, dev))
        w.img = g("img", (n_img,), f32, dev)
        w.gimg = g("gimg", (max(n_gimg, 1),), f32, dev)
        w.acc_sum = g("acc_sum", (F, B, nslots, nchunks), torch.float64, dev)
        w.acc_nnz = g("acc_nnz", (F, B, nslots, nchunks), i32, dev)
        w.den = g("den", (F * B * nslots + 2,), f32, dev)            # + scale, 1/scale of the deterministic gradient words
        w.nslots = nslots
        loss = torch.empty((1,), dtype=f32, device=dev)
        self._fill_workspace(d, w)
        d.acc_sum, d.acc_nnz, d.loss = w.acc_sum.data_ptr(), w.acc_nnz.data_ptr(), loss.data_ptr()
        fn = lib().tef_linear_forward if self._linear else lib().tef_iterative_forward
        check(fn(ctypes.byref(d), stream()), "tef_linear_forward" if self._linear else "tef_iterative_forward")
        w.hist_valid = False       # the scan turned the counts into offsets: a second forward() counts again by itself
        d.hist_done = 0
        w.desc = d
        w.consumed = False
        return loss.view(())

    @staticmethod
    def _fill_workspace(d, w):
        d.sort_bins, d.sort_sums, d.sorted_ev, d.posbuf, d.alivebuf = (x.data_ptr() for x in w.sort)
        d.img, d.den, d.gimg = w.img.data_ptr(), w.den.data_ptr(), w.gimg.data_ptr()

    def _backward_kernels(self, w, gout):
        if w.consumed:
            raise RuntimeError("the CM loss graph has already been back-propagated (the image buffers are reused in place)")
        F, B, H, W = w.shape
        P = self._max_passes()
        d = w.desc if w.desc is not None else self._desc(w)          # the descriptor of the forward call (same workspace)
        dev = w.packed.device
        gpacked = w.ws.get("gpacked", (w.n_gflow,), torch.float32, dev)
        if w.packedq is not None and w.n_gimgq > 0:
            d.gimgq = w.ws.get("gimgq", (w.n_gimgq,), torch.float32, dev).data_ptr()
        grads = torch.empty((P, F, B, 2, H, W), dtype=torch.float32, device=dev)     # handed to autograd: not pooled
        g = gout.detach().float().contiguous().reshape(1)
        self._fill_workspace(d, w)
        d.gflow, d.grad_out = gpacked.data_ptr(), g.data_ptr()
        fn = lib().tef_linear_backward if self._linear else lib().tef_iterative_backward
        check(fn(ctypes.byref(d), stream()), "tef_linear_backward" if self._linear else "tef_iterative_backward")
        det_scale = ctypes.c_void_p(w.den.data_ptr() + 4 * F * B * w.nslots) if self.deterministic else None
        check(lib().tef_unpack_flow_grad(ptr(gpacked), ptr(grads), F, P, B, H, W, int(self.deterministic), det_scale, stream()), "tef_unpack_flow_grad")
        w.consumed = True
        w.release()                # stream order makes reuse by the next window safe
        return grads

    def images(self):
        """Diagnostic view of the accumulated images of the last forward call (before backward):
        [F, B, slots, 4, H, W] with channels (count+, count-, time-weighted+, time-weighted-)."""
        w = self._win
        if w.img is None or w.consumed:
            raise RuntimeError("images() is only available between forward() and backward()")
        F, B, H, W = w.shape
        if self.deterministic:                                        # int64 fixed point: high words (2^-40), then low words (2^-88)
            v = w.img.view(torch.int64).view(2, F, B, w.nslots, 2, 2, H, w.Wp, 2)
            v = v[:, :, :, :, 0, :, :, 0:W, :] + v[:, :, :, :, 1, :, :, 1:W + 1, :]
            s = (v[0].double() * 2.0 ** -40 + v[1].double() * 2.0 ** -88).float()
        else:
            v = w.img.view(F, B, w.nslots, 2, 2, H, w.Wp, 2)          # [.., phase, pol, H, Wp, (count, tw)]
            s = v[:, :, :, 0, :, :, 0:W, :] + v[:, :, :, 1, :, :, 1:W + 1, :]
        return torch.stack([s[:, :, :, 0, :, :, 0], s[:, :, :, 1, :, :, 0], s[:, :, :, 0, :, :, 1], s[:, :, :, 1, :, :, 1]], dim=3)

    # ---------------------------------------------------------------- forward
    def _cm_loss(self):
        w = self._win
        P = self._max_passes()
        if self._passes < P or w.packed is None:
            # upstream indexes lists of length num_passes with range(max_passes)
            raise IndexError("the loss window needs %d passes, only %d were given to update()" % (P, self._passes))
        flat = [fl for per_pass in w.flows for fl in per_pass]
        if torch.is_grad_enabled() and any(fl.requires_grad for fl in flat):
            return _CMLoss.apply(self, w, *flat)
        return self._forward_kernels(w)

    def _smoothing_terms(self, loss):
        if self.flow_spat_smooth_weight is not None:
            loss = loss + self.flow_spatial_smoothing()
        if self.flow_temp_smooth_weight is not None and self._passes > 1:
            loss = loss + self.flow_temporal_smoothing()
        return loss

    # -------------------------------------------------- stand-alone pieces of the upstream API
    def iwe_formatting(self, warped_events, pol_mask, ts_list, tref, ts_scaling, interp_zeros=None, iwe_zeros=None):
        """Count and time-weighted images of warped events (upstream ``iwe_formatting``, loss/flow.py:81-110).

        Kept for callers of the upstream method; `forward` does not go through it (the fused kernels
        never materialise indices or weights)."""
        from ..utils.iwe import get_interpolation, interpolate

        norm_ts = 1 - torch.abs(tref - ts_list) / ts_scaling
        idx, weights = get_interpolation(warped_events, self.res, zeros=interp_zeros)
        iwe = torch.cat([interpolate(idx, weights, self.res, polarity_mask=pol_mask[:, :, c:c + 1], zeros=iwe_zeros) for c in range(2)], dim=1)
        wts = weights * norm_ts
        iwe_ts = torch.cat([interpolate(idx, wts, self.res, polarity_mask=pol_mask[:, :, c:c + 1], zeros=iwe_zeros) for c in range(2)], dim=1)
        return iwe, iwe_ts

    def focus_loss(self, iwe, iwe_ts):
        """Sum of squared per-pixel average timestamps, scaled by the number of pixels with events
        (upstream ``focus_loss``, loss/flow.py:112-129); summed over the batch."""
        sq = iwe_ts.reshape(iwe_ts.shape[0], 2, -1) ** 2
        loss = sq[:, 0].sum(1) + sq[:, 1].sum(1)
        if self.loss_scaling:
            nonzero = iwe.sum(1, keepdim=True).bool().reshape(iwe.shape[0], -1)
            loss = loss / (nonzero.sum(1) + 1e-9)
        return loss.sum()

    # Smoothness priors (upstream loss/flow.py:131-209; `Null` in every shipped config): fused kernels on the packed flow
    # maps `update` built (csrc/tef_cm_smooth.cu), gradients to the same flow tensors the CM loss differentiates.
    def _smoothness(self, temporal):
        w = self._win
        if w.packed is None:
            raise IndexError("no flow maps: call update() first")
        if self._passes > self._max_passes():
            # upstream's priors cover every pass given to update(); here only the passes of the loss window are packed
            raise NotImplementedError("the smoothness priors cover at most passes_loss = %d passes per window, update() was called %d times; "
                                      "call reset() after every loss like upstream's train_flow.py:136-137" % (self._max_passes(), self._passes))
        flat = [fl for per_pass in w.flows for fl in per_pass]
        return _Smoothness.apply(self, w, temporal, self._passes, *flat).sum()

    def flow_spatial_smoothing(self):
        """Charbonnier penalty on horizontal, vertical and both diagonal flow differences (upstream :170-209)."""
        return self.flow_spat_smooth_weight * self._smoothness(False)

    def flow_temporal_smoothing(self):
        """Charbonnier penalty between each flow map and the next one sampled where the flow points (upstream :131-168)."""
        return self.flow_temp_smooth_weight * self._smoothness(True)

    def forward(self):
        raise NotImplementedError


class Iterative(BaseEventWarping):
    """CM loss with iterative warping, all intermediate reference times and several temporal
    scales (upstream ``loss/flow.py:415-746``)."""

    def __init__(self, config, device, loss_scaling=True):
        if config["loss"]["iterative_mode"] == "four":
            config["data"]["passes_loss"] *= 2          # upstream mutates the caller's config (:422-423)
        super().__init__(config, device, loss_scaling=loss_scaling)
        mode = config["loss"]["iterative_mode"]
        self.delta_passes = []
        for passes in self.passes_loss:
            if mode in _MODES:
                self.delta_passes.append(passes // _MODES[mode])

    def update(self, flow_list, event_list, pol_mask, d_event_list, d_pol_mask):
        """Same contract as upstream ``Iterative.update`` (:443-476), including the in-place
        ``event_list[:, :, 0] += num_passes`` on the caller's tensors."""
        self.update_base(flow_list)
        self._update_pass(flow_list, event_list, pol_mask, d_event_list, d_pol_mask)
        self._passes += 1

    def forward(self):
        if self.config["loss"]["iterative_mode"] not in _MODES:
            raise IndexError("unknown iterative_mode %r" % (self.config["loss"]["iterative_mode"],))
        return self._smoothing_terms(self._cm_loss())


class Linear(BaseEventWarping):
    """CM loss of Hagenaars and Paredes-Valles et al. (NeurIPS 2021) with linear warping
    (upstream ``loss/flow.py:216-412``)."""

    _linear = True

    def __init__(self, config, device, loss_scaling=True):
        super().__init__(config, device, loss_scaling=loss_scaling)

    def update(self, flow_list, event_list, pol_mask, d_event_list, d_pol_mask):
        """Same contract as upstream ``Linear.update`` (:233-288).  Upstream samples the per-event flow here; the
        kernels sample the same values from this pass' packed map inside `forward`."""
        self.update_base(flow_list)
        self._update_pass(flow_list, event_list, pol_mask, d_event_list, d_pol_mask)
        self._passes += 1

    def forward(self):
        return self._smoothing_terms(self._cm_loss())


# BASELINE.json's north_star calls the loss "EventWarping"; upstream has no such class
# (SURVEY.md §0).  The trained/default configuration is Iterative, mode "two".
EventWarping = Iterative
