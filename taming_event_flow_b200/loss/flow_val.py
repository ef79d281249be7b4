"""Validation criteria (FWL, RSAT, AEE, window images), drop-in for the reference's ``loss/flow_val.py``.

Same classes and methods as upstream (``loss/flow_val.py:12-694``: ``BaseValidation``, ``Linear``, ``Iterative`` with
``update`` / ``reset`` / ``window_events`` / ``window_flow`` / ``window_iwe`` / ``rsat`` / ``fwl`` / ``compute_aee``).
This is the first "next" row of SURVEY.md §8f: evaluation only, no gradients.  The stages of ``Iterative.update`` (all
accumulated events one window forward, the new window back through every map, every older flow map carried forward, the
pixel trajectories) and ``forward_prop_flow`` are one fused kernel each (``csrc/tef_validation.cu``); the images and
metrics run through the CUDA primitives of ``taming_event_flow_b200.utils.iwe``; only the book-keeping around them
(concatenating windows, averaging, variances) is left to torch.  It is pinned by golden vectors made from the unmodified
reference (``tests/golden/make_golden.py``), not by the C oracle.  Batch size 1, like upstream (its index grids are
batch-1, ``loss/flow_val.py:30-38``).
"""
import ctypes

import torch

from .._lib import check, lib, ptr, require_cuda, stream
from ..utils.iwe import event_propagation, get_event_flow, get_interpolation, interpolate, purge_unfeasible

_l = ctypes.c_long
_f = ctypes.c_float


def _cat(old, new, dim=1):
    return new if old is None else torch.cat([old, new], dim=dim)


class _Rows:
    """``[1, n, C]`` per-event rows with amortised O(1) append (capacity doubling) instead of a ``torch.cat`` of the whole
    history per window; `view()` is the upstream tensor."""

    def __init__(self, C, device):
        self.C, self.device, self.buf, self.n = C, device, None, 0

    def reserve(self, extra):
        need = self.n + extra
        cap = 0 if self.buf is None else self.buf.shape[1]
        if need > cap:
            nb = torch.empty((1, max(need, 2 * cap, 4096), self.C), dtype=torch.float32, device=self.device)
            if self.n:
                nb[:, :self.n].copy_(self.buf[:, :self.n])
            self.buf = nb

    def at(self, row):
        return ctypes.c_void_p(self.buf.data_ptr() + 4 * self.C * row)

    def view(self):
        return None if self.buf is None else self.buf[:, :self.n]


class _Stack:
    """``[1, n, H, W]`` per-window maps, same idea."""

    def __init__(self, res, device):
        self.HW, self.res, self.device, self.buf, self.n = res[0] * res[1], res, device, None, 0

    def reserve(self, extra=1):
        need = self.n + extra
        cap = 0 if self.buf is None else self.buf.shape[1]
        if need > cap:
            nb = torch.empty((1, max(need, 2 * cap, 16), self.res[0], self.res[1]), dtype=torch.float32, device=self.device)
            if self.n:
                nb[:, :self.n].copy_(self.buf[:, :self.n])
            self.buf = nb

    def at(self, slot):
        return ctypes.c_void_p(self.buf.data_ptr() + 4 * self.HW * slot)

    def view(self):
        return None if self.buf is None else self.buf[:, :self.n]


class _AppendDesc(ctypes.Structure):
    """Mirror of ``tef_val_append``."""
    _fields_ = [("events", ctypes.c_void_p), ("pol_mask", ctypes.c_void_p), ("n", ctypes.c_long), ("pass_index", ctypes.c_float),
                ("ts_override", ctypes.c_void_p), ("ev_ts", ctypes.c_void_p), ("ev_loc", ctypes.c_void_p), ("ev_mask", ctypes.c_void_p),
                ("fw_ts", ctypes.c_void_p), ("fw_loc", ctypes.c_void_p), ("fw_mask", ctypes.c_void_p), ("bw_loc", ctypes.c_void_p),
                ("bw_mask", ctypes.c_void_p), ("flow", ctypes.c_void_p), ("event_mask", ctypes.c_void_p), ("H", ctypes.c_int), ("W", ctypes.c_int),
                ("map_x", ctypes.c_void_p), ("map_y", ctypes.c_void_p), ("map_e", ctypes.c_void_p), ("prop_x", ctypes.c_void_p),
                ("prop_y", ctypes.c_void_p)]


def _tile4(t):
    return torch.cat([t, t, t, t], dim=1)


class BaseValidation(torch.nn.Module):
    """State shared by both validation flavours (upstream ``BaseValidation``, loss/flow_val.py:12-314)."""

    def __init__(self, config, device):
        super().__init__()
        self.res = config["loader"]["resolution"]
        self.device = device
        self.config = config
        H, W = self.res
        yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        grid = torch.stack([yy, xx], 0).float().unsqueeze(0).to(device)          # [1,2,H,W] (y, x)
        self.indices_map = grid
        self.indices = grid.reshape(1, 2, -1).permute(0, 2, 1).contiguous()      # [1,HW,2]
        self.indices_mask = torch.ones((1, H * W, 1), device=device)
        self._clear_base()

    # ------------------------------------------------------------------ state
    def _clear_base(self):
        self._passes = 0
        dev = self.device
        self._r_ts, self._r_loc, self._r_mask = _Rows(1, dev), _Rows(2, dev), _Rows(2, dev)
        self._s_x, self._s_y, self._s_e = _Stack(self.res, dev), _Stack(self.res, dev), _Stack(self.res, dev)

    # upstream's attribute names, as views of the growing stores
    _event_ts = property(lambda self: self._r_ts.view())
    _event_loc = property(lambda self: self._r_loc.view())
    _event_pol_mask = property(lambda self: self._r_mask.view())
    _flow_maps_x = property(lambda self: self._s_x.view())
    _flow_maps_y = property(lambda self: self._s_y.view())
    _event_mask = property(lambda self: self._s_e.view())

    def reset_base(self):
        self._clear_base()

    @property
    def num_passes(self):
        return self._passes

    def _window_ts(self, event_list):
        ts = event_list[:, :, 0:1].clone()
        if self.config["loss"]["round_ts"]:
            ts[...] = ts.min() + 0.5                                             # loss/flow_val.py:87-88
        return ts

    def update_base(self, flow_list, event_list, pol_mask, event_mask, iterative=None):
        """Append this window's events, finest flow map and event mask (upstream :75-114); the pass index is added to the
        caller's timestamps in place, as upstream does.  One launch (``tef_val_append_window``) fills every list `update`
        appends to -- for the Iterative flavour (`iterative` = its row stores) also the forward / backward lists and the
        propagated-flow stack."""
        require_cuda(flow_list[-1], event_list, pol_mask, event_mask)
        if event_list.shape[0] != 1 or flow_list[-1].shape[0] != 1:
            raise RuntimeError("validation runs with batch size 1 (upstream builds batch-1 index grids, loss/flow_val.py:30-38)")
        H, W = self.res
        n, off, now = event_list.shape[1], self._r_ts.n, self._passes
        if event_list.dim() != 3 or event_list.shape[2] != 4 or tuple(pol_mask.shape) != (1, n, 2):
            raise ValueError("event_list must be [1,N,4] and pol_mask [1,N,2], got %s and %s" % (tuple(event_list.shape), tuple(pol_mask.shape)))
        if tuple(flow_list[-1].shape) != (1, 2, H, W) or event_mask.numel() != H * W:
            raise ValueError("flow must be [1,2,%d,%d] and the event mask [1,1,%d,%d]" % (H, W, H, W))
        keep = []
        d = _AppendDesc()
        if event_list.is_contiguous() and event_list.dtype == torch.float32:
            d.events, d.pass_index = event_list.data_ptr(), float(now)
        else:
            event_list[:, :, 0:1] += now
            ev = event_list.contiguous().float()
            keep.append(ev)
            d.events, d.pass_index = ev.data_ptr(), 0.0
        if self.config["loss"]["round_ts"] and n > 0:
            # ts.min() + 0.5 of the timestamps AFTER the in-place update (loss/flow_val.py:86-88): two fp32 roundings, like upstream
            ov = ((event_list[:, :, 0].min().float() + d.pass_index) + 0.5).reshape(1)
            keep.append(ov)
            d.ts_override = ov.data_ptr()
        mk = pol_mask if (pol_mask.is_contiguous() and pol_mask.dtype == torch.float32) else pol_mask.contiguous().float()
        fl = flow_list[-1] if (flow_list[-1].is_contiguous() and flow_list[-1].dtype == torch.float32) else flow_list[-1].contiguous().float()
        em = event_mask if (event_mask.is_contiguous() and event_mask.dtype == torch.float32) else event_mask.contiguous().float()
        keep += [mk, fl, em]
        rows = [self._r_ts, self._r_loc, self._r_mask] + (list(iterative["rows"]) if iterative else [])
        stacks = [self._s_x, self._s_y, self._s_e] + (list(iterative["stacks"]) if iterative else [])
        for r in rows:
            r.reserve(n)
        for st in stacks:
            st.reserve(1)
        d.pol_mask, d.n, d.flow, d.event_mask, d.H, d.W = mk.data_ptr(), n, fl.data_ptr(), em.data_ptr(), H, W
        d.ev_ts, d.ev_loc, d.ev_mask = self._r_ts.at(off), self._r_loc.at(off), self._r_mask.at(off)
        d.map_x, d.map_y, d.map_e = self._s_x.at(now), self._s_y.at(now), self._s_e.at(now)
        if iterative:
            fw_ts, fw_loc, fw_mask, bw_loc, bw_mask = iterative["rows"]
            d.fw_ts, d.fw_loc, d.fw_mask, d.bw_loc, d.bw_mask = fw_ts.at(off), fw_loc.at(off), fw_mask.at(off), bw_loc.at(off), bw_mask.at(off)
            d.prop_x, d.prop_y = iterative["stacks"][0].at(now), iterative["stacks"][1].at(now)
        check(lib().tef_val_append_window(ctypes.byref(d), stream()), "tef_val_append_window")
        for r in rows:
            r.n += n
        for st in stacks:
            st.n += 1
        return off, n

    # ------------------------------------------------------- building blocks
    def _pol_images(self, loc, pol_mask, round_idx, extra=None):
        """Per-polarity images of (optionally weighted) events at `loc`: [B,2,H,W].  Batch 1 with contiguous fp32 inputs (what
        the criteria hold) is one fused launch; anything else goes through the stand-alone operators, same arithmetic."""
        if (loc.shape[0] == 1 and loc.is_contiguous() and pol_mask.is_contiguous() and loc.dtype == torch.float32 and pol_mask.dtype == torch.float32
                and pol_mask.shape[-1] == 2 and (extra is None or (extra.is_contiguous() and extra.dtype == torch.float32 and extra.numel() == loc.shape[1]))):
            require_cuda(loc, pol_mask, extra)
            H, W = self.res
            out = torch.empty((1, 2, H, W), dtype=torch.float32, device=loc.device)
            check(lib().tef_val_pol_images(ptr(loc), ptr(pol_mask), ptr(extra), ptr(out), _l(loc.shape[1]), H, W, int(bool(round_idx)), stream()),
                  "tef_val_pol_images")
            return out
        idx, w = get_interpolation(loc, self.res, round_idx=round_idx)
        pm = pol_mask if round_idx else _tile4(pol_mask)
        if extra is not None:
            w = w * extra
        return torch.cat([interpolate(idx, w, self.res, polarity_mask=pm[:, :, c:c + 1]) for c in range(2)], dim=1)

    def _prop_flow(self, maps_x, maps_y, first, n_maps, tref, out_x, out_y):
        """Maps first..first+n_maps-1 of [1,P,H,W] carried to `tref` (None: each to the next time index), one splat and
        one normalisation kernel for all of them; out_x/out_y [1,P,H,W] may be the inputs."""
        H, W = self.res
        acc = torch.empty((n_maps, 3, H, W), dtype=torch.float32, device=maps_x.device)
        off = first * H * W * 4
        check(lib().tef_val_forward_prop_flow(ctypes.c_void_p(maps_x.data_ptr() + off), ctypes.c_void_p(maps_y.data_ptr() + off), first, n_maps,
                                              int(tref is None), _f(0.0 if tref is None else float(tref)), ptr(acc),
                                              ctypes.c_void_p(out_x.data_ptr() + off), ctypes.c_void_p(out_y.data_ptr() + off), H, W, stream()),
              "tef_val_forward_prop_flow")

    def forward_prop_flow(self, i, tref, flow_maps_x, flow_maps_y):
        """Flow map `i` carried forward to time `tref` by splatting it along itself (upstream :43-74)."""
        mx, my = flow_maps_x.contiguous().float(), flow_maps_y.contiguous().float()
        require_cuda(mx, my)
        fx, fy = torch.empty_like(mx), torch.empty_like(my)
        self._prop_flow(mx, my, i, 1, tref, fx, fy)
        return fx[:, i:i + 1], fy[:, i:i + 1]

    def window_events_base(self, round_idx=False):
        return self._pol_images(self._event_loc, self._event_pol_mask, round_idx)

    def window_flow_base(self, flow_maps_x, flow_maps_y, mask=False):
        """Per-pixel average of the non-zero flow vectors of the window (upstream :146-172)."""
        total = torch.cat([flow_maps_x[:, 0:1], flow_maps_y[:, 0:1]], dim=1)
        cnt = ((flow_maps_x[:, 0:1] != 0.0) | (flow_maps_y[:, 0:1] != 0.0)).float()
        for i in range(1, flow_maps_x.shape[1]):                                 # same left-to-right sums as upstream
            total = total + torch.cat([flow_maps_x[:, i:i + 1], flow_maps_y[:, i:i + 1]], dim=1)
            cnt = cnt + ((flow_maps_x[:, i:i + 1] != 0.0) | (flow_maps_y[:, i:i + 1] != 0.0)).float()
        if mask:
            total = total * (self._event_mask.sum(1, keepdim=True) > 0.0).float()
        return total / (cnt + 1e-9)

    def window_iwe_base(self, round_idx=False):
        warped = event_propagation(self._event_ts, self._event_loc, self._event_flow, self._passes)
        return self._pol_images(warped, self._event_pol_mask, round_idx)

    # ---------------------------------------------------------------- metrics
    def compute_fwl(self, fw_events, zero_events, fw_pol_mask, zero_pol_mask):
        """Flow Warp Loss: variance of the image of warped events over that of the raw events (upstream :189-212)."""
        fw = self._pol_images(fw_events, fw_pol_mask, True).sum(1, keepdim=True)
        zero = self._pol_images(zero_events, zero_pol_mask, True).sum(1, keepdim=True)
        return fw.var() / zero.var()

    def _sat(self, loc, pol_mask, ts_list):
        cnt = self._pol_images(loc, pol_mask, True)
        avg = self._pol_images(loc, pol_mask, True, extra=ts_list) / (cnt + 1e-9) / self._passes
        sq = (avg.flatten(2) ** 2).sum(2).sum(1)
        nonzero = (cnt.sum(1) > 0).flatten(1).float().sum(1)
        return sq / nonzero

    def compute_rsat(self, fw_events, zero_events, fw_pol_mask, zero_pol_mask, ts_list):
        """Ratio of the squared averaged timestamps, warped over raw events (upstream :214-274)."""
        return self._sat(fw_events, fw_pol_mask, ts_list) / self._sat(zero_events, zero_pol_mask, ts_list)

    def compute_aee(self, pred, gt, mask=None):
        """Average endpoint error over pixels with ground truth (and, optionally, with events), upstream :276-314."""
        err = (pred - gt).pow(2).sum(1).sqrt()
        valid = ~((gt[:, 0] == 0.0) & (gt[:, 1] == 0.0))
        if mask is not None:
            has_events = mask.sum(1) > 0
            metrics = self.config["metrics"]
            if "res_aee" in metrics.keys():
                yo = (self.res[0] - metrics["res_aee"][0]) // 2
                xo = (self.res[1] - metrics["res_aee"][1]) // 2
                has_events, err, valid = (t[:, yo:-yo, xo:-xo].contiguous() for t in (has_events, err, valid))
            if "vertical_crop_aee" in metrics.keys():
                rows = metrics["vertical_crop_aee"]
                has_events, err, valid = has_events[:, :rows], err[:, :rows], valid[:, :rows]
            valid = valid & has_events
        return err.flatten(1)[valid.flatten(1)].mean(0)


class Linear(BaseValidation):
    """Linear-warping validation (upstream ``Linear``, loss/flow_val.py:317-416)."""

    def __init__(self, config, device):
        super().__init__(config, device)
        self._event_flow = None

    def update(self, flow_list, event_list, pol_mask, event_mask):
        off, n = self.update_base(flow_list, event_list, pol_mask, event_mask)
        flow = get_event_flow(self._flow_maps_x[:, -1], self._flow_maps_y[:, -1], self._event_loc[:, off:off + n])
        self._event_flow = _cat(self._event_flow, flow)
        self._passes += 1

    def reset(self):
        self.reset_base()
        self._event_flow = None

    def window_events(self, round_idx=False):
        return self.window_events_base(round_idx)

    def window_flow(self, mode=None, mask=None):
        if mask is None:
            mask = self.config["vis"]["mask_output"]
        fx, fy = self._flow_maps_x.clone(), self._flow_maps_y.clone()
        if self._passes > 1:                                                     # every older map carried to the newest time
            self._prop_flow(self._flow_maps_x.contiguous(), self._flow_maps_y.contiguous(), 0, self._passes - 1, self._passes - 1, fx, fy)
        return self.window_flow_base(fx, fy, mask=mask)

    def window_iwe(self, mode=None, round_idx=False):
        return self.window_iwe_base(round_idx)

    def _warped(self):
        return event_propagation(self._event_ts, self._event_loc, self._event_flow, self._passes)

    def rsat(self):
        return self.compute_rsat(self._warped(), self._event_loc, self._event_pol_mask, self._event_pol_mask, self._event_ts)

    def fwl(self):
        return self.compute_fwl(self._warped(), self._event_loc, self._event_pol_mask, self._event_pol_mask)


class Iterative(BaseValidation):
    """Iterative-warping validation (upstream ``Iterative``, loss/flow_val.py:419-694): every `update` pushes all events
    seen so far one window forward with the newest flow map, pulls the new window back to time 0 through all maps, carries
    the older flow maps forward and accumulates the pixel trajectories."""

    def __init__(self, config, device):
        super().__init__(config, device)
        self._clear_iterative()

    def _clear_iterative(self):
        dev = self.device
        self._r_fw_ts, self._r_fw_loc, self._r_fw_mask = _Rows(1, dev), _Rows(2, dev), _Rows(2, dev)
        self._r_bw_loc, self._r_bw_mask = _Rows(2, dev), _Rows(2, dev)
        self._s_px, self._s_py = _Stack(self.res, dev), _Stack(self.res, dev)
        self._accum_flow_map_x = self._accum_flow_map_y = None
        self._flow_warping_indices = None
        self._prop_acc = None
        self._flow_out_mask = torch.zeros(1, 1, self.res[0], self.res[1], device=self.device)

    _fw_event_warp_ts = property(lambda self: self._r_fw_ts.view())
    _fw_event_loc = property(lambda self: self._r_fw_loc.view())
    _fw_event_pol_mask = property(lambda self: self._r_fw_mask.view())
    _bw_event_loc = property(lambda self: self._r_bw_loc.view())
    _bw_event_pol_mask = property(lambda self: self._r_bw_mask.view())
    _fw_prop_flow_maps_x = property(lambda self: self._s_px.view())
    _fw_prop_flow_maps_y = property(lambda self: self._s_py.view())

    def reset(self):
        self.reset_base()
        self._clear_iterative()

    def update(self, flow_list, event_list, pol_mask, event_mask):
        """Upstream ``Iterative.update`` (:477-607) in seven launches: append (all lists and stacks), forward step of every
        event so far, backward chain of the new window, flow propagation (clear, splat, normalise), pixel trajectories."""
        now = self._passes
        H, W = self.res
        L, st = lib(), stream()
        off, n = self.update_base(flow_list, event_list, pol_mask, event_mask,
                                  iterative={"rows": (self._r_fw_ts, self._r_fw_loc, self._r_fw_mask, self._r_bw_loc, self._r_bw_mask),
                                             "stacks": (self._s_px, self._s_py)})
        last_x, last_y = self._s_x.at(now), self._s_y.at(now)

        # all events so far, one window forward with the newest map (upstream :483-517)
        check(L.tef_val_forward_step(last_x, last_y, self._r_fw_loc.at(0), self._r_fw_ts.at(0), self._r_fw_mask.at(0), _f(now + 1), _l(off + n), H, W, st),
              "tef_val_forward_step")

        # the new window, back to time 0 through every map (upstream :519-556): its rows of the backward lists, in place
        check(L.tef_val_backward_chain(self._s_x.at(0), self._s_y.at(0), now + 1, self._r_bw_loc.at(off), self._r_ts.at(off), self._r_bw_mask.at(off),
                                       _l(n), H, W, st), "tef_val_backward_chain")

        # older flow maps carried one window forward (upstream :558-577), in place in the propagated stack
        if now > 0:
            if self._prop_acc is None or self._prop_acc.shape[0] < now:
                self._prop_acc = torch.empty((max(now, 2 * (0 if self._prop_acc is None else self._prop_acc.shape[0]), 16), 3, H, W), dtype=torch.float32,
                                             device=self.device)
            check(L.tef_val_forward_prop_flow(self._s_px.at(0), self._s_py.at(0), 0, now, 1, _f(0.0), ptr(self._prop_acc), self._s_px.at(0), self._s_py.at(0),
                                              H, W, st), "tef_val_forward_prop_flow")

        # pixel trajectories (upstream :579-605)
        if self._flow_warping_indices is None:
            self._flow_warping_indices = self.indices_map.clone()
            self._accum_flow_map_x = torch.empty((1, 1, H, W), dtype=torch.float32, device=self.device)
            self._accum_flow_map_y = torch.empty((1, 1, H, W), dtype=torch.float32, device=self.device)
        check(L.tef_val_trajectory_step(last_x, last_y, ptr(self._flow_warping_indices), ptr(self._flow_out_mask), ptr(self._accum_flow_map_x),
                                        ptr(self._accum_flow_map_y), H, W, st), "tef_val_trajectory_step")
        self._passes += 1

    def window_events(self, round_idx=False):
        return self.window_events_base(round_idx)

    def window_flow(self, mode=None, mask=None):
        if mask is None:
            mask = self.config["vis"]["mask_output"]
        if mode == "forward":
            return self.window_flow_base(self._fw_prop_flow_maps_x, self._fw_prop_flow_maps_y, mask=mask)
        if mode == "backward":
            return self.window_flow_base(self._accum_flow_map_x / self._flow_out_mask, self._accum_flow_map_y / self._flow_out_mask, mask=mask)
        return self.window_flow_base(self._flow_maps_x, self._flow_maps_y, mask=mask)

    def window_iwe(self, mode="forward", round_idx=False):
        if mode == "forward":
            loc, pol = self._fw_event_loc, self._fw_event_pol_mask
        elif mode == "backward":
            loc, pol = self._bw_event_loc, self._bw_event_pol_mask
        else:
            raise ValueError("Invalid IWE mode: {}".format(mode))
        return self._pol_images(loc, pol, round_idx)

    def rsat(self):
        return self.compute_rsat(self._fw_event_loc, self._event_loc, self._fw_event_pol_mask, self._event_pol_mask, self._event_ts)

    def fwl(self):
        return self.compute_fwl(self._fw_event_loc, self._event_loc, self._fw_event_pol_mask, self._event_pol_mask)
