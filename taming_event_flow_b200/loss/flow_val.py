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
        self._event_ts = self._event_loc = self._event_pol_mask = None
        self._flow_maps_x = self._flow_maps_y = self._event_mask = None

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

    def update_base(self, flow_list, event_list, pol_mask, event_mask):
        """Append this window's events, finest flow map and event mask (upstream :75-114); the pass index is added to the
        caller's timestamps in place, as upstream does."""
        require_cuda(flow_list[-1], event_list, pol_mask)
        if event_list.shape[0] != 1 or flow_list[-1].shape[0] != 1:
            raise RuntimeError("validation runs with batch size 1 (upstream builds batch-1 index grids, loss/flow_val.py:30-38)")
        event_list[:, :, 0:1] += self._passes
        self._event_ts = _cat(self._event_ts, self._window_ts(event_list))
        self._event_loc = _cat(self._event_loc, event_list[:, :, 1:3].clone())
        self._event_pol_mask = _cat(self._event_pol_mask, pol_mask.clone())
        flow = flow_list[-1]
        self._flow_maps_x = _cat(self._flow_maps_x, flow[:, 0:1])
        self._flow_maps_y = _cat(self._flow_maps_y, flow[:, 1:2])
        self._event_mask = _cat(self._event_mask, event_mask)

    # ------------------------------------------------------- building blocks
    def _pol_images(self, loc, pol_mask, round_idx, extra=None):
        """Per-polarity images of (optionally weighted) events at `loc`: [B,2,H,W]."""
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
        self.update_base(flow_list, event_list, pol_mask, event_mask)
        flow = get_event_flow(self._flow_maps_x[:, -1], self._flow_maps_y[:, -1], event_list[:, :, 1:3])
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
        self._fw_event_loc = self._fw_event_warp_ts = self._fw_event_pol_mask = None
        self._bw_event_loc = self._bw_event_pol_mask = None
        self._fw_prop_flow_maps_x = self._fw_prop_flow_maps_y = None
        self._accum_flow_map_x = self._accum_flow_map_y = None
        self._flow_warping_indices = None
        self._flow_out_mask = torch.zeros(1, 1, self.res[0], self.res[1], device=self.device)

    def reset(self):
        self.reset_base()
        self._clear_iterative()

    def update_fw_event_lists(self, event_list, event_pol_mask):
        self._fw_event_warp_ts = _cat(self._fw_event_warp_ts, self._window_ts(event_list)).contiguous()
        self._fw_event_loc = _cat(self._fw_event_loc, event_list[:, :, 1:3].clone()).contiguous()
        self._fw_event_pol_mask = _cat(self._fw_event_pol_mask, event_pol_mask.float().clone()).contiguous()

    def update_bw_event_lists(self, event_loc, event_pol_mask):
        self._bw_event_loc = _cat(self._bw_event_loc, event_loc.clone())
        self._bw_event_pol_mask = _cat(self._bw_event_pol_mask, event_pol_mask.clone())

    def update(self, flow_list, event_list, pol_mask, event_mask):
        self.update_base(flow_list, event_list, pol_mask, event_mask)
        now = self._passes
        H, W = self.res
        L, st = lib(), stream()
        maps_x, maps_y = self._flow_maps_x.contiguous(), self._flow_maps_y.contiguous()          # [1,now+1,H,W]
        last = now * H * W * 4
        last_x, last_y = ctypes.c_void_p(maps_x.data_ptr() + last), ctypes.c_void_p(maps_y.data_ptr() + last)

        # all events so far, one window forward with the newest map (upstream :483-517)
        self.update_fw_event_lists(event_list, pol_mask)
        check(L.tef_val_forward_step(last_x, last_y, ptr(self._fw_event_loc), ptr(self._fw_event_warp_ts), ptr(self._fw_event_pol_mask),
                                     _f(now + 1), _l(self._fw_event_loc.shape[1]), H, W, st), "tef_val_forward_step")

        # the new window, back to time 0 through every map (upstream :519-556)
        loc = event_list[:, :, 1:3].clone(memory_format=torch.contiguous_format)
        mask = pol_mask.float().clone(memory_format=torch.contiguous_format)
        ts = self._window_ts(event_list).contiguous()
        check(L.tef_val_backward_chain(ptr(maps_x), ptr(maps_y), now + 1, ptr(loc), ptr(ts), ptr(mask), _l(loc.shape[1]), H, W, st),
              "tef_val_backward_chain")
        self.update_bw_event_lists(loc, mask)

        # older flow maps carried one window forward (upstream :558-577)
        newest = flow_list[-1]
        self._fw_prop_flow_maps_x = _cat(self._fw_prop_flow_maps_x, newest[:, 0:1])
        self._fw_prop_flow_maps_y = _cat(self._fw_prop_flow_maps_y, newest[:, 1:2])
        if now > 0:
            self._prop_flow(self._fw_prop_flow_maps_x, self._fw_prop_flow_maps_y, 0, now, None, self._fw_prop_flow_maps_x, self._fw_prop_flow_maps_y)

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
