"""Event encodings, drop-in for the reference's ``dataloader/encodings.py``.

Same signatures as upstream (``dataloader/encodings.py:8-81``): 1-D coordinate / timestamp / polarity
tensors in, image-like tensors out; each call is one CUDA kernel (one pass over the events, also
for the voxel grid, where upstream makes ``num_bins`` passes).  Event counts are exact.
"""
import ctypes

import torch

from .._lib import check, lib, ptr, require_cuda, stream

_l = ctypes.c_long

# Upstream's ``index_put_`` raises IndexError for an event outside the sensor (dataloader/encodings.py:23-27).  Mirroring that
# needs the kernel's out-of-bounds flag on the host, i.e. one 4-byte read-back (a stream synchronisation) per call.  Set to
# False to skip the read-back: such events are then dropped silently.
CHECK_BOUNDS = True


def _prep(*ts):
    require_cuda(*ts)
    return [t.contiguous().float() for t in ts]


def _oob_flag(device):
    return torch.zeros((1,), dtype=torch.int32, device=device) if CHECK_BOUNDS else None


def _raise_if_oob(flag, sensor_size):
    if flag is not None and int(flag.item()) != 0:
        raise IndexError("event coordinates out of bounds for a sensor of size %s (upstream: index_put_ raises IndexError)" % (tuple(sensor_size),))


def events_to_image(xs, ys, ps, sensor_size=(180, 240), accumulate=True):
    """Accumulate events into an image (upstream ``dataloader/encodings.py:8-29``).

    Coordinates are truncated like ``.long()``; negative ones wrap like Python indexing; events beyond the sensor raise
    ``IndexError`` like upstream (see `CHECK_BOUNDS`).  ``accumulate=False`` keeps the last event of every pixel, the result of
    upstream's CPU ``index_put_`` (deterministic here: highest event index per pixel).
    """
    xs, ys, ps = _prep(xs, ys, ps)
    H, W = int(sensor_size[0]), int(sensor_size[1])
    img = torch.empty((H, W), dtype=torch.float32, device=xs.device)
    oob = _oob_flag(xs.device)
    check(lib().tef_events_to_image(ptr(xs), ptr(ys), ptr(ps), ptr(img), _l(xs.numel()), H, W, int(bool(accumulate)), ptr(oob), stream()),
          "tef_events_to_image")
    _raise_if_oob(oob, sensor_size)
    return img


def events_to_voxel(xs, ys, ts, ps, num_bins, sensor_size=(180, 240)):
    """Voxel grid with temporal bilinear interpolation (upstream ``dataloader/encodings.py:32-56``): [num_bins x H x W]."""
    assert len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps)
    xs, ys, ts, ps = _prep(xs, ys, ts, ps)
    H, W = int(sensor_size[0]), int(sensor_size[1])
    out = torch.empty((int(num_bins), H, W), dtype=torch.float32, device=xs.device)
    oob = _oob_flag(xs.device)
    check(lib().tef_events_to_voxel(ptr(xs), ptr(ys), ptr(ts), ptr(ps), ptr(out), _l(xs.numel()), int(num_bins), H, W, ptr(oob), stream()),
          "tef_events_to_voxel")
    _raise_if_oob(oob, sensor_size)
    return out


def events_to_channels(xs, ys, ps, sensor_size=(180, 240)):
    """Two-channel per-polarity event counts (upstream ``dataloader/encodings.py:59-81``): [2 x H x W], both positive."""
    assert len(xs) == len(ys) and len(ys) == len(ps)
    xs, ys, ps = _prep(xs, ys, ps)
    H, W = int(sensor_size[0]), int(sensor_size[1])
    out = torch.empty((2, H, W), dtype=torch.float32, device=xs.device)
    oob = _oob_flag(xs.device)
    check(lib().tef_events_to_channels(ptr(xs), ptr(ys), ptr(ps), ptr(out), _l(xs.numel()), H, W, ptr(oob), stream()), "tef_events_to_channels")
    _raise_if_oob(oob, sensor_size)
    return out


def events_to_channels_batched(event_list, sensor_size):
    """``events_to_channels`` for a whole zero-padded batch ``[B x N x 4]`` of (ts, y, x, p) rows in one launch:
    ``[B x 2 x H x W]``.  Not an upstream function: it replaces the per-sample loop + host round trip of the loader
    (upstream ``dataloader/base.py:164-167,291``); padding rows (p = 0) add nothing."""
    require_cuda(event_list)
    ev = event_list.contiguous().float()
    B, N = ev.shape[0], ev.shape[1]
    H, W = int(sensor_size[0]), int(sensor_size[1])
    out = torch.empty((B, 2, H, W), dtype=torch.float32, device=ev.device)
    check(lib().tef_events_to_channels_batched(ptr(ev), ptr(out), B, N, H, W, stream()), "tef_events_to_channels_batched")
    return out


def get_hot_event_mask(event_rate, idx, max_px=100, min_obvs=5, max_rate=0.8):
    """Binary mask that removes hot pixels.  **Not part of tudelft/taming_event_flow** (``BASELINE.json`` names it;
    SURVEY.md §0) -- parity is unpinned.  Specification: the routine of the same name in tudelft/event_flow's
    ``dataloader/encodings.py``: if ``idx > min_obvs``, up to ``max_px`` times take the arg-max of ``event_rate`` and, while it
    exceeds ``max_rate``, zero it (in place, as there) and clear the mask there.  Returns the ``[H x W]`` mask of ones/zeros."""
    require_cuda(event_rate)
    if not (event_rate.is_contiguous() and event_rate.dtype == torch.float32 and event_rate.dim() == 2):
        raise ValueError("event_rate must be a contiguous float32 [H, W] tensor (it is updated in place)")
    H, W = event_rate.shape
    mask = torch.empty_like(event_rate)
    check(lib().tef_get_hot_event_mask(ptr(event_rate), ptr(mask), H, W, int(idx), int(max_px), int(min_obvs), ctypes.c_float(max_rate), stream()),
          "tef_get_hot_event_mask")
    return mask
