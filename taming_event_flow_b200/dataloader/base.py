"""Loader -> loss contract, device side (upstream ``dataloader/base.py``).

Upstream formats each window on the device, copies every tensor back to the host (``dataloader/h5.py:409-421``),
zero-pads and stacks there (``custom_collate``, ``base.py:391-434``) and uploads 24 B per event again in the training
loop (``train_flow.py:106-116``).  Two ways in here:

* the upstream static methods under their own names and layouts (``create_list_encoding``, ``create_polarity_mask``,
  ``create_mask_encoding``, ``split_event_list``, ``custom_collate``), each on CUDA tensors;
* ``pack_events`` + ``format_windows``: a window crosses PCIe once as 8-byte packed events and ONE kernel writes the
  padded ``[B, N, 4]`` event list, the ``[B, N, 2]`` polarity mask and the 2-channel count image of a ragged batch —
  exactly the tensors ``Iterative.update`` / the flow network take, bit-identical to the upstream route.

There is no CPU path: tensors must be CUDA tensors (``pack_events`` is host-side by definition: it builds the wire format).
"""
import ctypes

import numpy as np
import torch

from .._lib import check, lib, ptr, require_cuda, stream
from .encodings import events_to_channels

_l = ctypes.c_long

EVENT_KEYS = ("event_list", "event_list_pol_mask", "d_event_list", "d_event_list_pol_mask")
MAX_COORD = 1 << 14


def pack_events(xs, ys, ts, ps):
    """Host side: raw events -> the 8-byte wire format ``{fp32 ts, x | y << 14 | pol << 28}`` (``uint64`` numpy array).

    ``xs, ys``: integer pixel coordinates (< 16384); ``ts``: raw timestamps, converted with ``astype(np.float32)`` like
    upstream ``event_formatting`` (``base.py:163-166``) — normalisation happens on the device; ``ps``: raw polarity in
    {0, 1} (upstream maps it to ``ps * 2 - 1``, ``:167``).
    """
    xs, ys, ps = (np.asarray(a) for a in (xs, ys, ps))
    if not (len(xs) == len(ys) == len(ts) == len(ps)):
        raise AssertionError("xs, ys, ts, ps must have the same length")           # base.py:161
    if len(xs) and (xs.min() < 0 or ys.min() < 0 or xs.max() >= MAX_COORD or ys.max() >= MAX_COORD):
        raise ValueError("event coordinates must lie in [0, 16384)")
    if len(ps) and not np.isin(ps, (0, 1)).all():
        raise ValueError("raw polarities must be 0 or 1")
    t32 = np.asarray(ts).astype(np.float32).view(np.uint32).astype(np.uint64)
    hi = xs.astype(np.uint64) | (ys.astype(np.uint64) << np.uint64(14)) | (ps.astype(np.uint64) << np.uint64(28))
    return t32 | (hi << np.uint64(32))


def create_list_encoding(xs, ys, ts, ps):
    """``[4 x N]`` list representation ``(ts, y, x, p)`` (upstream ``base.py:247-262``)."""
    require_cuda(xs, ys, ts, ps)
    return torch.stack([ts, ys, xs, ps])


def create_polarity_mask(ps):
    """``[2 x N]`` polarity mask: row 0 marks positive, row 1 negative events (upstream ``base.py:264-278``)."""
    require_cuda(ps)
    ps = ps.contiguous().float()
    mask = torch.empty((2, ps.numel()), dtype=torch.float32, device=ps.device)
    check(lib().tef_create_polarity_mask(ptr(ps), ptr(mask), _l(ps.numel()), stream()), "tef_create_polarity_mask")
    return mask


def create_mask_encoding(event_cnt):
    """Per-pixel event mask of a ``[2 x H x W]`` (or batched ``[B x 2 x H x W]``) count image (upstream ``base.py:302-314``)."""
    require_cuda(event_cnt)
    cnt = event_cnt.contiguous().float()
    batched = cnt.dim() == 4
    B = cnt.shape[0] if batched else 1
    H, W = cnt.shape[-2:]
    out = torch.empty((B, 1, H, W), dtype=torch.float32, device=cnt.device)
    check(lib().tef_create_mask_encoding(ptr(cnt), ptr(out), B, H, W, stream()), "tef_create_mask_encoding")
    return out if batched else out[0]


def _draw_seed(generator=None):
    """64-bit key of the split permutation, from torch's RNG (``torch.manual_seed`` reproduces a run)."""
    if generator is not None and generator.device.type != "cpu":
        return int(torch.randint(0, 1 << 62, (1,), device=generator.device, generator=generator).item())   # synchronises
    return int(torch.randint(0, 1 << 62, (1,), generator=generator).item())


def _split_rows(events, masks, counts, k, generator=None, offsets=None):
    """events [B,N,4], masks [B,N,2] (zero-padded), per-sample counts -> grad / detached padded lists."""
    B, N = events.shape[0], events.shape[1]
    dev = events.device
    Ng = max([min(c, k) for c in counts] + [0])
    Nd = max([max(c - k, 0) for c in counts] + [0])
    if offsets is None:
        offsets = torch.tensor([0] + list(np.cumsum(counts)), dtype=torch.int64, device=dev)
    g_ev = torch.empty((B, Ng, 4), dtype=torch.float32, device=dev)
    g_mk = torch.empty((B, Ng, 2), dtype=torch.float32, device=dev)
    d_ev = torch.empty((B, Nd, 4), dtype=torch.float32, device=dev)
    d_mk = torch.empty((B, Nd, 2), dtype=torch.float32, device=dev)
    check(lib().tef_split_events(ptr(events), ptr(masks), ptr(offsets), ctypes.c_ulonglong(_draw_seed(generator)), B, _l(N), _l(k), ptr(g_ev), ptr(g_mk), _l(Ng),
                                 ptr(d_ev), ptr(d_mk), _l(Nd), stream()), "tef_split_events")
    return g_ev, g_mk, d_ev, d_mk


def split_event_list(event_list, event_list_pol_mask, max_num_grad_events, generator=None):
    """Random split into a gradient-carrying list of at most ``max_num_grad_events`` events and a detached rest
    (upstream ``base.py:347-377``; same ``[4 x N]`` / ``[2 x N]`` layouts).  Exactly ``max_num_grad_events`` events are
    sampled without replacement, every event with the same probability, by a keyed pseudo-random permutation computed
    per event (no sort); the key comes from torch's RNG (``generator``).  The detached events come in permutation order
    rather than in their original order (every consumer is order-invariant)."""
    require_cuda(event_list, event_list_pol_mask)
    N = event_list.shape[1]
    if max_num_grad_events is None or N <= max_num_grad_events:
        dev = event_list.device
        return event_list, event_list_pol_mask, torch.zeros((4, 0), device=dev), torch.zeros((2, 0), device=dev)
    rows = _collate([event_list], N, 4)
    mrows = _collate([event_list_pol_mask], N, 2)
    g_ev, g_mk, d_ev, d_mk = _split_rows(rows, mrows, [N], int(max_num_grad_events), generator)
    return g_ev[0].t(), g_mk[0].t(), d_ev[0].t(), d_mk[0].t()


def _collate(items, N, C):
    dev = items[0].device
    out = torch.empty((len(items), N, C), dtype=torch.float32, device=dev)
    for b, it in enumerate(items):
        require_cuda(it)
        if it.shape[0] != C:
            raise RuntimeError("expected a [%d x N] tensor, got %s" % (C, tuple(it.shape)))
        src = it.contiguous().float()
        check(lib().tef_collate_events(ptr(src), ptr(out[b]), _l(src.shape[1]), _l(N), C, stream()), "tef_collate_events")
    return out


def custom_collate(batch):
    """Batch a list of per-sample dictionaries (upstream ``base.py:391-434``): event lists are zero-padded to the longest
    sample and come out as ``[B, N, C]``; every other entry is stacked (3-D results transposed, as upstream does)."""
    out = {}
    for key in batch[0].keys():
        items = [entry[key] for entry in batch]
        if items[0] is None:
            out[key] = None
        elif key in EVENT_KEYS:
            out[key] = _collate(items, max(it.shape[1] for it in items), items[0].shape[0])
        else:
            item = torch.stack(items)
            out[key] = item.transpose(2, 1) if item.dim() == 3 else item
    return out


class BaseDataLoader:
    """Name-compatible holder of the static methods above (upstream keeps them on ``BaseDataLoader``)."""
    create_list_encoding = staticmethod(create_list_encoding)
    create_polarity_mask = staticmethod(create_polarity_mask)
    create_mask_encoding = staticmethod(create_mask_encoding)
    split_event_list = staticmethod(split_event_list)
    custom_collate = staticmethod(custom_collate)

    @staticmethod
    def create_cnt_encoding(xs, ys, ps, sensor_size):
        """Per-pixel, per-polarity event counts (upstream ``base.py:280-300`` without rectification)."""
        return events_to_channels(xs, ys, ps, sensor_size=sensor_size)


class PackedBatch:
    """A ragged batch of packed windows in one pinned host buffer; ``upload`` is a single asynchronous copy."""

    def __init__(self, windows):
        self.counts = [int(len(w)) for w in windows]
        total = sum(self.counts)
        self.host = torch.empty((max(total, 1),), dtype=torch.int64).pin_memory()
        if total:
            self.host[:total] = torch.from_numpy(np.concatenate([np.asarray(w, dtype=np.uint64) for w in windows]).view(np.int64))
        self.offsets_host = torch.tensor([0] + list(np.cumsum(self.counts)), dtype=torch.int64).pin_memory()
        self.nbytes = 8 * total + 8 * len(self.offsets_host)

    def upload(self, device):
        return self.host.to(device, non_blocking=True), self.offsets_host.to(device, non_blocking=True)


def format_windows(windows, sensor_size, device, max_num_grad_events=None, with_cnt=True, generator=None, uploaded=None):
    """Packed windows (``pack_events`` arrays, one per sample, or a ``PackedBatch``) -> the batch dictionary upstream's
    ``custom_collate`` hands to the training loop, already on ``device``:

    ``event_list [B,N,4]``, ``event_list_pol_mask [B,N,2]``, ``d_event_list``, ``d_event_list_pol_mask`` (empty unless
    ``max_num_grad_events`` splits), ``event_cnt [B,2,H,W]`` and ``event_mask [B,1,H,W]`` (``with_cnt``).
    ``uploaded``: the result of an earlier ``PackedBatch.upload`` (a prefetching loader issues it on its copy stream).
    """
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("format_windows needs a CUDA device (no CPU path)")
    batch = windows if isinstance(windows, PackedBatch) else PackedBatch(windows)
    B, N = len(batch.counts), max(batch.counts + [0])
    H, W = int(sensor_size[0]), int(sensor_size[1])
    with torch.cuda.device(device):
        packed, offsets = batch.upload(device) if uploaded is None else uploaded
        ev = torch.empty((B, N, 4), dtype=torch.float32, device=device)
        mk = torch.empty((B, N, 2), dtype=torch.float32, device=device)
        cnt = torch.empty((B, 2, H, W), dtype=torch.float32, device=device) if with_cnt else None
        check(lib().tef_format_events(ptr(packed), ptr(offsets), B, _l(N), ptr(ev), ptr(mk), ptr(cnt) if with_cnt else None, H, W, stream()),
              "tef_format_events")
        out = {"event_list": ev, "event_list_pol_mask": mk,
               "d_event_list": torch.zeros((B, 0, 4), device=device), "d_event_list_pol_mask": torch.zeros((B, 0, 2), device=device)}
        if max_num_grad_events is not None and N > max_num_grad_events:
            g_ev, g_mk, d_ev, d_mk = _split_rows(ev, mk, batch.counts, int(max_num_grad_events), generator, offsets)
            out.update(event_list=g_ev, event_list_pol_mask=g_mk, d_event_list=d_ev, d_event_list_pol_mask=d_mk)
        if with_cnt:
            out["event_cnt"] = cnt
            out["event_mask"] = create_mask_encoding(cnt)
    return out
