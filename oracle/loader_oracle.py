"""numpy restatement of the reference's loader -> loss contract (``dataloader/base.py``), SURVEY.md §8f-2.

TEST INFRASTRUCTURE ONLY (same rule as ``cm_oracle.py``): imported by ``tests/`` as the checker for
``taming_event_flow_b200/dataloader/base.py``.  Pinned against the unmodified reference by
``tests/golden/loader.npz`` (``tests/golden/make_golden.py::make_loader_cases``).
"""
import numpy as np


def event_formatting(xs, ys, ts, ps):
    """``base.py:139-170``: fp32 casts, polarity {0,1} -> {-1,+1}, timestamps normalised to [0, 1] in fp32."""
    xs = np.asarray(xs).astype(np.float32)
    ys = np.asarray(ys).astype(np.float32)
    ts = np.asarray(ts).astype(np.float32)
    ps = np.asarray(ps).astype(np.float32) * np.float32(2) - np.float32(1)
    if ts.shape[0] > 0:
        with np.errstate(invalid="ignore", divide="ignore"):
            ts = (ts - ts[0]) / (ts[-1] - ts[0])
    return xs, ys, ts, ps


def create_list_encoding(xs, ys, ts, ps):
    """``base.py:247-262``: ``[4 x N]`` rows (ts, y, x, p)."""
    return np.stack([ts, ys, xs, ps])


def create_polarity_mask(ps):
    """``base.py:264-278``; the sign of zero is kept (row 1 of a positive event is -0.0)."""
    ps = np.asarray(ps, np.float32)
    m = np.stack([ps, ps]).copy()
    m[0][m[0] < 0] = 0
    m[0][m[0] > 0] = 1
    m[1][m[1] < 0] = -1
    m[1][m[1] > 0] = 0
    m[1] *= np.float32(-1)
    return m


def create_mask_encoding(event_cnt):
    """``base.py:302-314``: ``[2 x H x W]`` -> ``[1 x H x W]``."""
    m = np.sum(np.asarray(event_cnt, np.float32), axis=0, keepdims=True)
    m[m > 0.0] = 1.0
    return m


def collate_events(items):
    """``custom_collate`` (``base.py:416-428``) for one event key: list of ``[C x n_i]`` -> ``[B, N, C]`` zero-padded."""
    N = max(it.shape[1] for it in items)
    out = np.zeros((len(items), N, items[0].shape[0]), np.float32)
    for b, it in enumerate(items):
        out[b, : it.shape[1]] = np.asarray(it, np.float32).T
    return out


def unpack_events(packed):
    """Inverse of ``taming_event_flow_b200.dataloader.base.pack_events``: raw (xs, ys, fp32 ts, ps in {0,1})."""
    packed = np.asarray(packed, np.uint64)
    ts = (packed & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.float32)
    hi = (packed >> np.uint64(32)).astype(np.uint32)
    return (hi & 0x3FFF).astype(np.int64), ((hi >> 14) & 0x3FFF).astype(np.int64), ts, ((hi >> 28) & 1).astype(np.int64)


def events_to_channels(xs, ys, ps, sensor_size):
    """``dataloader/encodings.py:59-81`` for in-sensor integer coordinates: positive counts per polarity (exact)."""
    H, W = sensor_size
    out = np.zeros((2, H, W), np.float32)
    xi, yi = np.asarray(xs).astype(np.int64), np.asarray(ys).astype(np.int64)
    ps = np.asarray(ps)
    np.add.at(out[0], (yi[ps > 0], xi[ps > 0]), 1.0)
    np.add.at(out[1], (yi[ps < 0], xi[ps < 0]), 1.0)
    return out


def format_windows(windows, sensor_size):
    """What ``format_windows`` must produce for a list of raw windows ``(xs, ys, ts, ps01)``: the upstream route
    event_formatting -> create_list_encoding / create_polarity_mask / events_to_channels -> custom_collate."""
    lists, masks, cnts = [], [], []
    for xs, ys, ts, ps in windows:
        x, y, t, p = event_formatting(xs, ys, ts, ps)
        lists.append(create_list_encoding(x, y, t, p))
        masks.append(create_polarity_mask(p))
        cnts.append(events_to_channels(x, y, p, sensor_size))
    cnt = np.stack(cnts)
    return {"event_list": collate_events(lists), "event_list_pol_mask": collate_events(masks), "event_cnt": cnt,
            "event_mask": np.stack([create_mask_encoding(c) for c in cnts])}
