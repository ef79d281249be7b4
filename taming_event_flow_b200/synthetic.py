"""Seeded synthetic event streams and flow fields.

Used by the parity tests, the golden-vector generator and ``bench.py`` so that
every implementation sees identical bits.  Everything is generated on the CPU
with ``torch.manual_seed`` and copied to the device afterwards.

Event tensors follow the loader contract of the reference
(``dataloader/base.py:147-171,264-278,391-434``): ``[B,N,4]`` rows of
``(ts, y, x, p)`` with ``ts`` normalised to ``[0,1]`` per window, integer pixel
coordinates stored as fp32, ``p`` in ``{-1,+1}``; a ``[B,N,2]`` ``(pos,neg)`` mask;
shorter samples are zero-padded to the batch maximum.
"""
import math

import torch
import torch.nn.functional as F


def make_window(gen, B, N, H, W, ragged=False, distribution="uniform", t_index=0):
    """One window of events for a batch: returns ``events [B,N,4]`` and ``pol_mask [B,N,2]``."""
    events = torch.zeros(B, N, 4)
    mask = torch.zeros(B, N, 2)
    for b in range(B):
        n = N
        if ragged and N > 0:
            n = int(torch.randint(max(1, int(0.3 * N)), N + 1, (1,), generator=gen))
        if n == 0:
            continue
        ts, _ = torch.sort(torch.rand(n, generator=gen))
        if n > 1:
            ts = (ts - ts[0]) / (ts[-1] - ts[0])
        else:
            ts = torch.zeros(1)
        if distribution == "uniform":
            y = torch.randint(0, H, (n,), generator=gen).float()
            x = torch.randint(0, W, (n,), generator=gen).float()
        elif distribution == "edges":
            # events on 32 moving line segments: heavy pixel reuse (atomic contention)
            nseg = 32
            seg = torch.randint(0, nseg, (n,), generator=gen)
            g2 = torch.Generator().manual_seed(1234 + b)
            c = torch.rand(nseg, 2, generator=g2) * torch.tensor([H - 1.0, W - 1.0])
            ang = torch.rand(nseg, generator=g2) * math.pi
            length = 0.25 * min(H, W)
            vel = (torch.rand(nseg, 2, generator=g2) - 0.5) * 6.0
            u = (torch.rand(n, generator=gen) - 0.5) * length
            y = c[seg, 0] + u * torch.sin(ang[seg]) + vel[seg, 0] * (t_index + ts)
            x = c[seg, 1] + u * torch.cos(ang[seg]) + vel[seg, 1] * (t_index + ts)
            y = y.round().clamp(0, H - 1)
            x = x.round().clamp(0, W - 1)
        else:
            raise ValueError(distribution)
        p = (torch.randint(0, 2, (n,), generator=gen) * 2 - 1).float()
        events[b, :n, 0] = ts
        events[b, :n, 1] = y
        events[b, :n, 2] = x
        events[b, :n, 3] = p
        mask[b, :n, 0] = (p > 0).float()
        mask[b, :n, 1] = (p < 0).float()
    return events, mask


def make_flow(gen, B, H, W, sigma=3.0, coarse=16):
    """Smooth flow field ``[B,2,H,W]`` (ch0 = x, ch1 = y): bicubic upsample of N(0, sigma^2)."""
    h, w = max(2, H // coarse), max(2, W // coarse)
    z = torch.randn(B, 2, h, w, generator=gen) * sigma
    return F.interpolate(z, size=(H, W), mode="bicubic", align_corners=True).contiguous()


def make_sequence(seed, B, P, N, Nd, H, W, F_scales=1, sigma=3.0, ragged=False, distribution="uniform"):
    """A full loss window: ``P`` passes of flows, events and detached events.

    Returns a dict of lists indexed by pass: ``flows[t][f]``, ``events[t]``, ``masks[t]``,
    ``d_events[t]``, ``d_masks[t]``.
    """
    gen = torch.Generator().manual_seed(seed)
    out = {"flows": [], "events": [], "masks": [], "d_events": [], "d_masks": []}
    for t in range(P):
        out["flows"].append([make_flow(gen, B, H, W, sigma) for _ in range(F_scales)])
        ev, mk = make_window(gen, B, N, H, W, ragged, distribution, t)
        dev, dmk = make_window(gen, B, Nd, H, W, ragged, distribution, t)
        out["events"].append(ev)
        out["masks"].append(mk)
        out["d_events"].append(dev)
        out["d_masks"].append(dmk)
    return out


def loss_config(H, W, B, P=10, scales_loss=1, iterative_mode="two", round_ts=False, warping="Iterative"):
    """The nested dict the reference's loss classes read (``configs/train_flow.yml``)."""
    return {
        "loader": {"resolution": [H, W], "batch_size": B},
        "loss": {
            "warping": warping,
            "iterative_mode": iterative_mode,
            "round_ts": round_ts,
            "flow_spat_smooth_weight": None,
            "flow_temp_smooth_weight": None,
        },
        "data": {"passes_loss": P, "scales_loss": scales_loss},
    }
