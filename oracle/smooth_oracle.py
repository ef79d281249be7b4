"""numpy restatement of the reference's smoothness priors (``loss/flow.py:131-209``), SURVEY.md §8f-3.

TEST INFRASTRUCTURE ONLY (same rule as ``cm_oracle.py``).  Values only; the gradients of the CUDA kernels are pinned by
the reference's own autograd (``tests/golden/smoothness.npz``).  ``flows[t][f]``: ``[B,2,H,W]`` (ch0 = x, ch1 = y).
"""
import numpy as np


def _stack(flows, f, dtype):
    return np.stack([np.asarray(per[f], dtype) for per in flows], 1)                # [B,P,2,H,W]


def flow_spatial_smoothing(flows, dtype=np.float64):
    """``loss/flow.py:170-209`` with weight 1: Charbonnier (eps 1e-6) of the four directional differences."""
    F = len(flows[0])
    total = 0
    for f in range(F):
        fl = _stack(flows, f, dtype)
        B, P = fl.shape[:2]
        pairs = ((fl[..., :, :-1], fl[..., :, 1:]), (fl[..., :-1, :], fl[..., 1:, :]),
                 (fl[..., :-1, :-1], fl[..., 1:, 1:]), (fl[..., 1:, :-1], fl[..., :-1, 1:]))
        acc = 0
        for a, b in pairs:
            d = np.sqrt((a - b) ** 2 + dtype(1e-6)).sum(2)                          # x and y components
            acc = acc + d.reshape(B, P, -1).mean(2).mean(1)
        total = total + acc / 4
    return (total / F).sum()


def _sample(mp, y, x):
    """bilinear sample of mp [H,W] at (y, x) arrays, zeros outside (grid_sample, align_corners=True)."""
    H, W = mp.shape
    y0, x0 = np.floor(y).astype(np.int64), np.floor(x).astype(np.int64)
    ay, ax = y - y0, x - x0
    out = 0
    for dy, dx, w in ((0, 0, (1 - ay) * (1 - ax)), (0, 1, (1 - ay) * ax), (1, 0, ay * (1 - ax)), (1, 1, ay * ax)):
        yy, xx = y0 + dy, x0 + dx
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        out = out + np.where(ok, mp[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], 0) * w
    return out


def flow_temporal_smoothing(flows, dtype=np.float64):
    """``loss/flow.py:131-168`` with weight 1: each map against the next one sampled where its flow points."""
    F, P = len(flows[0]), len(flows)
    total = 0
    for f in range(F):
        fl = _stack(flows, f, dtype)
        B, _, _, H, W = fl.shape
        yy, xx = np.meshgrid(np.arange(H, dtype=dtype), np.arange(W, dtype=dtype), indexing="ij")
        for j in range(P - 1):
            for b in range(B):
                vx, vy = fl[b, j, 0], fl[b, j, 1]
                ty, tx = yy + vy, xx + vx
                inside = ((ty >= 0) & (ty <= H - 1) & (tx >= 0) & (tx <= W - 1)).astype(dtype)
                ny, nx = _sample(fl[b, j + 1, 1], ty, tx), _sample(fl[b, j + 1, 0], ty, tx)
                d = np.sqrt((vy - ny) ** 2 + dtype(1e-9)) + np.sqrt((vx - nx) ** 2 + dtype(1e-9))
                total = total + (d * inside).sum() / (inside.sum() + dtype(1e-9))
    return total / F / (P - 1)
