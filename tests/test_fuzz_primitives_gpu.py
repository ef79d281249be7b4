"""Seeded random shapes for the utils/iwe.py and dataloader/encodings.py drop-ins against the CPU oracle: odd resolutions,
tiny and empty inputs, locations far outside the sensor, exact-integer and half-integer positions (ties of round / floor).
Per-event outputs bit-exact; accumulated real-valued images within 1e-5 norm-relative."""
import numpy as np
import pytest
import torch

from oracle import cm_oracle as orc
from taming_event_flow_b200.dataloader import encodings as enc
from taming_event_flow_b200.utils import iwe
from util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def same(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else a
    return a.shape == b.shape and np.array_equal(a, b)


def _locations(r, B, N, H, W):
    loc = np.stack([r.uniform(-4, H + 3, (B, N)), r.uniform(-4, W + 3, (B, N))], -1).astype(np.float32)
    k = r.integers(0, 4, (B, N))
    loc[k == 1] = np.round(loc[k == 1])                              # exact integers (bilinear weight 1 / 0)
    loc[k == 2] = np.floor(loc[k == 2]) + 0.5                        # ties of torch.round (half to even)
    return loc


@pytest.mark.parametrize("seed", range(25))
def test_random_primitives(seed):
    r = np.random.default_rng(100 + seed)
    B, N = int(r.integers(1, 4)), int(r.integers(0, 300))
    H, W = int(r.integers(2, 60)), int(r.integers(2, 75))
    res = (H, W)
    mx, my = (r.normal(0, 3, (B, H, W)).astype(np.float32) for _ in range(2))
    loc = _locations(r, B, N, H, W)
    ts = r.uniform(0, 1, (B, N, 1)).astype(np.float32)
    mask = (r.random((B, N, 2)) < 0.5).astype(np.float32)
    tref = float(r.integers(0, 5))

    flow = iwe.get_event_flow(cu(mx), cu(my), cu(loc))
    o_flow = orc.get_event_flow(mx, my, loc)
    assert same(flow, o_flow)
    warped = iwe.event_propagation(cu(ts), cu(loc), flow, tref)
    o_warped = orc.event_propagation(ts, loc, o_flow, tref)
    assert same(warped, o_warped)
    pl, pm = iwe.purge_unfeasible(warped, cu(mask), res)
    o_pl, o_pm = orc.purge_unfeasible(o_warped, mask, res)
    assert same(pl, o_pl) and same(pm, o_pm)
    for round_idx in (False, True):
        idx, w = iwe.get_interpolation(pl, res, round_idx=round_idx)
        o_idx, o_w = orc.get_interpolation(o_pl, res, round_idx=round_idx)
        assert same(idx, o_idx) and same(w, o_w), round_idx
        pol = np.concatenate([o_pm[..., 0:1]] * (1 if round_idx else 4), 1)
        img = iwe.interpolate(idx, w, res, polarity_mask=cu(pol))
        o_img = orc.interpolate(o_idx, o_w, res, polarity_mask=pol)
        assert img.shape == (B, 1, H, W)
        if round_idx:
            assert same(img, o_img.reshape(B, 1, H, W))               # integer counts
        elif N:
            linf, l2 = rel_err(img.cpu().numpy(), o_img.reshape(B, 1, H, W))
            assert (linf < TOL and l2 < TOL) or not o_img.any()


@pytest.mark.parametrize("seed", range(15))
def test_random_encodings_and_deblur(seed):
    r = np.random.default_rng(500 + seed)
    H, W, n, bins = int(r.integers(1, 50)), int(r.integers(1, 70)), int(r.integers(0, 3000)), int(r.integers(1, 8))
    xs = (r.integers(0, W, n) + (r.random(n) < 0.1) * 0.6).astype(np.float32)      # .long() truncation
    ys = r.integers(0, H, n).astype(np.float32)
    ts = np.sort(r.random(n)).astype(np.float32)
    ps = (r.integers(0, 2, n) * 2 - 1).astype(np.float32)
    assert same(enc.events_to_channels(cu(xs), cu(ys), cu(ps), (H, W)), orc.events_to_channels(xs, ys, ps, (H, W)))
    assert same(enc.events_to_image(cu(xs), cu(ys), cu(ps), (H, W)), orc.events_to_image(xs, ys, ps, (H, W)))   # signed integer sums
    vox, o_vox = enc.events_to_voxel(cu(xs), cu(ys), cu(ts), cu(ps), bins, (H, W)), orc.events_to_voxel(xs, ys, ts, ps, bins, (H, W))
    assert vox.shape == (bins, H, W)
    if o_vox.any():
        linf, l2 = rel_err(vox.cpu().numpy(), o_vox)
        assert linf < TOL and l2 < TOL
    if H >= 2 and W >= 2 and n:
        B = 2
        ev = np.stack([np.stack([ts, ys, np.floor(xs), ps], -1)] * B)             # integer pixels (round_flow indexes with them)
        ev[1, :, 1:3] = ev[1, ::-1, 1:3]
        flow = r.normal(0, 2, (B, 2, H, W)).astype(np.float32)
        pol = np.stack([(ev[..., 3] > 0), (ev[..., 3] < 0)], -1).astype(np.float32)
        for round_idx in (True, False):
            got = iwe.compute_pol_iwe(cu(flow), cu(ev), (H, W), cu(pol), round_idx=round_idx, round_flow=True)
            want = orc.compute_pol_iwe(flow, ev, (H, W), pol, round_idx=round_idx, round_flow=True)
            assert got.shape == want.shape
            if round_idx:
                assert same(got, want)
            elif want.any():
                linf, l2 = rel_err(got.cpu().numpy(), want)
                assert linf < TOL and l2 < TOL
