"""CPU, build container only: the oracle's restatement of utils/iwe.py and dataloader/encodings.py against the UNMODIFIED
reference functions imported live, on seeded random shapes (odd resolutions, empty inputs, locations outside the sensor,
exact-integer and half-integer positions).  Per-event outputs bit-identical; images bit-identical too (same CPU
summation order).  Skipped where the reference is not mounted."""
import os
import sys

import numpy as np
import pytest
import torch

REF = os.environ.get("TEF_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "utils")), reason="reference not mounted")

from oracle import cm_oracle as orc  # noqa: E402


def _ref():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from dataloader import encodings as ref_enc  # reference
    from utils import iwe as ref_iwe  # reference

    return ref_iwe, ref_enc


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def same(a, b):
    a = a.detach().numpy() if torch.is_tensor(a) else a
    return a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize("seed", range(30))
def test_primitives(seed):
    ref_iwe, _ = _ref()
    r = np.random.default_rng(300 + seed)
    B, N = int(r.integers(1, 4)), int(r.integers(1, 300))
    H, W = int(r.integers(2, 60)), int(r.integers(2, 75))
    res = (H, W)
    mx, my = (r.normal(0, 3, (B, H, W)).astype(np.float32) for _ in range(2))
    loc = np.stack([r.uniform(-4, H + 3, (B, N)), r.uniform(-4, W + 3, (B, N))], -1).astype(np.float32)
    k = r.integers(0, 4, (B, N))
    loc[k == 1] = np.round(loc[k == 1])
    loc[k == 2] = np.floor(loc[k == 2]) + 0.5
    ts = r.uniform(0, 1, (B, N, 1)).astype(np.float32)
    mask = (r.random((B, N, 2)) < 0.5).astype(np.float32)
    tref = float(r.integers(0, 5))

    flow = ref_iwe.get_event_flow(T(mx), T(my), T(loc))
    o_flow = orc.get_event_flow(mx, my, loc)
    assert same(flow, o_flow)
    warped = ref_iwe.event_propagation(T(ts), T(loc), flow, tref)
    o_warped = orc.event_propagation(ts, loc, o_flow, tref)
    assert same(warped, o_warped)
    pl, pm = ref_iwe.purge_unfeasible(warped, T(mask), res)
    o_pl, o_pm = orc.purge_unfeasible(o_warped, mask, res)
    assert same(pl, o_pl) and same(pm, o_pm)
    for round_idx in (False, True):
        idx, w = ref_iwe.get_interpolation(pl.clone(), res, round_idx=round_idx)
        o_idx, o_w = orc.get_interpolation(o_pl, res, round_idx=round_idx)
        assert same(idx, o_idx) and same(w, o_w), round_idx
        pol = np.concatenate([o_pm[..., 0:1]] * (1 if round_idx else 4), 1)
        img = ref_iwe.interpolate(idx, w, res, polarity_mask=T(pol))
        o_img = orc.interpolate(o_idx, o_w, res, polarity_mask=pol)
        assert same(img, o_img.reshape(img.shape)), round_idx


@pytest.mark.parametrize("seed", range(15))
def test_encodings_and_deblur(seed):
    ref_iwe, ref_enc = _ref()
    r = np.random.default_rng(900 + seed)
    H, W, n, bins = int(r.integers(2, 50)), int(r.integers(2, 70)), int(r.integers(1, 3000)), int(r.integers(2, 8))
    xs = (r.integers(0, W, n) + (r.random(n) < 0.1) * 0.6).astype(np.float32)
    ys = r.integers(0, H, n).astype(np.float32)
    ts = np.sort(r.random(n)).astype(np.float32)
    ps = (r.integers(0, 2, n) * 2 - 1).astype(np.float32)
    assert same(ref_enc.events_to_channels(T(xs), T(ys), T(ps), sensor_size=(H, W)), orc.events_to_channels(xs, ys, ps, (H, W)))
    assert same(ref_enc.events_to_image(T(xs), T(ys), T(ps), sensor_size=(H, W)), orc.events_to_image(xs, ys, ps, (H, W)))
    assert same(ref_enc.events_to_voxel(T(xs), T(ys), T(ts), T(ps), bins, sensor_size=(H, W)), orc.events_to_voxel(xs, ys, ts, ps, bins, (H, W)))
    B = 2
    ev = np.stack([np.stack([ts, ys, np.floor(xs), ps], -1)] * B)
    ev[1, :, 1:3] = ev[1, ::-1, 1:3]
    flow = r.normal(0, 2, (B, 2, H, W)).astype(np.float32)
    pol = np.stack([(ev[..., 3] > 0), (ev[..., 3] < 0)], -1).astype(np.float32)
    for round_idx in (True, False):
        got = orc.compute_pol_iwe(flow, ev, (H, W), pol, round_idx=round_idx, round_flow=True)
        want = ref_iwe.compute_pol_iwe(T(flow), T(ev), (H, W), T(pol), round_idx=round_idx, round_flow=True)
        assert same(want, got), round_idx


def test_reference_encoding_error_and_last_writer_behaviour():
    """What the CUDA encodings mirror (tests/test_primitives_gpu.py): an event outside the sensor raises IndexError in the
    reference's index_put_ (dataloader/encodings.py:23-27), negative coordinates wrap, and accumulate=False keeps the LAST
    event of a pixel."""
    import torch

    _, ref_enc = _ref()
    for bad_x, bad_y in ((20.0, 0.0), (0.0, 12.0), (-21.0, 0.0), (0.0, -13.0)):
        with pytest.raises(IndexError):
            ref_enc.events_to_image(torch.tensor([bad_x]), torch.tensor([bad_y]), torch.tensor([1.0]), sensor_size=(12, 20))
        for fn in (orc.events_to_image, orc.events_to_channels):
            with pytest.raises(IndexError):
                fn(np.float32([bad_x]), np.float32([bad_y]), np.float32([1.0]), (12, 20))
        with pytest.raises(IndexError):
            orc.events_to_voxel(np.float32([bad_x]), np.float32([bad_y]), np.float32([0.5]), np.float32([1.0]), 3, (12, 20))
    img = ref_enc.events_to_image(torch.tensor([-1.0, 3, 3]), torch.tensor([-12.0, 1, 1]), torch.tensor([1.0, 2, 5]), sensor_size=(12, 20), accumulate=False)
    assert img[0, 19] == 1.0 and img[1, 3] == 5.0
    # the oracle follows: negative coordinates wrap, accumulate=False keeps the last event of a pixel
    r = np.random.default_rng(5)
    n, H, W = 500, 12, 20
    xs = r.integers(-W, W, n).astype(np.float32)
    ys = r.integers(-H, H, n).astype(np.float32)
    ts = np.sort(r.random(n)).astype(np.float32)
    ps = r.normal(size=n).astype(np.float32)
    for acc in (True, False):
        assert same(ref_enc.events_to_image(T(xs), T(ys), T(ps), sensor_size=(H, W), accumulate=acc), orc.events_to_image(xs, ys, ps, (H, W), accumulate=acc))
    assert same(ref_enc.events_to_channels(T(xs), T(ys), T(ps), sensor_size=(H, W)), orc.events_to_channels(xs, ys, ps, (H, W)))
    assert same(ref_enc.events_to_voxel(T(xs), T(ys), T(ts), T(ps), 4, sensor_size=(H, W)), orc.events_to_voxel(xs, ys, ts, ps, 4, (H, W)))
