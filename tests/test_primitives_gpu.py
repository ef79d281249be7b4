"""GPU parity of the utils/iwe.py and dataloader/encodings.py drop-ins against the golden vectors made from
the unmodified reference, and against the CPU oracle on larger seeded inputs.
Per-event outputs (flow samples, positions, indices, weights) and event counts must be bit-exact;
accumulated real-valued images within 1e-5 norm-relative (fp32 summation order)."""
import os

import numpy as np
import pytest
import torch

from oracle import cm_oracle as orc
from util import GOLDEN, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def cu(a, grad=False):
    t = torch.as_tensor(np.asarray(a)).cuda()
    return t.requires_grad_(True) if grad else t


@pytest.fixture(scope="module")
def z():
    return np.load(os.path.join(GOLDEN, "primitives.npz"))


def test_get_event_flow_forward_backward(z):
    from taming_event_flow_b200.utils.iwe import get_event_flow

    mx, my, loc = cu(z["mapx"], True), cu(z["mapy"], True), cu(z["loc"], True)
    out = get_event_flow(mx, my, loc)
    assert np.array_equal(out.detach().cpu().numpy(), z["gef_out"])
    out.backward(cu(z["gef_gout"]))
    for got, ref in ((mx.grad, z["gef_gmapx"]), (my.grad, z["gef_gmapy"]), (loc.grad, z["gef_gloc"])):
        linf, l2 = rel_err(got.cpu().numpy(), ref)
        assert linf < TOL and l2 < TOL, (linf, l2)


def test_event_propagation_and_purge(z):
    from taming_event_flow_b200.utils.iwe import event_propagation, purge_unfeasible

    ts, loc, fl = cu(z["ts"], True), cu(z["loc"], True), cu(z["gef_out"], True)
    out = event_propagation(ts, loc, fl, 1)
    assert np.array_equal(out.detach().cpu().numpy(), z["prop_out"])
    g = torch.randn_like(out)
    out.backward(g)
    ts2, loc2, fl2 = (t.detach().clone().requires_grad_(True) for t in (ts, loc, fl))
    (loc2 + (1 - ts2) * fl2).backward(g)
    for a, b in ((ts, ts2), (loc, loc2), (fl, fl2)):
        assert rel_err(a.grad.cpu().numpy(), b.grad.cpu().numpy())[0] < 1e-6
    res = (int(z["H"]), int(z["W"]))
    pl, pm = purge_unfeasible(cu(z["loc"]), cu(z["mask"]), res)
    assert np.array_equal(pl.cpu().numpy(), z["purge_loc"]) and np.array_equal(pm.cpu().numpy(), z["purge_mask"])
    lg = cu(z["loc"], True)
    pl, _ = purge_unfeasible(lg, cu(z["mask"]), res)
    pl.sum().backward()
    inside = (z["purge_loc"] != 0) | ((z["loc"] == 0) & True)
    assert np.array_equal(lg.grad.cpu().numpy() != 0, np.broadcast_to(((z["loc"][..., 0:1] >= 0) & (z["loc"][..., 0:1] <= res[0] - 1) & (z["loc"][..., 1:2] >= 0) & (z["loc"][..., 1:2] <= res[1] - 1)), z["loc"].shape))


def test_get_interpolation_forward_backward(z):
    from taming_event_flow_b200.utils.iwe import get_interpolation

    res = (int(z["H"]), int(z["W"]))
    loc = cu(z["loc"], True)
    idx, w = get_interpolation(loc, res)
    assert np.array_equal(idx.cpu().numpy(), z["gi_idx"]) and np.array_equal(w.detach().cpu().numpy(), z["gi_w"])
    w.backward(cu(z["gi_gw"]))
    linf, l2 = rel_err(loc.grad.cpu().numpy(), z["gi_gloc"])      # includes the tie / integer-coordinate sub-gradients
    assert linf < TOL and l2 < TOL, (linf, l2)
    idx, w = get_interpolation(cu(z["loc"]), res, round_idx=True)
    assert np.array_equal(idx.cpu().numpy(), z["gi_ridx"]) and np.array_equal(w.cpu().numpy(), z["gi_rw"])


def test_interpolate_forward_backward(z):
    from taming_event_flow_b200.utils.iwe import interpolate

    res = (int(z["H"]), int(z["W"]))
    idx, w = cu(z["gi_idx"]), cu(z["gi_w"], True)
    pol4 = cu(np.concatenate([z["mask"][:, :, 0:1]] * 4, 1))
    for got, ref in ((interpolate(idx, w, res), z["interp_nopol"]), (interpolate(idx, w, res, polarity_mask=pol4), z["interp_pol"]),
                     (interpolate(idx, w, res, polarity_mask=pol4, zeros=cu(z["interp_zeros_in"])), z["interp_zeros"])):
        assert got.shape == ref.shape
        linf, l2 = rel_err(got.detach().cpu().numpy(), ref)
        assert linf < TOL and l2 < TOL
    out = interpolate(idx, w, res, polarity_mask=pol4)
    g = torch.randn_like(out)
    out.backward(g)
    ref = torch.gather(g.view(g.shape[0], -1, 1), 1, idx.long()) * pol4
    assert torch.equal(w.grad, ref)


@pytest.mark.parametrize("ri", [True, False])
@pytest.mark.parametrize("rf", [True, False])
def test_compute_pol_iwe(z, ri, rf):
    from taming_event_flow_b200.utils.iwe import compute_pol_iwe, deblur_events

    res = (int(z["H"]), int(z["W"]))
    ev = cu(z["db_ev_int"] if rf else z["db_ev_frac"])
    out = compute_pol_iwe(cu(z["db_flow"]), ev, res, cu(z["mask"]), round_idx=ri, round_flow=rf)
    ref = z["pol_iwe_ri%d_rf%d" % (ri, rf)]
    assert out.shape == ref.shape
    if ri:
        assert np.array_equal(out.cpu().numpy(), ref)             # integer counts
    else:
        linf, l2 = rel_err(out.cpu().numpy(), ref)
        assert linf < TOL and l2 < TOL
    one = deblur_events(cu(z["db_flow"]), ev, res, round_idx=ri, polarity_mask=cu(z["mask"][:, :, 0:1]), round_flow=rf)
    if ri:
        assert torch.equal(one[:, 0], out[:, 0])
    else:                                                            # two launches: fp32 summation order differs
        assert rel_err(one[:, 0].cpu().numpy(), out[:, 0].cpu().numpy())[0] < TOL


def test_iwe_formatting_and_focus_loss_match_fused_forward():
    """The stand-alone upstream-style methods agree with the oracle's images for one reference time."""
    from taming_event_flow_b200 import synthetic as syn
    from taming_event_flow_b200.loss.flow import Iterative

    B, N, H, W = 2, 3000, 40, 48
    g = torch.Generator().manual_seed(5)
    warped = torch.rand(B, N, 2, generator=g) * torch.tensor([H + 2.0, W + 2.0]) - 1.0
    ts = torch.rand(B, N, 1, generator=g) * 4
    p = (torch.randint(0, 2, (B, N), generator=g) * 2 - 1).float()
    mask = torch.stack([(p > 0).float(), (p < 0).float()], -1)
    m = Iterative(syn.loss_config(H, W, B, 4), "cuda")
    iwe, iwe_ts = m.iwe_formatting(warped.cuda(), torch.cat([mask] * 4, 1).cuda(), torch.cat([ts] * 4, 1).cuda(), 2, 2)
    idx, w = orc.get_interpolation(warped.numpy(), (H, W))
    nts = (1 - np.abs(2 - np.concatenate([ts.numpy()] * 4, 1)) / 2).astype(np.float32)
    for c in range(2):
        pol = np.concatenate([mask.numpy()[:, :, c:c + 1]] * 4, 1)
        assert rel_err(iwe[:, c:c + 1].cpu().numpy(), orc.interpolate(idx, w, (H, W), pol))[0] < TOL
        assert rel_err(iwe_ts[:, c:c + 1].cpu().numpy(), orc.interpolate(idx, w * nts, (H, W), pol))[0] < TOL
    a = iwe_ts / (iwe + 1e-9)
    loss = m.focus_loss(iwe, a)
    ref = 0.0
    for b in range(B):
        nz = (iwe[b].sum(0) != 0).sum().item()
        ref += (a[b] ** 2).sum().item() / (nz + 1e-9)
    assert loss.item() == pytest.approx(ref, rel=1e-5)


def test_encodings_golden():
    from taming_event_flow_b200.dataloader.encodings import events_to_channels, events_to_image, events_to_voxel

    z = np.load(os.path.join(GOLDEN, "encodings.npz"))
    ss = (int(z["H"]), int(z["W"]))
    xs, ys, ts, ps = cu(z["xs"]), cu(z["ys"]), cu(z["ts"]), cu(z["ps"])
    assert np.array_equal(events_to_image(xs, ys, ps, ss).cpu().numpy(), z["image"])            # +-1 sums: exact
    assert np.array_equal(events_to_channels(xs, ys, ps, ss).cpu().numpy(), z["channels"])      # counts: bit-exact
    v = events_to_voxel(xs, ys, ts, ps, int(z["bins"]), ss).cpu().numpy()
    assert v.shape == z["voxel"].shape
    linf, l2 = rel_err(v, z["voxel"])
    assert linf < TOL and l2 < TOL
    assert rel_err(v, z["voxel64"])[1] <= 1.05 * rel_err(z["voxel"], z["voxel64"])[1] + 1e-7


@pytest.mark.parametrize("n", [0, 1, 1_000_000])
def test_encodings_sizes(n):
    from taming_event_flow_b200.dataloader.encodings import events_to_channels, events_to_voxel

    H, W, bins = 480, 640, 5
    g = torch.Generator().manual_seed(n)
    xs = torch.randint(0, W, (n,), generator=g).float()
    ys = torch.randint(0, H, (n,), generator=g).float()
    ts = torch.rand(n, generator=g)
    ps = (torch.randint(0, 2, (n,), generator=g) * 2 - 1).float()
    ch = events_to_channels(xs.cuda(), ys.cuda(), ps.cuda(), (H, W)).cpu().numpy()
    assert np.array_equal(ch, orc.events_to_channels(xs.numpy(), ys.numpy(), ps.numpy(), (H, W)))
    assert ch.sum() == n                                                                           # a checksum of checksums
    v = events_to_voxel(xs.cuda(), ys.cuda(), ts.cuda(), ps.cuda(), bins, (H, W)).cpu().numpy()
    ref = orc.events_to_voxel(xs.numpy(), ys.numpy(), ts.numpy(), ps.numpy(), bins, (H, W))
    assert rel_err(v, ref)[0] < TOL if n else not v.any()


def test_get_hot_event_mask_against_its_specification():
    """get_hot_event_mask is not in the reference (parity unpinned); the kernel is checked bit-exactly against
    the CPU restatement of its published specification, including ties, the min_obvs gate and the max_px cap."""
    from taming_event_flow_b200.dataloader.encodings import get_hot_event_mask

    g = torch.Generator().manual_seed(4)
    H, W = 60, 80
    rate = torch.rand(H, W, generator=g) * 0.7
    hot = torch.randperm(H * W, generator=g)[:150]
    rate.view(-1)[hot] = 0.8 + torch.rand(150, generator=g)
    rate.view(-1)[hot[:10]] = 1.25                                   # ties: the lowest flat index goes first
    for idx, max_px in ((3, 100), (6, 100), (6, 1000), (6, 0)):
        r = rate.clone().cuda()
        m = get_hot_event_mask(r, idx, max_px=max_px)
        em, er = orc.get_hot_event_mask(rate.numpy(), idx, max_px=max_px)
        assert np.array_equal(m.cpu().numpy(), em) and np.array_equal(r.cpu().numpy(), er)
        assert int((m == 0).sum()) == (0 if idx <= 5 else min(max_px, 150))


def test_encodings_out_of_range_raises_like_index_put():
    """An event outside the sensor: the reference's index_put_ raises IndexError (dataloader/encodings.py:23-27; checked live in
    tests/test_primitives_vs_reference_live.py).  Negative coordinates down to -size wrap like Python indexing."""
    from taming_event_flow_b200.dataloader import encodings as enc

    H, W = 12, 20
    xs = torch.tensor([3.0, 19.0, -1.0, 5.7]).cuda()
    ys = torch.tensor([2.0, 11.0, -12.0, 0.2]).cuda()
    ts = torch.tensor([0.0, 0.3, 0.6, 1.0]).cuda()
    ps = torch.tensor([1.0, -1.0, 1.0, -1.0]).cuda()
    img = enc.events_to_image(xs, ys, ps, (H, W)).cpu().numpy()
    ref = orc.events_to_image(xs.cpu().numpy(), ys.cpu().numpy(), ps.cpu().numpy(), (H, W))
    assert np.array_equal(img, ref) and img[0, 19] == 1.0 and img[0, 5] == -1.0       # (-12, -1) wraps to (0, 19); (0.2, 5.7) truncates
    for bad_x, bad_y in ((20.0, 0.0), (0.0, 12.0), (-21.0, 0.0), (0.0, -13.0), (1e9, 0.0)):
        bx, by = xs.clone(), ys.clone()
        bx[1], by[1] = bad_x, bad_y
        with pytest.raises(IndexError):
            enc.events_to_image(bx, by, ps, (H, W))
        with pytest.raises(IndexError):
            enc.events_to_image(bx, by, ps, (H, W), accumulate=False)
        with pytest.raises(IndexError):
            enc.events_to_channels(bx, by, ps, (H, W))
        with pytest.raises(IndexError):
            enc.events_to_voxel(bx, by, ts, ps, 3, (H, W))
    enc.CHECK_BOUNDS = False                     # no read-back: the event is dropped
    try:
        bx = xs.clone()
        bx[1] = 20.0
        ch = enc.events_to_channels(bx, ys, ps, (H, W))
        assert ch.sum().item() == 3.0
    finally:
        enc.CHECK_BOUNDS = True


@pytest.mark.parametrize("n,offset", [(1, 0), (3, 1), (4, 0), (1001, 1), (1002, 2), (4096, 0), (50001, 3)])
def test_encodings_vector_loads_tails_and_unaligned_views(n, offset):
    """Four events per thread with 16-byte loads: counts that are not multiples of four and views that start at an
    unaligned element take the element-wise path; same results bit for bit."""
    from taming_event_flow_b200.dataloader.encodings import events_to_channels, events_to_image, events_to_voxel

    H, W, bins = 40, 56, 5
    g = torch.Generator().manual_seed(n * 7 + offset)
    xs = torch.randint(0, W, (n + offset,), generator=g).float()
    ys = torch.randint(0, H, (n + offset,), generator=g).float()
    ts = torch.rand(n + offset, generator=g)
    ps = (torch.randint(0, 2, (n + offset,), generator=g) * 2 - 1).float()
    cx, cy, ct, cp = (a.cuda()[offset:] for a in (xs, ys, ts, ps))
    hx, hy, ht, hp = (a[offset:].numpy() for a in (xs, ys, ts, ps))
    assert np.array_equal(events_to_channels(cx, cy, cp, (H, W)).cpu().numpy(), orc.events_to_channels(hx, hy, hp, (H, W)))
    assert np.array_equal(events_to_image(cx, cy, cp, (H, W)).cpu().numpy(), orc.events_to_image(hx, hy, hp, (H, W)))
    v = events_to_voxel(cx, cy, ct, cp, bins, (H, W)).cpu().numpy()
    assert rel_err(v, orc.events_to_voxel(hx, hy, ht, hp, bins, (H, W)))[0] < TOL


@pytest.mark.parametrize("n", [0, 5, 20000])
def test_events_to_image_without_accumulation_keeps_the_last_event(n):
    """accumulate=False: index_put_ without accumulation; the reference's CPU kernel applies the events in order, so the last
    event of a pixel wins.  The GPU result must be that, deterministically (run twice, heavy pixel reuse)."""
    from taming_event_flow_b200.dataloader.encodings import events_to_image

    H, W = 16, 24
    g = torch.Generator().manual_seed(n)
    xs = torch.randint(0, W, (n,), generator=g).float()
    ys = torch.randint(0, H, (n,), generator=g).float()
    ps = torch.randn(n, generator=g)
    ref = np.zeros((H, W), np.float32)
    for i in range(n):
        ref[int(ys[i]), int(xs[i])] = ps[i]
    for _ in range(2):
        out = events_to_image(xs.cuda(), ys.cuda(), ps.cuda(), (H, W), accumulate=False).cpu().numpy()
        assert np.array_equal(out, ref)


def test_mask_shapes_are_validated():
    from taming_event_flow_b200._lib import TefShapeError
    from taming_event_flow_b200.utils.iwe import compute_pol_iwe, deblur_events, purge_unfeasible

    B, N, H, W = 2, 10, 8, 12
    loc = torch.rand(B, N, 2).cuda() * 6
    out_l, out_m = purge_unfeasible(loc, torch.ones(B, N, 1).cuda(), (H, W))        # [B,N,1] broadcasts like upstream
    assert out_m.shape == (B, N, 1)
    for bad in (torch.ones(B, N - 1, 2), torch.ones(B, N, 3), torch.ones(N, 2)):
        with pytest.raises(TefShapeError):
            purge_unfeasible(loc, bad.cuda(), (H, W))
    flow, ev = torch.zeros(B, 2, H, W).cuda(), torch.zeros(B, N, 4).cuda()
    deblur_events(flow, ev, (H, W), polarity_mask=torch.ones(B, N, 1).cuda())
    with pytest.raises(TefShapeError):          # a two-column mask would be read with the wrong stride
        deblur_events(flow, ev, (H, W), polarity_mask=torch.ones(B, N, 2).cuda())
    with pytest.raises(TefShapeError):
        compute_pol_iwe(flow, ev, (H, W), torch.ones(B, N, 1).cuda())
    with pytest.raises(TefShapeError):
        deblur_events(torch.zeros(B, 2, H, W + 1).cuda(), ev, (H, W))
