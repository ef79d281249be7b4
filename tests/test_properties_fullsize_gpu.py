"""Parity at BASELINE.json's full size (480x640, 1 M events/window, 10 passes) through size-independent properties --
the CPU oracle would need minutes and ~30 GB there:

* zero flow: every event stays on its integer pixel, so each count image must equal, bit for bit, the sum of the
  per-window event-count encodings (ties the loss kernels to the events_to_channels kernel), the time-weighted image is
  bounded by it, and no flow gradient may be NaN;
* duplicating every event doubles both images, which leaves the normalised timestamps and the loss unchanged;
* swapping the polarity channels of the masks leaves loss and gradients unchanged;
* permuting the events of each window changes nothing beyond fp32 summation order."""
import numpy as np
import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu
H, W, P, N = 480, 640, 10, 1_000_000


def _windows(seed, n=N):
    g = torch.Generator().manual_seed(seed)
    evs, mks = [], []
    for t in range(P):
        ts, _ = torch.sort(torch.rand(1, n, generator=g), dim=1)
        ts = (ts - ts[:, :1]) / (ts[:, -1:] - ts[:, :1])
        ev = torch.stack([ts, torch.randint(0, H, (1, n), generator=g).float(), torch.randint(0, W, (1, n), generator=g).float(),
                          (torch.randint(0, 2, (1, n), generator=g) * 2 - 1).float()], -1)
        evs.append(ev)
        mks.append(torch.stack([(ev[..., 3] > 0).float(), (ev[..., 3] < 0).float()], -1))
    return evs, mks


def _flows(seed, sigma):
    from taming_event_flow_b200 import synthetic as syn

    g = torch.Generator().manual_seed(seed)
    return [syn.make_flow(g, 1, H, W, sigma) for _ in range(P)]


def _run(evs, mks, flows, want_images=False):
    from taming_event_flow_b200 import synthetic as syn
    from taming_event_flow_b200.loss.flow import Iterative

    m = Iterative(syn.loss_config(H, W, 1, P), "cuda")
    fl = [f.cuda().requires_grad_(True) for f in flows]
    empty_e, empty_m = torch.zeros(1, 0, 4, device="cuda"), torch.zeros(1, 0, 2, device="cuda")
    for t in range(P):
        m.update([fl[t]], evs[t].cuda().clone(), mks[t].cuda(), empty_e.clone(), empty_m)
    loss = m()
    img = m.images() if want_images else None
    loss.backward()
    grads = torch.stack([f.grad for f in fl]).cpu().numpy()
    return loss.item(), grads, img


def test_zero_flow_images_equal_event_counts_exactly():
    from taming_event_flow_b200.dataloader.encodings import events_to_channels

    evs, mks = _windows(1)
    loss, grads, img = _run(evs, mks, [torch.zeros(1, 2, H, W) for _ in range(P)], want_images=True)
    assert np.isfinite(loss) and np.isfinite(grads).all()
    counts = [events_to_channels(e[0, :, 2].cuda(), e[0, :, 1].cuda(), e[0, :, 3].cuda(), (H, W)) for e in evs]   # [2,H,W] each
    delta = P // 2
    for tref in range(P + 1):
        lo, hi = max(0, tref - delta), min(P, tref + delta)          # windows feeding this reference time (loss/flow.py:685-686)
        expect = torch.stack(counts[lo:hi]).sum(0)
        got = img[0, 0, tref, 0:2]
        assert torch.equal(got, expect), "count image of tref %d differs from the event-count encoding" % tref
        tw = img[0, 0, tref, 2:4]
        assert bool((tw <= got + 1e-3).all()) and bool((tw >= 0).all())
    assert float(torch.stack(counts).sum()) == P * N                 # a checksum of checksums


def test_duplication_polarity_swap_and_permutation_invariance():
    evs, mks = _windows(2)
    flows = _flows(3, 3.0)
    loss, grads, _ = _run(evs, mks, flows)
    # every event twice
    l2, g2, _ = _run([torch.cat([e, e], 1) for e in evs], [torch.cat([m, m], 1) for m in mks], flows)
    assert abs(l2 - loss) <= 2e-6 * abs(loss)
    # (the gradients are not compared here: the 1e-9 in iwe_ts / (iwe + 1e-9), loss/flow.py:727, does not scale with the
    #  event count, and pixels touched by a ~1e-6 corner weight dominate the gradient norm)
    assert np.isfinite(g2).all()
    # polarity channels swapped
    l3, g3, _ = _run(evs, [m.flip(-1) for m in mks], flows)
    assert abs(l3 - loss) <= 2e-6 * abs(loss)
    assert rel_err(g3, grads)[1] < 1e-5
    # events permuted inside each window
    g = torch.Generator().manual_seed(5)
    perms = [torch.randperm(N, generator=g) for _ in range(P)]
    l4, g4, _ = _run([e[:, p] for e, p in zip(evs, perms)], [m[:, p] for m, p in zip(mks, perms)], flows)
    assert abs(l4 - loss) <= 2e-6 * abs(loss)
    assert rel_err(g4, grads)[1] < 1e-5
