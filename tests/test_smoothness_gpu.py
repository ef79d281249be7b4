"""GPU parity of the fused smoothness priors (csrc/tef_cm_smooth.cu): values and gradients against the
unmodified reference's autograd (tests/golden/smoothness.npz, fp32 and its fp64 run) and the numpy oracle at a larger size.
Tolerance 1e-5 norm-relative (north_star's bound for losses and gradients)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import smooth_oracle as so  # noqa: E402
from taming_event_flow_b200 import synthetic as syn  # noqa: E402
from taming_event_flow_b200.loss import flow as tef_flow  # noqa: E402
from util import rel_err  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-5
G = np.load(os.path.join(ROOT, "tests", "golden", "smoothness.npz"))


def _module(kind, B, P, H, W, flows_np, events=None, masks=None, n_updates=None):
    cfg = syn.loss_config(H, W, B, P, 1, "two", warping=kind)
    cfg["loss"]["flow_spat_smooth_weight"], cfg["loss"]["flow_temp_smooth_weight"] = 1.0, 1.0
    m = getattr(tef_flow, kind)(cfg, torch.device("cuda"))
    flows = [[torch.from_numpy(f).cuda().requires_grad_(True) for f in per] for per in flows_np]
    for t in range(P if n_updates is None else n_updates):
        ev = torch.from_numpy(events[t]).cuda() if events is not None else torch.zeros(B, 0, 4, device="cuda")
        mk = torch.from_numpy(masks[t]).cuda() if masks is not None else torch.zeros(B, 0, 2, device="cuda")
        m.update(flows[t], ev, mk, torch.zeros(B, 0, 4, device="cuda"), torch.zeros(B, 0, 2, device="cuda"))
    return m, flows


@pytest.mark.parametrize("kind", ["Iterative", "Linear"])
def test_priors_match_reference(kind):
    B, P, H, W, F = (int(G[k]) for k in ("B", "P", "H", "W", "F"))
    flows_np = [[G["flow%d_%d" % (t, f)] for f in range(F)] for t in range(P)]
    ev, mk = [G["ev%d" % t] for t in range(P)], [G["mk%d" % t] for t in range(P)]
    m, flows = _module(kind, B, P, H, W, flows_np, ev, mk)
    for name, fn in (("spat", m.flow_spatial_smoothing), ("temp", m.flow_temporal_smoothing)):
        for per in flows:
            for f in per:
                f.grad = None
        val = fn()
        val.backward()
        for tag in ("32", "64"):
            ref = float(G["%s_%s" % (name, tag)])
            assert abs(val.item() - ref) <= TOL * abs(ref), (name, tag, val.item(), ref)
            g = np.stack([np.stack([flows[t][f].grad.cpu().numpy() for t in range(P)]) for f in range(F)])
            linf, l2 = rel_err(g, G["%s_grad_%s" % (name, tag)])
            assert linf < TOL and l2 < TOL, (name, tag, linf, l2)
    # after two of the four updates (upstream's tensors only hold the passes seen so far)
    m2, _ = _module(kind, B, P, H, W, flows_np, ev, mk, n_updates=2)
    assert abs(m2.flow_spatial_smoothing().item() - float(G["spat_2passes_32"])) <= TOL * float(G["spat_2passes_32"])
    assert abs(m2.flow_temporal_smoothing().item() - float(G["temp_2passes_32"])) <= TOL * float(G["temp_2passes_32"])


def test_priors_seeded_vs_oracle_and_scaling():
    """Larger maps (odd sizes, big flows so that many targets leave the image): values against the numpy oracle, and the
    gradient against a directional finite difference of the oracle in fp64."""
    B, P, H, W, F = 2, 3, 61, 83, 1
    rng = np.random.default_rng(3)
    flows_np = [[(rng.normal(0, 6, (B, 2, H, W))).astype(np.float32)] for _ in range(P)]
    m, flows = _module("Iterative", B, P, H, W, flows_np)
    for fn, orc in ((m.flow_spatial_smoothing, so.flow_spatial_smoothing), (m.flow_temporal_smoothing, so.flow_temporal_smoothing)):
        for per in flows:
            per[0].grad = None
        val = fn()
        ref = orc(flows_np)
        assert abs(val.item() - ref) <= TOL * abs(ref), (val.item(), ref)
        (3.0 * val).backward()                                                   # upstream gradient 3
        d = [[rng.normal(0, 1, (B, 2, H, W))] for _ in range(P)]
        h = 1e-6
        plus = [[flows_np[t][0].astype(np.float64) + h * d[t][0]] for t in range(P)]
        minus = [[flows_np[t][0].astype(np.float64) - h * d[t][0]] for t in range(P)]
        fd = 3.0 * (orc(plus) - orc(minus)) / (2 * h)
        an = sum(float((flows[t][0].grad.double().cpu().numpy() * d[t][0]).sum()) for t in range(P))
        assert abs(an - fd) <= 2e-4 * max(abs(fd), 1e-3), (an, fd)


def test_priors_without_grad_and_in_the_total_loss():
    B, P, H, W, F = (int(G[k]) for k in ("B", "P", "H", "W", "F"))
    flows_np = [[G["flow%d_%d" % (t, f)] for f in range(F)] for t in range(P)]
    ev, mk = [G["ev%d" % t] for t in range(P)], [G["mk%d" % t] for t in range(P)]
    m, flows = _module("Iterative", B, P, H, W, flows_np, ev, mk)
    with torch.no_grad():
        s, t = m.flow_spatial_smoothing(), m.flow_temporal_smoothing()
    assert not s.requires_grad and abs(s.item() - float(G["spat_32"])) <= TOL * float(G["spat_32"])
    m.flow_spat_smooth_weight, m.flow_temp_smooth_weight = 0.5, 2.0
    total = m()
    m.flow_spat_smooth_weight = m.flow_temp_smooth_weight = None
    m2, _ = _module("Iterative", B, P, H, W, flows_np, ev, mk)
    m2.flow_spat_smooth_weight = m2.flow_temp_smooth_weight = None
    cm = m2()
    want = cm.item() + 0.5 * s.item() + 2.0 * t.item()
    assert abs(total.item() - want) <= TOL * abs(want)


@pytest.mark.parametrize("seed", range(8))
def test_priors_random_shapes(seed):
    """Odd resolutions, 1-3 flow scales, 2-6 passes, small and large flows: values against the numpy oracle, gradients
    against a central finite difference of the oracle in fp64 along a random direction."""
    r = np.random.default_rng(70 + seed)
    B, P, F = int(r.integers(1, 4)), int(r.integers(2, 7)), int(r.integers(1, 4))
    H, W = int(r.integers(3, 45)), int(r.integers(3, 60))
    sigma = float(r.choice([0.5, 3.0, 8.0]))
    flows_np = [[r.normal(0, sigma, (B, 2, H, W)).astype(np.float32) for _ in range(F)] for _ in range(P)]
    m, flows = _module("Iterative" if seed % 2 == 0 else "Linear", B, P, H, W, flows_np)
    for fn, orc_fn in ((m.flow_spatial_smoothing, so.flow_spatial_smoothing), (m.flow_temporal_smoothing, so.flow_temporal_smoothing)):
        for per in flows:
            for f in per:
                f.grad = None
        val = fn()
        ref = orc_fn(flows_np)
        assert abs(val.item() - ref) <= TOL * abs(ref), (val.item(), ref)
        val.backward()
        d = [[r.normal(0, 1, (B, 2, H, W)) for _ in range(F)] for _ in range(P)]
        h = 1e-6
        plus = [[flows_np[t][f].astype(np.float64) + h * d[t][f] for f in range(F)] for t in range(P)]
        minus = [[flows_np[t][f].astype(np.float64) - h * d[t][f] for f in range(F)] for t in range(P)]
        fd = (orc_fn(plus) - orc_fn(minus)) / (2 * h)
        an = sum(float((flows[t][f].grad.double().cpu().numpy() * d[t][f]).sum()) for t in range(P) for f in range(F))
        assert abs(an - fd) <= 5e-4 * max(abs(fd), 1e-3), (an, fd)
