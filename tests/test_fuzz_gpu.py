"""Seeded random small configurations (odd resolutions, tiny and ragged event counts, every mode, 1-3 temporal scales,
border compensation on/off, detached lists present or empty, events outside the sensor): CUDA path against the CPU
oracle, same tolerances as the fixed cases (1e-5 norm-relative; non-zero pixel sets exact).  Configurations the reference
rejects (an empty window list at some scale) must be rejected by both."""
import numpy as np
import pytest
import torch

from oracle import cm_oracle as orc
from taming_event_flow_b200 import synthetic as syn
from taming_event_flow_b200.loss import flow as tef_flow
from util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _draw(seed):
    r = np.random.default_rng(seed)
    kind = "iterative" if r.random() < 0.7 else "linear"
    mode = ["one", "two", "four"][int(r.integers(0, 3))] if kind == "iterative" else "two"
    border = bool(r.random() < 0.6) and mode != "four"
    P = int(r.integers(1, 13))
    if mode == "four":
        P = 2 * int(r.integers(1, 6))                        # the module doubles config passes_loss (loss/flow.py:422-423)
    return dict(kind=kind, mode=mode, border=border, P=P, S=int(r.integers(1, 4)), B=int(r.integers(1, 4)), F=int(r.integers(1, 4)),
                H=int(r.integers(5, 70)), W=int(r.integers(5, 90)), N=int(r.integers(1, 400)), Nd=int(r.integers(0, 3)) * int(r.integers(1, 200)),
                sigma=float(r.choice([0.3, 2.0, 6.0])), ragged=bool(r.random() < 0.5), outside=bool(r.random() < 0.3))


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("TEF_FUZZ_SEEDS", "60"))))
def test_random_configuration(seed):
    c = _draw(seed)
    seq = syn.make_sequence(1000 + seed, c["B"], c["P"], c["N"], c["Nd"], c["H"], c["W"], c["F"], c["sigma"], c["ragged"], "uniform")
    if c["outside"]:                                         # a few events beyond the sensor (and at fractional positions)
        for ev in seq["events"]:
            ev[:, ::7, 1] += c["H"] * 0.75
            ev[:, ::5, 2] -= 3.25
    P_cfg = c["P"] // 2 if c["mode"] == "four" else c["P"]
    cfg = syn.loss_config(c["H"], c["W"], c["B"], P_cfg, c["S"], c["mode"], warping="Iterative" if c["kind"] == "iterative" else "Linear")
    oc = orc.make_cfg(c["B"], c["H"], c["W"], c["P"], c["F"], c["S"], c["mode"], c["border"])
    fn = orc.iterative if c["kind"] == "iterative" else orc.linear
    m = (tef_flow.Iterative if c["kind"] == "iterative" else tef_flow.Linear)(cfg, torch.device("cuda"))
    m.border_compensation = c["border"]
    flows = [[f.cuda().requires_grad_(True) for f in per] for per in seq["flows"]]
    for t in range(c["P"]):
        m.update(flows[t], seq["events"][t].cuda().clone(), seq["masks"][t].cuda(), seq["d_events"][t].cuda().clone(), seq["d_masks"][t].cuda())
    try:
        o = fn(oc, seq["flows"], seq["events"], seq["masks"], seq["d_events"], seq["d_masks"], np.float32, want_grad=True, want_iwe=True)
    except orc.OracleError:
        with pytest.raises((RuntimeError, TypeError)):
            m()
        return
    loss = m()
    iwe = m.images().cpu().numpy()
    loss.backward()
    g = np.stack([np.stack([flows[t][f].grad.cpu().numpy() for t in range(c["P"])]) for f in range(c["F"])])
    assert abs(loss.item() - o["loss"]) <= TOL * max(abs(o["loss"]), 1e-6), (c, loss.item(), o["loss"])
    assert np.array_equal(iwe != 0, o["iwe"] != 0), c
    linf, l2 = rel_err(iwe, o["iwe"])
    assert linf < TOL and l2 < TOL, (c, "iwe", linf, l2)
    if np.abs(o["gflow"]).max() > 0:
        linf, l2 = rel_err(g, o["gflow"])
        assert linf < TOL and l2 < TOL, (c, "grad", linf, l2)
    else:
        assert not g.any()


@pytest.mark.parametrize("N,Nd", [(0, 50), (0, 0), (1, 0), (0, 1)])
def test_empty_event_lists(N, Nd):
    """No gradient-carrying events (or no events at all): the reference returns the detached events' loss (or 0) and zero
    gradients; so must the kernels, without touching empty buffers."""
    B, P, H, W, F = 2, 4, 16, 20, 1
    seq = syn.make_sequence(3, B, P, max(N, 1), max(Nd, 1), H, W, F, 2.0, False, "uniform")
    if N == 0:
        seq["events"], seq["masks"] = [e[:, :0] for e in seq["events"]], [m[:, :0] for m in seq["masks"]]
    if Nd == 0:
        seq["d_events"], seq["d_masks"] = [e[:, :0] for e in seq["d_events"]], [m[:, :0] for m in seq["d_masks"]]
    o = orc.iterative(orc.make_cfg(B, H, W, P, F, 1, "two", True), seq["flows"], seq["events"], seq["masks"], seq["d_events"], seq["d_masks"],
                      np.float32, want_grad=True, want_iwe=True)
    for kind in (tef_flow.Iterative, tef_flow.Linear):
        m = kind(syn.loss_config(H, W, B, P, 1, "two", warping=kind.__name__), torch.device("cuda"))
        flows = [[f.cuda().requires_grad_(True) for f in per] for per in seq["flows"]]
        for t in range(P):
            m.update(flows[t], seq["events"][t].cuda().clone(), seq["masks"][t].cuda(), seq["d_events"][t].cuda().clone(), seq["d_masks"][t].cuda())
        loss = m()
        loss.backward()
        g = torch.stack([f.grad for per in flows for f in per])
        assert torch.isfinite(loss) and torch.isfinite(g).all()
        if kind is tef_flow.Iterative:
            assert abs(loss.item() - o["loss"]) <= TOL * max(abs(o["loss"]), 1e-6)
            linf = np.abs(g.cpu().numpy().reshape(o["gflow"].shape) - o["gflow"]).max()
            assert linf <= TOL * max(np.abs(o["gflow"]).max(), 1e-12)
        if N == 0:
            assert not g.any()
