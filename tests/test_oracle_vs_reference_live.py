"""CPU, build container only: the C oracle against the UNMODIFIED reference imported live from /root/reference on seeded
random small configurations (every mode, 1-3 temporal scales, border compensation on/off, ragged batches, empty detached
lists, events beyond the sensor).  Complements the committed golden vectors; skipped where the reference is not mounted
(the GPU box)."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

REF = os.environ.get("TEF_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "loss")), reason="reference not mounted")

from oracle import cm_oracle as orc  # noqa: E402
from taming_event_flow_b200 import synthetic as syn  # noqa: E402
from util import rel_err  # noqa: E402


def _reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import warnings

    warnings.filterwarnings("ignore")
    from loss.flow import Iterative, Linear  # reference

    return Iterative, Linear


def _draw(seed):
    r = np.random.default_rng(seed)
    kind = "iterative" if r.random() < 0.7 else "linear"
    mode = ["one", "two", "four"][int(r.integers(0, 3))] if kind == "iterative" else "two"
    border = bool(r.random() < 0.6) and mode != "four"
    P = int(r.integers(1, 9))
    if mode == "four":
        P = 2 * int(r.integers(1, 4))
    return dict(kind=kind, mode=mode, border=border, P=P, S=int(r.integers(1, 4)), B=int(r.integers(1, 3)), F=int(r.integers(1, 3)),
                H=int(r.integers(5, 36)), W=int(r.integers(5, 44)), N=int(r.integers(1, 200)), Nd=int(r.integers(0, 2)) * int(r.integers(1, 100)),
                sigma=float(r.choice([0.3, 2.0, 5.0])), ragged=bool(r.random() < 0.5), outside=bool(r.random() < 0.3))


@pytest.mark.parametrize("seed", range(48))
def test_oracle_matches_live_reference(seed):
    Iterative, Linear = _reference()
    c = _draw(7000 + seed)
    seq = syn.make_sequence(2000 + seed, c["B"], c["P"], c["N"], c["Nd"], c["H"], c["W"], c["F"], c["sigma"], c["ragged"], "uniform")
    if c["outside"]:
        for ev in seq["events"]:
            ev[:, ::7, 1] += c["H"] * 0.75
            ev[:, ::5, 2] -= 3.25
    P_cfg = c["P"] // 2 if c["mode"] == "four" else c["P"]
    cfg = syn.loss_config(c["H"], c["W"], c["B"], P_cfg, c["S"], c["mode"])
    m = (Iterative if c["kind"] == "iterative" else Linear)(copy.deepcopy(cfg), "cpu")
    m.border_compensation = c["border"]
    flows = [[f.clone().requires_grad_(True) for f in per] for per in seq["flows"]]
    oc = orc.make_cfg(c["B"], c["H"], c["W"], c["P"], c["F"], c["S"], c["mode"], c["border"])
    fn = orc.iterative if c["kind"] == "iterative" else orc.linear
    try:
        for t in range(c["P"]):
            m.update(flows[t], seq["events"][t].clone(), seq["masks"][t].clone(), seq["d_events"][t].clone(), seq["d_masks"][t].clone())
        loss = m()
    except (RuntimeError, ValueError, TypeError, IndexError):                    # the reference rejects the configuration: so must the oracle
        with pytest.raises(orc.OracleError):
            fn(oc, seq["flows"], seq["events"], seq["masks"], seq["d_events"], seq["d_masks"], np.float32, want_grad=True)
        return
    loss.backward()
    o = fn(oc, seq["flows"], seq["events"], seq["masks"], seq["d_events"], seq["d_masks"], np.float32, want_grad=True)
    assert abs(o["loss"] - loss.item()) <= 1e-5 * max(abs(loss.item()), 1e-6), (c, o["loss"], loss.item())
    g = np.stack([np.stack([(flows[t][f].grad if flows[t][f].grad is not None else torch.zeros_like(flows[t][f])).numpy()
                            for t in range(c["P"])]) for f in range(c["F"])])
    if np.abs(g).max() > 0:
        linf, l2 = rel_err(o["gflow"], g)
        assert linf < 1e-5 and l2 < 1e-5, (c, linf, l2)
    else:
        assert not o["gflow"].any()
