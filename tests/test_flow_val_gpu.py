"""Validation criteria (taming_event_flow_b200.loss.flow_val, first "next" row of SURVEY.md §8f) against golden vectors made
from the unmodified reference's loss/flow_val.py: FWL, RSAT, AEE and every window image after 1, 3 and 5 windows, for the
Linear and the Iterative flavour.  Rounded-index images are integer counts (bit-exact); the rest within 1e-5 norm-relative."""
import copy
import os

import numpy as np
import pytest
import torch

from util import GOLDEN, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("name", ["linear", "iterative"])
def test_flow_val_golden(name):
    from taming_event_flow_b200.loss import flow_val as fv

    z = np.load(os.path.join(GOLDEN, "flow_val.npz"))
    H, W, P = int(z["H"]), int(z["W"]), int(z["P"])
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"round_ts": False}, "vis": {"mask_output": True}, "metrics": {"name": ["AEE"]}}
    m = (fv.Linear if name == "linear" else fv.Iterative)(copy.deepcopy(cfg), "cuda")
    gt = torch.tensor(z["gt"]).cuda()
    checked = 0
    for t in range(P):
        ev, mk, flow, em = (torch.tensor(z["%s%d" % (k, t)]).cuda() for k in ("ev", "mk", "flow", "emask"))
        ev_in = ev.clone()
        m.update([flow.clone()], ev_in, mk.clone(), em.clone())
        assert torch.equal(ev_in[:, :, 0], ev[:, :, 0] + t)                  # in-place timestamp update, as upstream
        if t not in (0, 2, P - 1):
            continue
        key = "%s_t%d_" % (name, t)
        got = {"fwl": m.fwl(), "rsat": m.rsat(), "events": m.window_events(), "events_round": m.window_events(round_idx=True),
               "aee": m.compute_aee(flow, gt, mask=m._event_mask)}
        modes = (None,) if name == "linear" else (None, "forward", "backward")
        for mode in modes:
            got["flow_%s" % mode] = m.window_flow(mode=mode)
            got["flow_nomask_%s" % mode] = m.window_flow(mode=mode, mask=False)
        for mode in ((None,) if name == "linear" else ("forward", "backward")):
            for ri in (False, True):
                got["iwe_%s_%d" % (mode, ri)] = m.window_iwe(mode=mode, round_idx=ri)
        for k, v in got.items():
            ref, v = z[key + k], v.cpu().numpy()
            assert v.shape == ref.shape, (key + k, v.shape, ref.shape)
            if k in ("events", "events_round") or k.endswith("_1"):
                assert np.array_equal(v, ref), key + k                        # integer counts
            else:
                linf, l2 = rel_err(v, ref)
                assert linf < TOL and l2 < TOL, (key + k, linf, l2)
            checked += 1
    assert checked >= 27
    m.reset()
    assert m.num_passes == 0
