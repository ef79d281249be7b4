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


@pytest.mark.parametrize("fixture", ["flow_val.npz", "flow_val_b.npz"])
@pytest.mark.parametrize("name", ["linear", "iterative"])
def test_flow_val_golden(name, fixture):
    from taming_event_flow_b200.loss import flow_val as fv

    z = np.load(os.path.join(GOLDEN, fixture))
    H, W, P = int(z["H"]), int(z["W"]), int(z["P"])
    round_ts = bool(int(z["round_ts"])) if "round_ts" in z.files else False
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"round_ts": round_ts}, "vis": {"mask_output": True}, "metrics": {"name": ["AEE"]}}
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


def test_forward_prop_flow_matches_operator_route():
    """The fused splat/normalise kernels against the same computation spelled out with the stand-alone operators
    (upstream loss/flow_val.py:43-74 line by line)."""
    from taming_event_flow_b200.loss import flow_val as fv
    from taming_event_flow_b200.utils.iwe import event_propagation, get_event_flow, get_interpolation, interpolate, purge_unfeasible

    H, W, P = 37, 53, 4
    g = torch.Generator().manual_seed(2)
    mx = (torch.randn(1, P, H, W, generator=g) * 4).cuda()
    my = (torch.randn(1, P, H, W, generator=g) * 4).cuda()
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"round_ts": False}, "vis": {"mask_output": True}, "metrics": {"name": ["AEE"]}}
    m = fv.Iterative(cfg, "cuda")
    for i, tref in ((0, 1), (2, 3), (1, 3)):
        px_flow = get_event_flow(mx[:, i], my[:, i], m.indices)
        warped, mask = purge_unfeasible(event_propagation(i, m.indices, px_flow, tref), m.indices_mask.clone(), m.res)
        idx, w = get_interpolation(warped, m.res)
        mask4, flow4 = torch.cat([mask] * 4, 1), torch.cat([px_flow] * 4, 1)
        norm = interpolate(idx, w, m.res, polarity_mask=mask4)
        want_y = interpolate(idx, w * flow4[..., 0:1], m.res, polarity_mask=mask4) / (norm + 1e-9)
        want_x = interpolate(idx, w * flow4[..., 1:2], m.res, polarity_mask=mask4) / (norm + 1e-9)
        fx, fy = m.forward_prop_flow(i, tref, mx, my)
        assert fx.shape == (1, 1, H, W) and fy.shape == (1, 1, H, W)
        for got, want in ((fx, want_x), (fy, want_y)):
            linf, l2 = rel_err(got.cpu().numpy(), want.cpu().numpy())
            assert linf < 1e-6 and l2 < 1e-6, (i, tref, linf, l2)
        assert torch.equal(norm == 0, fx == 0)                                # untouched pixels stay exactly 0


def test_validation_rejects_batches():
    from taming_event_flow_b200.loss import flow_val as fv

    cfg = {"loader": {"resolution": [8, 8]}, "loss": {"round_ts": False}, "vis": {"mask_output": True}, "metrics": {"name": ["AEE"]}}
    m = fv.Iterative(cfg, "cuda")
    with pytest.raises(RuntimeError):
        m.update([torch.zeros(2, 2, 8, 8, device="cuda")], torch.zeros(2, 5, 4, device="cuda"), torch.zeros(2, 5, 2, device="cuda"),
                 torch.zeros(2, 1, 8, 8, device="cuda"))
