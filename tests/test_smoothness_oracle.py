"""CPU: numpy restatement of the smoothness priors (oracle/smooth_oracle.py) against the reference (tests/golden/smoothness.npz)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import smooth_oracle as so  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "smoothness.npz"))
P, F = int(G["P"]), int(G["F"])
FLOWS = [[G["flow%d_%d" % (t, f)] for f in range(F)] for t in range(P)]


def test_priors_match_reference_fp64():
    assert abs(so.flow_spatial_smoothing(FLOWS) - float(G["spat_64"])) < 1e-12
    assert abs(so.flow_temporal_smoothing(FLOWS) - float(G["temp_64"])) < 1e-11
    assert abs(so.flow_spatial_smoothing(FLOWS[:2]) - float(G["spat_2passes_64"])) < 1e-12
    assert abs(so.flow_temporal_smoothing(FLOWS[:2]) - float(G["temp_2passes_64"])) < 1e-11


def test_priors_fp32_within_tolerance():
    assert abs(so.flow_spatial_smoothing(FLOWS, np.float32) - float(G["spat_32"])) < 1e-5 * float(G["spat_32"])
    assert abs(so.flow_temporal_smoothing(FLOWS, np.float32) - float(G["temp_32"])) < 1e-5 * float(G["temp_32"])


import pytest  # noqa: E402

REF = os.environ.get("TEF_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "loss")), reason="reference not mounted")
@pytest.mark.parametrize("seed", range(12))
def test_priors_match_live_reference(seed):
    """Random shapes against the unmodified reference imported live (build container only), fp64."""
    import copy
    import warnings

    import torch

    sys.path.insert(0, ROOT)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    from loss.flow import Iterative  # reference
    from taming_event_flow_b200 import synthetic as syn

    r = np.random.default_rng(40 + seed)
    B, P, F = int(r.integers(1, 4)), int(r.integers(2, 6)), int(r.integers(1, 3))
    H, W = int(r.integers(3, 40)), int(r.integers(3, 50))
    flows = [[r.normal(0, float(r.choice([0.5, 3.0, 8.0])), (B, 2, H, W)) for _ in range(F)] for _ in range(P)]
    cfg = syn.loss_config(H, W, B, P, 1, "two")
    cfg["loss"]["flow_spat_smooth_weight"], cfg["loss"]["flow_temp_smooth_weight"] = 1.0, 1.0
    torch.set_default_dtype(torch.float64)
    try:
        m = Iterative(copy.deepcopy(cfg), "cpu")
        for t in range(P):
            m.update([torch.from_numpy(f) for f in flows[t]], torch.zeros(B, 0, 4), torch.zeros(B, 0, 2), torch.zeros(B, 0, 4), torch.zeros(B, 0, 2))
        spat, temp = m.flow_spatial_smoothing().item(), m.flow_temporal_smoothing().item()
    finally:
        torch.set_default_dtype(torch.float32)
    assert abs(so.flow_spatial_smoothing(flows) - spat) <= 1e-11 * abs(spat)
    assert abs(so.flow_temporal_smoothing(flows) - temp) <= 1e-10 * abs(temp)
