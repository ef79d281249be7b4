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
