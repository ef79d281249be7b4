"""The CPU oracle against the golden vectors made from the unmodified reference
(tests/golden/make_golden.py).  Runs without a GPU."""
import numpy as np
import pytest

from oracle import cm_oracle as orc
from util import GOLDEN, load_loss_case, loss_case_names, rel_err

import os


def _run(c, dtype):
    cfg = orc.make_cfg(c["B"], c["H"], c["W"], c["P"], c["F"], c["S"], c["mode"], bool(c["border"]))
    fn = orc.iterative if c["kind"] == "iterative" else orc.linear
    return fn(cfg, c["flow_list"], c["events"], c["masks"], c["d_events"], c["d_masks"], dtype, want_grad=True, want_iwe=True)


@pytest.mark.parametrize("name", loss_case_names("cm_only"))
def test_loss_oracle_fp64_matches_reference_fp64(name):
    c = load_loss_case(name)
    o = _run(c, np.float64)
    assert abs(o["loss"] - c["loss64"]) <= 1e-12 * abs(c["loss64"])
    linf, l2 = rel_err(o["gflow"], c["grad64"])
    assert linf < 1e-11 and l2 < 1e-11, (linf, l2)


@pytest.mark.parametrize("name", loss_case_names("cm_only"))
def test_loss_oracle_fp32_matches_reference_fp32(name):
    c = load_loss_case(name)
    o = _run(c, np.float32)
    # forward images: same fp32 operations in the same order as the reference's CPU scatter_add_
    assert np.array_equal(o["iwe"], c["iwe32"]), "IWE/IWT images are not bit-identical to the reference"
    assert abs(o["loss"] - c["loss32"]) <= 2e-6 * abs(c["loss32"])
    linf, l2 = rel_err(o["gflow"], c["grad32"])
    assert linf < 1e-5 and l2 < 1e-5, (linf, l2)
    # triangulation: the oracle is as close to the fp64 truth as the reference's own fp32 run
    e_or = rel_err(o["gflow"], c["grad64"])[1]
    e_ref = rel_err(c["grad32"], c["grad64"])[1]
    assert e_or <= 1.05 * e_ref + 1e-6


def _prim():
    return np.load(os.path.join(GOLDEN, "primitives.npz"))


def test_get_event_flow_bit_exact():
    z = _prim()
    out = orc.get_event_flow(z["mapx"], z["mapy"], z["loc"])
    assert np.array_equal(out, z["gef_out"])


def test_event_propagation_and_purge_bit_exact():
    z = _prim()
    out = orc.event_propagation(z["ts"], z["loc"], z["gef_out"], 1.0)
    assert np.array_equal(out, z["prop_out"])
    loc, mk = orc.purge_unfeasible(z["loc"], z["mask"], (int(z["H"]), int(z["W"])))
    assert np.array_equal(loc, z["purge_loc"]) and np.array_equal(mk, z["purge_mask"])


def test_get_interpolation_bit_exact():
    z = _prim()
    res = (int(z["H"]), int(z["W"]))
    idx, w = orc.get_interpolation(z["loc"], res)
    assert np.array_equal(idx, z["gi_idx"]) and np.array_equal(w, z["gi_w"])
    idx, w = orc.get_interpolation(z["loc"], res, round_idx=True)
    assert np.array_equal(idx, z["gi_ridx"]) and np.array_equal(w, z["gi_rw"])


def test_interpolate_bit_exact():
    z = _prim()
    res = (int(z["H"]), int(z["W"]))
    pol4 = np.concatenate([z["mask"][:, :, 0:1]] * 4, 1)
    assert np.array_equal(orc.interpolate(z["gi_idx"], z["gi_w"], res), z["interp_nopol"])
    assert np.array_equal(orc.interpolate(z["gi_idx"], z["gi_w"], res, pol4), z["interp_pol"])
    assert np.array_equal(orc.interpolate(z["gi_idx"], z["gi_w"], res, pol4, zeros=z["interp_zeros_in"]), z["interp_zeros"])


@pytest.mark.parametrize("ri", [True, False])
@pytest.mark.parametrize("rf", [True, False])
def test_compute_pol_iwe_bit_exact(ri, rf):
    z = _prim()
    res = (int(z["H"]), int(z["W"]))
    ev = z["db_ev_int"] if rf else z["db_ev_frac"]
    out = orc.compute_pol_iwe(z["db_flow"], ev, res, z["mask"], round_idx=ri, round_flow=rf)
    assert np.array_equal(out, z["pol_iwe_ri%d_rf%d" % (ri, rf)])


def test_encodings_bit_exact():
    z = np.load(os.path.join(GOLDEN, "encodings.npz"))
    ss = (int(z["H"]), int(z["W"]))
    assert np.array_equal(orc.events_to_image(z["xs"], z["ys"], z["ps"], ss), z["image"])
    assert np.array_equal(orc.events_to_channels(z["xs"], z["ys"], z["ps"], ss), z["channels"])
    assert np.array_equal(orc.events_to_voxel(z["xs"], z["ys"], z["ts"], z["ps"], int(z["bins"]), ss), z["voxel"])
    v64 = orc.events_to_voxel(z["xs"], z["ys"], z["ts"], z["ps"], int(z["bins"]), ss, dtype=np.float64)
    assert np.allclose(v64, z["voxel64"], rtol=0, atol=1e-12)


def test_error_codes_mirror_reference_failures():
    # passes_loss=1 in mode two: delta 0 -> torch.cat of an empty list in the reference
    c = load_loss_case("iter_two_small")
    cfg = orc.make_cfg(c["B"], c["H"], c["W"], 1, c["F"], 1, "two", True)
    with pytest.raises(orc.OracleError):
        orc.iterative(cfg, c["flow_list"], c["events"], c["masks"], c["d_events"], c["d_masks"])
    # mode four with border compensation: TypeError in the reference
    cfg = orc.make_cfg(c["B"], c["H"], c["W"], 4, c["F"], 1, "four", True)
    with pytest.raises(orc.OracleError):
        orc.iterative(cfg, c["flow_list"], c["events"], c["masks"], c["d_events"], c["d_masks"])


def test_oracle_thread_control():
    """bench.py sets the OpenMP team size explicitly (torchrun exports OMP_NUM_THREADS=1) and reports what it got."""
    before = orc.set_threads(0)
    assert orc.set_threads(2) == 2
    assert orc.set_threads(before) == before
