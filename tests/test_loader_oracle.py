"""CPU: the numpy restatement of the loader -> loss contract (oracle/loader_oracle.py) against the golden vectors made
from the unmodified reference (tests/golden/loader.npz), and the host-side wire format."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import loader_oracle as lo  # noqa: E402
from taming_event_flow_b200.dataloader import base as tef_base  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "loader.npz"))
NB = len(G["counts"])


def same_bits(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def raw(b):
    return G["xs%d" % b], G["ys%d" % b], G["ts%d" % b], G["ps%d" % b]


@pytest.mark.parametrize("b", range(NB))
def test_event_formatting_and_lists_match_reference(b):
    x, y, t, p = lo.event_formatting(*raw(b))
    assert same_bits(t, G["fmt_ts%d" % b]) and same_bits(p, G["fmt_ps%d" % b])
    assert same_bits(lo.create_list_encoding(x, y, t, p), G["list%d" % b])
    assert same_bits(lo.create_polarity_mask(p), G["mask%d" % b])              # incl. the -0.0 of positive events
    cnt = lo.events_to_channels(x, y, p, (int(G["H"]), int(G["W"])))
    assert same_bits(cnt, G["cnt%d" % b])
    assert same_bits(lo.create_mask_encoding(cnt), G["emask%d" % b])


def test_collate_matches_reference():
    assert same_bits(lo.collate_events([G["list%d" % b] for b in range(NB)]), G["collate_event_list"])
    assert same_bits(lo.collate_events([G["mask%d" % b] for b in range(NB)]), G["collate_event_list_pol_mask"])
    out = lo.format_windows([raw(b) for b in range(NB)], (int(G["H"]), int(G["W"])))
    assert same_bits(out["event_list"], G["collate_event_list"])
    assert same_bits(out["event_list_pol_mask"], G["collate_event_list_pol_mask"])
    assert same_bits(out["event_cnt"], G["collate_event_cnt"])


def test_reference_split_is_a_partition():
    """What the golden split documents: k sampled columns + the rest, together a permutation of the input."""
    g, d = G["split_g"], G["split_d"]
    assert g.shape == (4, 500) and d.shape == (4, 700)
    both = np.concatenate([g, d], 1)
    key = lambda a: a[:, np.lexsort(a[::-1])]                                   # noqa: E731
    assert np.array_equal(key(both), key(G["list3"]))


def test_pack_events_roundtrip_and_errors():
    for b in range(NB):
        xs, ys, ts, ps = raw(b)
        packed = tef_base.pack_events(xs, ys, ts, ps)
        assert packed.dtype == np.uint64 and packed.shape == xs.shape
        ux, uy, ut, up = lo.unpack_events(packed)
        assert np.array_equal(ux, xs) and np.array_equal(uy, ys) and np.array_equal(up, ps)
        assert same_bits(ut, ts.astype(np.float32))
    with pytest.raises(ValueError):
        tef_base.pack_events([1 << 14], [0], [0.0], [1])
    with pytest.raises(ValueError):
        tef_base.pack_events([0], [0], [0.0], [-1])
    with pytest.raises(AssertionError):
        tef_base.pack_events([0, 1], [0], [0.0], [1])


def test_device_functions_refuse_cpu_tensors():
    import torch

    from taming_event_flow_b200._lib import TefError

    with pytest.raises(TefError):
        tef_base.create_polarity_mask(torch.ones(4))
    with pytest.raises(RuntimeError):
        tef_base.format_windows([np.zeros(0, np.uint64)], (4, 4), "cpu")


REF = os.environ.get("TEF_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "dataloader")), reason="reference not mounted")
@pytest.mark.parametrize("seed", range(10))
def test_loader_oracle_matches_live_reference(seed):
    """Random ragged batches against the unmodified BaseDataLoader static methods imported live (build container only)."""
    import types
    import warnings

    import torch

    if REF not in sys.path:
        sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    from dataloader.base import BaseDataLoader  # reference
    from dataloader.encodings import events_to_channels  # reference

    r = np.random.default_rng(600 + seed)
    H, W = int(r.integers(2, 50)), int(r.integers(2, 70))
    me = types.SimpleNamespace(device=torch.device("cpu"))
    wins, batch = [], []
    for _ in range(int(r.integers(1, 5))):
        n = int(r.integers(0, 400))
        xs, ys = r.integers(0, W, n), r.integers(0, H, n)
        ts, ps = np.sort(r.uniform(1e3, 9e5, n)), r.integers(0, 2, n)
        wins.append((xs, ys, ts, ps))
        fx, fy, ft, fp = BaseDataLoader.event_formatting(me, xs, ys, ts, ps)
        ev, mk = BaseDataLoader.create_list_encoding(fx, fy, ft, fp), BaseDataLoader.create_polarity_mask(fp)
        batch.append({"event_list": ev, "event_list_pol_mask": mk, "event_cnt": events_to_channels(fx, fy, fp, sensor_size=(H, W))})
        packed = tef_base.pack_events(xs, ys, ts, ps)
        ux, uy, ut, up = lo.unpack_events(packed)
        assert np.array_equal(ux, xs) and np.array_equal(uy, ys) and np.array_equal(up, ps) and same_bits(ut, ts.astype(np.float32))
    want = BaseDataLoader.custom_collate(batch)
    got = lo.format_windows(wins, (H, W))
    for k in ("event_list", "event_list_pol_mask", "event_cnt"):
        a, b = got[k], want[k].numpy()
        assert a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b))
        ok = ~np.isnan(a)
        assert np.array_equal(a.view(np.uint32)[ok], b.view(np.uint32)[ok]), k
