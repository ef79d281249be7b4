"""GPU parity of the loader -> loss contract (taming_event_flow_b200/dataloader/base.py) against the
golden vectors of the unmodified reference and the numpy oracle.  Bit-exact (NaN payloads aside: a one-event window
normalises to 0/0 in the reference too)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import loader_oracle as lo  # noqa: E402
from taming_event_flow_b200 import synthetic as syn  # noqa: E402
from taming_event_flow_b200.dataloader import base as tef_base  # noqa: E402
from taming_event_flow_b200.loss import flow as tef_flow  # noqa: E402

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(ROOT, "tests", "golden", "loader.npz"))
NB = len(G["counts"])
RES = (int(G["H"]), int(G["W"]))


def same_bits(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    if a.shape != b.shape or not np.array_equal(np.isnan(a), np.isnan(b)):
        return False
    ok = ~np.isnan(a)
    return np.array_equal(a.view(np.uint32)[ok], b.view(np.uint32)[ok])


def raw(b):
    return G["xs%d" % b], G["ys%d" % b], G["ts%d" % b], G["ps%d" % b]


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_format_windows_matches_reference_route():
    out = tef_base.format_windows([tef_base.pack_events(*raw(b)) for b in range(NB)], RES, "cuda")
    assert same_bits(out["event_list"], G["collate_event_list"])
    assert same_bits(out["event_list_pol_mask"], G["collate_event_list_pol_mask"])
    assert same_bits(out["event_cnt"], G["collate_event_cnt"])
    assert same_bits(out["event_mask"], np.stack([G["emask%d" % b] for b in range(NB)]))
    assert out["d_event_list"].shape == (NB, 0, 4) and out["d_event_list_pol_mask"].shape == (NB, 0, 2)
    no_cnt = tef_base.format_windows([tef_base.pack_events(*raw(0))], RES, "cuda", with_cnt=False)
    assert "event_cnt" not in no_cnt and same_bits(no_cnt["event_list"][0].t(), G["list0"])


def test_format_windows_large_seeded_vs_oracle():
    rng = np.random.default_rng(5)
    H, W = 480, 640
    wins = []
    for n in (200_000, 0, 123_457):
        wins.append((rng.integers(0, W, n), rng.integers(0, H, n), np.sort(rng.uniform(0, 1e5, n)), rng.integers(0, 2, n)))
    want = lo.format_windows(wins, (H, W))
    got = tef_base.format_windows([tef_base.pack_events(*w) for w in wins], (H, W), "cuda")
    for k in ("event_list", "event_list_pol_mask", "event_cnt", "event_mask"):
        assert same_bits(got[k], want[k]), k


@pytest.mark.parametrize("b", [0, 3])
def test_static_methods_match_reference(b):
    fp = cuda(G["fmt_ps%d" % b])
    assert same_bits(tef_base.BaseDataLoader.create_polarity_mask(fp), G["mask%d" % b])
    ev = tef_base.create_list_encoding(cuda(G["xs%d" % b].astype(np.float32)), cuda(G["ys%d" % b].astype(np.float32)),
                                       cuda(G["fmt_ts%d" % b]), fp)
    assert same_bits(ev, G["list%d" % b])
    cnt = tef_base.BaseDataLoader.create_cnt_encoding(ev[2], ev[1], ev[3], RES)
    assert same_bits(cnt, G["cnt%d" % b])
    assert same_bits(tef_base.create_mask_encoding(cnt), G["emask%d" % b])
    assert same_bits(tef_base.create_polarity_mask(cuda(np.array([0.0, -0.0, 2.5, -3.0], np.float32))),
                     lo.create_polarity_mask(np.array([0.0, -0.0, 2.5, -3.0], np.float32)))


def test_custom_collate_matches_reference():
    batch = [{"event_list": cuda(G["list%d" % b]), "event_list_pol_mask": cuda(G["mask%d" % b]), "event_cnt": cuda(G["cnt%d" % b]),
              "d_event_list": torch.zeros((4, 0), device="cuda"), "d_event_list_pol_mask": torch.zeros((2, 0), device="cuda"),
              "K_rect": torch.arange(16.0, device="cuda").view(4, 4) + b, "gt": None} for b in range(NB)]
    out = tef_base.custom_collate(batch)
    for k in ("event_list", "event_list_pol_mask", "event_cnt", "d_event_list", "d_event_list_pol_mask"):
        assert same_bits(out[k], G["collate_" + k]), k
    assert out["gt"] is None
    assert torch.equal(out["K_rect"][1], (torch.arange(16.0, device="cuda").view(4, 4) + 1).t())   # upstream transposes 3-D items


def _sorted_cols(a):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else a
    return a[:, np.lexsort(a[::-1])]


def test_split_event_list_partition_and_uniformity():
    ev, mk = cuda(G["list3"]), cuda(G["mask3"])
    gen = torch.Generator().manual_seed(1)
    g, gm, d, dm = tef_base.split_event_list(ev, mk, 500, generator=gen)
    assert g.shape == (4, 500) and gm.shape == (2, 500) and d.shape == (4, 700) and dm.shape == (2, 700)
    both = torch.cat([torch.cat([g, gm]), torch.cat([d, dm])], 1)                  # rows stay paired with their masks
    assert np.array_equal(_sorted_cols(both), _sorted_cols(np.concatenate([G["list3"], G["mask3"]])))
    same = tef_base.split_event_list(ev, mk, 5000)
    assert same[0] is ev and same[2].shape == (4, 0) and same[3].shape == (2, 0)
    assert tef_base.split_event_list(ev, mk, None)[0] is ev
    # every event is sampled with probability k/N: 400 draws of 100 out of 1200 (row 0 made a unique id)
    ev = ev.clone()
    ev[0] = torch.arange(ev.shape[1], device="cuda")
    hits = torch.zeros(ev.shape[1], device="cuda")
    for _ in range(400):
        gi = tef_base.split_event_list(ev, mk, 100, generator=gen)[0]
        hits[gi[0].long()] += 1
    p = hits / 400
    assert abs(p.mean().item() - 100 / 1200) < 1e-6 and p.max().item() < 0.15 and p.min().item() > 0.03   # sd 0.0138


def test_format_windows_ragged_split_feeds_the_loss():
    """format_windows + split on a ragged batch: per-sample partitions are exact, and the CM loss does not depend on
    which route produced its inputs."""
    rng = np.random.default_rng(9)
    H, W, B, P, k = 32, 40, 3, 2, 300
    counts = [500, 120, 301]
    cfg = syn.loss_config(H, W, B, P, 1, "two")
    flows = [torch.from_numpy(rng.normal(0, 2, (B, 2, H, W)).astype(np.float32)).cuda() for _ in range(P)]
    losses = []
    for route in ("packed_split", "oracle_nosplit"):
        rng_w = np.random.default_rng(21)
        mod = tef_flow.Iterative(cfg, torch.device("cuda"))
        for t in range(P):
            wins = [(rng_w.integers(0, W, n), rng_w.integers(0, H, n), np.sort(rng_w.uniform(0, 1e4, n)), rng_w.integers(0, 2, n)) for n in counts]
            want = lo.format_windows(wins, (H, W))
            if route == "packed_split":
                out = tef_base.format_windows([tef_base.pack_events(*w) for w in wins], (H, W), "cuda", max_num_grad_events=k,
                                              generator=torch.Generator().manual_seed(t))
                assert out["event_list"].shape == (B, 300, 4) and out["d_event_list"].shape == (B, 200, 4)
                for b, n in enumerate(counts):
                    rows = torch.cat([torch.cat([out["event_list"][b], out["event_list_pol_mask"][b]], 1),
                                      torch.cat([out["d_event_list"][b], out["d_event_list_pol_mask"][b]], 1)])
                    rows = rows[rows[:, 4:].abs().sum(1) > 0]                     # drop padding rows
                    ref = np.concatenate([want["event_list"][b, :n], want["event_list_pol_mask"][b, :n]], 1)
                    assert np.array_equal(_sorted_cols(rows.t()), _sorted_cols(ref.T)), (t, b)
                    assert (out["event_list"][b, min(n, k):] == 0).all() and (out["d_event_list"][b, max(n - k, 0):] == 0).all()
                assert same_bits(out["event_cnt"], want["event_cnt"])
                mod.update([flows[t]], out["event_list"], out["event_list_pol_mask"], out["d_event_list"], out["d_event_list_pol_mask"])
            else:
                mod.update([flows[t]], cuda(want["event_list"]), cuda(want["event_list_pol_mask"]),
                           torch.zeros((B, 0, 4), device="cuda"), torch.zeros((B, 0, 2), device="cuda"))
        losses.append(mod().item())
    assert abs(losses[0] - losses[1]) <= 1e-5 * abs(losses[1])


@pytest.mark.parametrize("seed", range(10))
def test_format_windows_random_ragged_batches(seed):
    r = np.random.default_rng(800 + seed)
    H, W, B = int(r.integers(2, 60)), int(r.integers(2, 80)), int(r.integers(1, 6))
    wins = []
    for _ in range(B):
        n = int(r.choice([0, 1, 2, int(r.integers(3, 600))]))
        wins.append((r.integers(0, W, n), r.integers(0, H, n), np.sort(r.uniform(1e3, 9e5, n)), r.integers(0, 2, n)))
    want = lo.format_windows(wins, (H, W))
    packed = [tef_base.pack_events(*w) for w in wins]
    got = tef_base.format_windows(packed, (H, W), "cuda")
    for key in ("event_list", "event_list_pol_mask", "event_cnt", "event_mask"):
        assert same_bits(got[key], want[key]), key
    k = int(r.integers(1, 300))
    counts = [len(w[0]) for w in wins]
    out = tef_base.format_windows(packed, (H, W), "cuda", max_num_grad_events=k, generator=torch.Generator().manual_seed(seed))
    if max(counts) <= k:
        assert out["d_event_list"].shape[1] == 0 and same_bits(out["event_list"], want["event_list"])
        return
    assert out["event_list"].shape[1] == max(min(c, k) for c in counts) and out["d_event_list"].shape[1] == max(max(c - k, 0) for c in counts)
    for b, n in enumerate(counts):
        rows = torch.cat([torch.cat([out["event_list"][b], out["event_list_pol_mask"][b]], 1),
                          torch.cat([out["d_event_list"][b], out["d_event_list_pol_mask"][b]], 1)])
        rows = rows[rows[:, 4:].abs().sum(1) > 0].cpu().numpy()
        ref = np.concatenate([want["event_list"][b, :n], want["event_list_pol_mask"][b, :n]], 1)
        assert rows.shape == ref.shape
        if n > 1:                                            # one-event windows normalise to NaN timestamps (0/0), like upstream
            assert np.array_equal(_sorted_cols(rows.T), _sorted_cols(ref.T)), b
