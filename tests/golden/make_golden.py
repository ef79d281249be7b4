"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

The reference ships no tests or fixtures, so the pin for the
oracle and for the CUDA path is the reference's own eager PyTorch code, imported
live here on the CPU, evaluated on seeded synthetic inputs
(taming_event_flow_b200/synthetic.py) in fp32 and, for triangulation, in fp64
(same code, torch default dtype switched).  Inputs are stored next to the
outputs so the fixtures do not depend on torch's RNG stream.
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("TEF_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from taming_event_flow_b200 import synthetic as syn  # noqa: E402

from loss.flow import Iterative, Linear  # noqa: E402  (reference)
from utils import iwe as ref_iwe  # noqa: E402  (reference)
from dataloader import encodings as ref_enc  # noqa: E402  (reference)


def run_loss(kind, cfg, seq, dtype, border, loss_scaling=True, record_iwe=False):
    torch.set_default_dtype(dtype)
    try:
        cfg = copy.deepcopy(cfg)
        cls = Iterative if kind == "iterative" else Linear
        m = cls(cfg, "cpu", loss_scaling=loss_scaling)
        m.border_compensation = border
        rec = []
        if record_iwe:
            orig = m.iwe_formatting

            def hooked(*a, **k):
                out = orig(*a, **k)
                rec.append((out[0].detach().clone(), out[1].detach().clone()))
                return out

            m.iwe_formatting = hooked
        flows = [[f.to(dtype).clone().requires_grad_(True) for f in fl] for fl in seq["flows"]]
        for t in range(len(flows)):
            m.update(
                flows[t],
                seq["events"][t].to(dtype).clone(),
                seq["masks"][t].to(dtype).clone(),
                seq["d_events"][t].to(dtype).clone(),
                seq["d_masks"][t].to(dtype).clone(),
            )
        loss = m()
        loss.backward()
        P, F = len(flows), len(flows[0])
        g = np.stack([np.stack([flows[t][f].grad.numpy() for t in range(P)]) for f in range(F)])
        iwe = None
        if record_iwe:
            # calls come in (grad, detached) pairs per image slot (Linear: fw, bw, d_fw, d_bw)
            imgs = []
            if kind == "iterative":
                for k in range(0, len(rec), 2):
                    cnt = rec[k][0] + rec[k + 1][0]
                    tsi = rec[k][1] + rec[k + 1][1]
                    imgs.append(torch.cat([cnt, tsi], 1))  # [B,4,H,W]: cnt+,cnt-,ts+,ts-
            else:
                for k in range(0, len(rec), 4):
                    for e in range(2):
                        cnt = rec[k + e][0] + rec[k + 2 + e][0]
                        tsi = rec[k + e][1] + rec[k + 2 + e][1]
                        imgs.append(torch.cat([cnt, tsi], 1))
            nsl = len(imgs) // F
            iwe = torch.stack(imgs).view(F, nsl, *imgs[0].shape).permute(0, 2, 1, 3, 4, 5).contiguous().numpy()
        return loss.item(), g, iwe
    finally:
        torch.set_default_dtype(torch.float32)


LOSS_CASES = [
    # name, kind, B, P, N, Nd, H, W, F, S, mode, sigma, ragged, border, seed
    ("iter_two_small", "iterative", 2, 4, 300, 200, 32, 32, 1, 1, "two", 3.0, False, True, 0),
    ("iter_two_f2_p10", "iterative", 2, 10, 400, 250, 48, 40, 2, 1, "two", 3.0, False, True, 1),
    ("iter_two_s2", "iterative", 2, 8, 300, 200, 32, 40, 1, 2, "two", 3.0, False, True, 2),
    ("iter_one", "iterative", 2, 4, 300, 200, 32, 32, 1, 1, "one", 3.0, False, True, 3),
    ("iter_ragged_nodetached", "iterative", 3, 4, 300, 0, 32, 32, 1, 1, "two", 3.0, True, True, 4),
    ("iter_four_noborder", "iterative", 1, 8, 300, 100, 32, 32, 1, 1, "four", 3.0, False, False, 5),
    ("iter_two_noborder", "iterative", 2, 6, 300, 100, 40, 32, 1, 1, "two", 3.0, False, False, 6),
    ("iter_smallflow", "iterative", 1, 10, 1500, 0, 60, 80, 1, 1, "two", 0.5, False, True, 7),
    ("lin_f2", "linear", 2, 4, 300, 200, 32, 32, 2, 1, "two", 3.0, False, True, 8),
    ("lin_s2_noborder", "linear", 2, 8, 300, 200, 32, 40, 1, 2, "two", 3.0, False, False, 9),
    ("lin_ragged", "linear", 2, 10, 300, 0, 32, 40, 1, 1, "two", 3.0, True, True, 10),
    # smoothness priors switched on (flow_spat_smooth_weight, flow_temp_smooth_weight)
    ("iter_smooth", "iterative", 2, 4, 300, 100, 32, 40, 2, 1, "two", 2.0, False, True, 11, (0.3, 0.7)),
    ("lin_smooth", "linear", 2, 4, 300, 100, 32, 40, 1, 1, "two", 2.0, False, True, 12, (0.5, 0.25)),
]


def make_loss_cases():
    for case in LOSS_CASES:
        (name, kind, B, P, N, Nd, H, W, F, S, mode, sigma, ragged, border, seed) = case[:15]
        smooth = case[15] if len(case) > 15 else (None, None)
        seq = syn.make_sequence(seed, B, P, N, Nd, H, W, F, sigma, ragged)
        P_cfg = P // 2 if (mode == "four" and kind == "iterative") else P  # Iterative.__init__ doubles it (loss/flow.py:422-423)
        cfg = syn.loss_config(H, W, B, P_cfg, S, mode)
        cfg["loss"]["flow_spat_smooth_weight"], cfg["loss"]["flow_temp_smooth_weight"] = smooth
        l32, g32, iwe32 = run_loss(kind, cfg, seq, torch.float32, border, record_iwe=True)
        l64, g64, _ = run_loss(kind, cfg, seq, torch.float64, border)
        out = {
            "kind": kind, "B": B, "P": P, "H": H, "W": W, "F": F, "S": S, "mode": mode, "border": border,
            "smooth_spat": np.float64(smooth[0] if smooth[0] is not None else -1.0),
            "smooth_temp": np.float64(smooth[1] if smooth[1] is not None else -1.0),
            "loss32": np.float32(l32), "loss64": np.float64(l64), "grad32": g32.astype(np.float32), "grad64": g64,
            "iwe32": iwe32.astype(np.float32),
            "flows": np.stack([np.stack([seq["flows"][t][f].numpy() for t in range(P)]) for f in range(F)]),
        }
        for t in range(P):
            out["ev%d" % t] = seq["events"][t].numpy()
            out["mk%d" % t] = seq["masks"][t].numpy()
            out["dev%d" % t] = seq["d_events"][t].numpy()
            out["dmk%d" % t] = seq["d_masks"][t].numpy()
        np.savez_compressed(os.path.join(HERE, "loss_%s.npz" % name), **out)
        print("loss case", name, "loss32=%.8g" % l32, "loss64=%.12g" % l64)


def make_primitive_cases():
    g = torch.Generator().manual_seed(123)
    B, N, H, W = 2, 257, 24, 31
    out = {"B": B, "N": N, "H": H, "W": W}
    loc = torch.rand(B, N, 2, generator=g) * torch.tensor([H + 6.0, W + 6.0]) - 3.0
    loc[:, : N // 4] = torch.floor(loc[:, : N // 4])  # exactly-integer coordinates (tie sub-gradients)
    loc[:, N // 4] = torch.tensor([H - 1.0, W - 1.0])
    loc[:, N // 4 + 1] = torch.tensor([0.0, 0.0])
    loc[:, N // 4 + 2] = torch.tensor([1.0 - 2.0 ** -24, 3.5])  # y+1 rounds up to 2 in fp32
    mapx = torch.randn(B, H, W, generator=g)
    mapy = torch.randn(B, H, W, generator=g)
    ts = torch.rand(B, N, 1, generator=g)
    p = (torch.randint(0, 2, (B, N), generator=g) * 2 - 1).float()
    mask = torch.stack([(p > 0).float(), (p < 0).float()], -1)
    out.update(loc=loc.numpy(), mapx=mapx.numpy(), mapy=mapy.numpy(), ts=ts.numpy(), mask=mask.numpy())

    # get_event_flow forward + backward (grad to maps and to locations)
    mx, my, lc = mapx.clone().requires_grad_(True), mapy.clone().requires_grad_(True), loc.clone().requires_grad_(True)
    fl = ref_iwe.get_event_flow(mx, my, lc)
    gout = torch.randn(fl.shape, generator=g)
    fl.backward(gout)
    out.update(gef_out=fl.detach().numpy(), gef_gout=gout.numpy(), gef_gmapx=mx.grad.numpy(), gef_gmapy=my.grad.numpy(), gef_gloc=lc.grad.numpy())

    # event_propagation
    out["prop_out"] = ref_iwe.event_propagation(ts, loc, fl.detach(), 1).numpy()
    # purge_unfeasible
    pl, pm = ref_iwe.purge_unfeasible(loc, mask, [H, W])
    out.update(purge_loc=pl.numpy(), purge_mask=pm.numpy())
    # get_interpolation (both branches) + its backward through the weights
    lc2 = loc.clone().requires_grad_(True)
    idx, wts = ref_iwe.get_interpolation(lc2, [H, W])
    gw = torch.randn(wts.shape, generator=g)
    wts.backward(gw)
    out.update(gi_idx=idx.detach().numpy(), gi_w=wts.detach().numpy(), gi_gw=gw.numpy(), gi_gloc=lc2.grad.numpy())
    ridx, rw = ref_iwe.get_interpolation(loc.clone(), [H, W], round_idx=True)
    out.update(gi_ridx=ridx.numpy(), gi_rw=rw.numpy())
    # interpolate with and without polarity mask, and with a `zeros` start image
    pol4 = torch.cat([mask[:, :, 0:1]] * 4, 1)
    out["interp_nopol"] = ref_iwe.interpolate(idx.detach(), wts.detach(), [H, W]).numpy()
    out["interp_pol"] = ref_iwe.interpolate(idx.detach(), wts.detach(), [H, W], polarity_mask=pol4).numpy()
    z0 = torch.rand(B, H * W, 1, generator=g)
    out["interp_zeros_in"] = z0.numpy()
    out["interp_zeros"] = ref_iwe.interpolate(idx.detach(), wts.detach(), [H, W], polarity_mask=pol4, zeros=z0).numpy()

    # deblur_events / compute_pol_iwe (all four flag combinations).  round_flow=True indexes the
    # flow with y*W+x un-rounded (utils/iwe.py:185-191), i.e. it assumes integer sensor coordinates,
    # so that branch gets integer locations (some outside the sensor); round_flow=False gets
    # fractional ones.
    flow = torch.stack([mapx, mapy], 1) * 2.0
    loc_i = torch.stack([torch.randint(-2, H + 2, (B, N), generator=g), torch.randint(-2, W + 2, (B, N), generator=g)], -1).float()
    ev_i = torch.cat([ts, loc_i, p.unsqueeze(-1)], -1)
    ev_f = torch.cat([ts, loc, p.unsqueeze(-1)], -1)
    out.update(db_ev_int=ev_i.numpy(), db_ev_frac=ev_f.numpy(), db_flow=flow.numpy())
    for ri in (True, False):
        for rf in (True, False):
            key = "pol_iwe_ri%d_rf%d" % (ri, rf)
            src = ev_i if rf else ev_f
            out[key] = ref_iwe.compute_pol_iwe(flow, src.clone(), [H, W], mask, round_idx=ri, round_flow=rf).numpy()
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), **out)
    print("primitive cases written")


def make_encoding_cases():
    g = torch.Generator().manual_seed(7)
    H, W, N, bins = 36, 52, 5000, 5
    xs = torch.randint(0, W, (N,), generator=g).float()
    ys = torch.randint(0, H, (N,), generator=g).float()
    xs[:50] += 0.75  # .long() truncation (dataloader/encodings.py:23-26)
    ts, _ = torch.sort(torch.rand(N, generator=g))
    ts = (ts - ts[0]) / (ts[-1] - ts[0])
    ps = (torch.randint(0, 2, (N,), generator=g) * 2 - 1).float()
    out = {"H": H, "W": W, "N": N, "bins": bins, "xs": xs.numpy(), "ys": ys.numpy(), "ts": ts.numpy(), "ps": ps.numpy()}
    out["image"] = ref_enc.events_to_image(xs, ys, ps, sensor_size=(H, W)).numpy()
    out["channels"] = ref_enc.events_to_channels(xs, ys, ps, sensor_size=(H, W)).numpy()
    out["voxel"] = ref_enc.events_to_voxel(xs, ys, ts, ps, bins, sensor_size=(H, W)).numpy()
    torch.set_default_dtype(torch.float64)
    out["voxel64"] = ref_enc.events_to_voxel(xs.double(), ys.double(), ts.double(), ps.double(), bins, sensor_size=(H, W)).numpy()
    torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(HERE, "encodings.npz"), **out)
    print("encoding cases written")


def make_flow_val_cases():
    """loss/flow_val.py (first "next" row of SURVEY.md §8f): Linear and Iterative validation on synthetic window sequences."""
    _flow_val_case("flow_val.npz", 40, 56, 5, 1200, 2.0, 8, False, 77)
    _flow_val_case("flow_val_b.npz", 37, 53, 4, 500, 4.0, 4, True, 78)      # odd resolution, larger flow, round_ts


def _flow_val_case(fname, H, W, P, N, sigma, coarse, round_ts, seed):
    from loss import flow_val as ref_val

    gen = torch.Generator().manual_seed(seed)
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"round_ts": round_ts}, "vis": {"mask_output": True}, "metrics": {"name": ["AEE"]}}
    wins = []
    for t in range(P):
        ev, mk = syn.make_window(gen, 1, N, H, W)
        flow = syn.make_flow(gen, 1, H, W, sigma, coarse=coarse)
        emask = (torch.rand(1, 1, H, W, generator=gen) > 0.3).float()
        wins.append((ev, mk, flow, emask))
    gt = syn.make_flow(gen, 1, H, W, sigma, coarse=coarse)
    gt[:, :, :5] = 0.0                                                    # pixels without ground truth
    out = {"H": H, "W": W, "P": P, "gt": gt.numpy(), "round_ts": int(round_ts)}
    for t, (ev, mk, flow, emask) in enumerate(wins):
        out["ev%d" % t], out["mk%d" % t], out["flow%d" % t], out["emask%d" % t] = ev.numpy(), mk.numpy(), flow.numpy(), emask.numpy()
    for name, cls in (("linear", ref_val.Linear), ("iterative", ref_val.Iterative)):
        m = cls(copy.deepcopy(cfg), "cpu")
        for t, (ev, mk, flow, emask) in enumerate(wins):
            m.update([flow.clone()], ev.clone(), mk.clone(), emask.clone())
            if t in (0, 2, P - 1):                                        # metrics after 1, 3 and 5 windows
                key = "%s_t%d_" % (name, t)
                out[key + "fwl"] = m.fwl().numpy()
                out[key + "rsat"] = m.rsat().numpy()
                out[key + "events"] = m.window_events().numpy()
                out[key + "events_round"] = m.window_events(round_idx=True).numpy()
                out[key + "aee"] = m.compute_aee(flow, gt, mask=m._event_mask).numpy()
                modes = (None,) if name == "linear" else (None, "forward", "backward")
                for mode in modes:
                    out[key + "flow_%s" % mode] = m.window_flow(mode=mode).numpy()
                    out[key + "flow_nomask_%s" % mode] = m.window_flow(mode=mode, mask=False).numpy()
                for mode in ((None,) if name == "linear" else ("forward", "backward")):
                    for ri in (False, True):
                        out[key + "iwe_%s_%d" % (mode, ri)] = m.window_iwe(mode=mode, round_idx=ri).numpy()
        m.reset()
        assert m.num_passes == 0
    np.savez_compressed(os.path.join(HERE, fname), **out)
    print("flow_val case written:", fname)


def make_smoothness_cases():
    """loss/flow.py:131-209: each smoothness prior on its own (weight 1), value and gradients from the
    unmodified reference's autograd, fp32 and fp64, plus the value after fewer updates than passes_loss."""
    out = {}
    B, P, H, W, F = 2, 4, 28, 36, 2
    seq = syn.make_sequence(21, B, P, 50, 0, H, W, F_scales=F, sigma=4.0)
    cfg = syn.loss_config(H, W, B, P, 1, "two")
    cfg["loss"]["flow_spat_smooth_weight"], cfg["loss"]["flow_temp_smooth_weight"] = 1.0, 1.0
    out.update({"B": B, "P": P, "H": H, "W": W, "F": F})
    for t in range(P):
        out["ev%d" % t], out["mk%d" % t] = seq["events"][t].numpy(), seq["masks"][t].numpy()
        for f in range(F):
            out["flow%d_%d" % (t, f)] = seq["flows"][t][f].numpy()
    for dtype, tag in ((torch.float32, "32"), (torch.float64, "64")):
        torch.set_default_dtype(dtype)
        try:
            m = Iterative(copy.deepcopy(cfg), "cpu")
            flows = [[f.to(dtype).clone().requires_grad_(True) for f in fl] for fl in seq["flows"]]
            empty_e, empty_m = torch.zeros(B, 0, 4, dtype=dtype), torch.zeros(B, 0, 2, dtype=dtype)
            for t in range(P):
                m.update(flows[t], seq["events"][t].to(dtype).clone(), seq["masks"][t].to(dtype).clone(), empty_e, empty_m)
                if t == 1:
                    out["spat_2passes_" + tag] = m.flow_spatial_smoothing().item()
                    out["temp_2passes_" + tag] = m.flow_temporal_smoothing().item()
            for name, fn in (("spat", m.flow_spatial_smoothing), ("temp", m.flow_temporal_smoothing)):
                for fl in flows:
                    for f in fl:
                        f.grad = None
                val = fn()
                val.backward()
                out["%s_%s" % (name, tag)] = val.item()
                out["%s_grad_%s" % (name, tag)] = np.stack([np.stack([flows[t][f].grad.numpy() for t in range(P)]) for f in range(F)])
        finally:
            torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(HERE, "smoothness.npz"), **out)
    print("smoothness cases written", out["spat_32"], out["temp_32"])


def make_loader_cases():
    """dataloader/base.py: the static methods of the UNMODIFIED BaseDataLoader on three ragged windows."""
    import types

    from dataloader.base import BaseDataLoader  # noqa: E402  (reference)

    rng = np.random.default_rng(11)
    H, W = 40, 56
    me = types.SimpleNamespace(device=torch.device("cpu"))
    out = {"H": H, "W": W, "counts": np.array([700, 1, 0, 1200])}
    batch = []
    for b, n in enumerate(out["counts"]):
        xs = rng.integers(0, W, n).astype(np.int64)
        ys = rng.integers(0, H, n).astype(np.int64)
        ts = np.sort(rng.uniform(2.0e5, 2.5e5, n))                          # raw float64 timestamps (us since file start)
        ps = rng.integers(0, 2, n).astype(np.int64)
        out.update({"xs%d" % b: xs, "ys%d" % b: ys, "ts%d" % b: ts, "ps%d" % b: ps})
        fx, fy, ft, fp = BaseDataLoader.event_formatting(me, xs, ys, ts, ps)
        ev = BaseDataLoader.create_list_encoding(fx, fy, ft, fp)
        mk = BaseDataLoader.create_polarity_mask(fp)
        cnt = ref_enc.events_to_channels(fx, fy, fp, sensor_size=(H, W))
        out.update({"fmt_ts%d" % b: ft.numpy(), "fmt_ps%d" % b: fp.numpy(), "list%d" % b: ev.numpy(), "mask%d" % b: mk.numpy(),
                    "cnt%d" % b: cnt.numpy(), "emask%d" % b: BaseDataLoader.create_mask_encoding(cnt).numpy()})
        batch.append({"event_list": ev, "event_list_pol_mask": mk, "event_cnt": cnt, "d_event_list": torch.zeros((4, 0)),
                      "d_event_list_pol_mask": torch.zeros((2, 0))})
    col = BaseDataLoader.custom_collate(batch)
    for k, v in col.items():
        out["collate_" + k] = v.numpy()
    # split_event_list: random; stored to document shapes and the partition property
    torch.manual_seed(3)
    g, gm, d, dm = BaseDataLoader.split_event_list(batch[3]["event_list"], batch[3]["event_list_pol_mask"], 500)
    out.update({"split_g": g.numpy(), "split_gm": gm.numpy(), "split_d": d.numpy(), "split_dm": dm.numpy()})
    np.savez_compressed(os.path.join(HERE, "loader.npz"), **out)
    print("loader cases written")


if __name__ == "__main__":
    import warnings

    warnings.filterwarnings("ignore")
    make_loss_cases()
    make_primitive_cases()
    make_encoding_cases()
    make_flow_val_cases()
    make_loader_cases()
    make_smoothness_cases()
    print("torch", torch.__version__, "cpu capability", torch.backends.cpu.get_cpu_capability())
