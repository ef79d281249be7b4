"""ctypes/numpy front-end of the CPU oracle (``oracle/cm_oracle.c``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs as the checker and
CPU baseline.  Nothing under ``taming_event_flow_b200/`` imports this module.

The functions mirror the reference's call pattern (``update`` once per pass,
then one loss evaluation, upstream ``loss/flow.py:443-476,588-746``) but take
everything at once:

* ``flows[t][f]``     ``[B,2,H,W]``  flow map of pass ``t``, flow scale ``f`` (ch0 = x, ch1 = y)
* ``events[t]``       ``[B,N_t,4]``  (ts, y, x, p) with ts in [0, 1] (raw, before ``+= pass``)
* ``masks[t]``        ``[B,N_t,2]``  (pos, neg)
* ``d_events/d_masks`` the detached lists, same layout
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libcm_oracle.so")
_lib = None

MODES = {"one": 1, "two": 2, "four": 4}


class OrcCfg(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int) for k in ("B", "H", "W", "P", "F", "S", "mode", "border_comp", "loss_scaling", "round_ts")]


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc, a second or two)."""
    src_newer = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("cm_oracle.c", "cm_oracle_impl.h", "cm_oracle.h", "Makefile")
    )
    if force or src_newer:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def set_threads(n=0):
    """Set (n > 0) and return the OpenMP team size the oracle runs with."""
    return int(lib().orc_set_threads(int(n)))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _np(x, dtype):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=dtype)


def _pack_events(ev_list, mk_list, B, dtype):
    n = np.array([int(np.shape(e)[1]) for e in ev_list], dtype=np.int32)
    ev = [_np(e, dtype).reshape(-1, 4) for e in ev_list]
    mk = [_np(m, dtype).reshape(-1, 2) for m in mk_list]
    for e in ev_list:
        assert np.shape(e)[0] == B
    ev = np.concatenate(ev, 0) if len(ev) else np.zeros((0, 4), dtype)
    mk = np.concatenate(mk, 0) if len(mk) else np.zeros((0, 2), dtype)
    if ev.shape[0] == 0:  # keep a valid pointer
        ev = np.zeros((1, 4), dtype)[:0]
        mk = np.zeros((1, 2), dtype)[:0]
    return np.ascontiguousarray(ev), np.ascontiguousarray(mk), n


def make_cfg(B, H, W, P, F, scales_loss=1, iterative_mode="two", border_compensation=True, loss_scaling=True, round_ts=False):
    return OrcCfg(B, H, W, P, F, scales_loss, MODES[iterative_mode], int(border_compensation), int(loss_scaling), int(round_ts))


def num_slots(cfg, linear=False):
    return lib().orc_num_slots(ctypes.byref(cfg), int(linear))


class OracleError(RuntimeError):
    pass


_ERRORS = {
    -1: "configuration outside the oracle's static limits",
    -2: "empty window list (reference: torch.cat of an empty list raises RuntimeError)",
    -3: "round_ts on an empty event tensor (reference: min() of an empty tensor raises)",
    -4: "iterative_mode 'four' with border compensation (reference raises TypeError)",
}


def _run(kind, cfg, flows, events, masks, d_events, d_masks, dtype, want_grad, want_iwe, want_nodes, exact_sums=False):
    L = lib()
    sfx = "f32" if dtype == np.float32 else "f64"
    if exact_sums:
        assert dtype == np.float32, "exact_sums: fp32 per-event arithmetic with double accumulators"
        sfx = "f32x"
    P, F, B, H, W = cfg.P, cfg.F, cfg.B, cfg.H, cfg.W
    assert len(flows) >= P and len(events) >= P
    flow = np.empty((F, P, B, 2, H, W), dtype)
    for t in range(P):
        assert len(flows[t]) == F
        for f in range(F):
            flow[f, t] = _np(flows[t][f], dtype)
    ev, mk, n = _pack_events(events[:P], masks[:P], B, dtype)
    dev, dmk, dn = _pack_events(d_events[:P], d_masks[:P], B, dtype)
    loss = np.zeros(1, dtype)
    gflow = np.zeros_like(flow) if want_grad else None
    ns = num_slots(cfg, linear=(kind == "linear"))
    iwe = np.zeros((F, B, ns, 4, H, W), dtype) if want_iwe else None
    out = {}
    if kind == "iterative":
        E = ev.shape[0]
        nodes = np.zeros((F, P + 1, max(E, 1), 2), dtype) if want_nodes else None
        alive = np.zeros((F, P + 1, max(E, 1)), np.uint8) if want_nodes else None
        rc = getattr(L, "orc_iterative_" + sfx)(
            ctypes.byref(cfg), _ptr(flow), _ptr(ev), _ptr(mk), _ptr(n), _ptr(dev), _ptr(dmk), _ptr(dn),
            _ptr(loss), _ptr(gflow), _ptr(iwe), _ptr(nodes), _ptr(alive))
        if want_nodes:
            out["nodes"] = nodes[:, :, :E]
            out["alive"] = alive[:, :, :E]
    else:
        rc = getattr(L, "orc_linear_" + sfx)(
            ctypes.byref(cfg), _ptr(flow), _ptr(ev), _ptr(mk), _ptr(n), _ptr(dev), _ptr(dmk), _ptr(dn),
            _ptr(loss), _ptr(gflow), _ptr(iwe))
    if rc != 0:
        raise OracleError(_ERRORS.get(rc, "oracle error %d" % rc))
    out["loss"] = loss[0]
    if want_grad:
        # back to the reference's shape: grads[t][f] is [B,2,H,W]
        out["grads"] = [[gflow[f, t] for f in range(F)] for t in range(P)]
        out["gflow"] = gflow
    if want_iwe:
        out["iwe"] = iwe
    return out


def iterative(cfg, flows, events, masks, d_events, d_masks, dtype=np.float32, want_grad=True, want_iwe=False, want_nodes=False, exact_sums=False):
    """``Iterative`` loss forward (+ analytic backward), upstream ``loss/flow.py:415-746``.

    ``exact_sums=True`` is NOT the reference: same fp32 per-event arithmetic, but every image pixel and flow-gradient tap is
    summed in double and rounded to fp32 once -- the yardstick that is free of any fp32 summation order."""
    return _run("iterative", cfg, flows, events, masks, d_events, d_masks, dtype, want_grad, want_iwe, want_nodes, exact_sums)


def linear(cfg, flows, events, masks, d_events, d_masks, dtype=np.float32, want_grad=True, want_iwe=False, exact_sums=False):
    """``Linear`` loss forward (+ analytic backward), upstream ``loss/flow.py:216-412``."""
    return _run("linear", cfg, flows, events, masks, d_events, d_masks, dtype, want_grad, want_iwe, False, exact_sums)


# ---------------------------------------------------------------------------
# stand-alone primitives (upstream utils/iwe.py, dataloader/encodings.py)
# ---------------------------------------------------------------------------
def _sfx(dtype):
    return "f32" if dtype == np.float32 else "f64"


def get_event_flow(flow_map_x, flow_map_y, event_loc, dtype=np.float32):
    mx, my, loc = _np(flow_map_x, dtype), _np(flow_map_y, dtype), _np(event_loc, dtype)
    B, H, W = mx.shape
    N = loc.shape[1]
    out = np.zeros((B, N, 2), dtype)
    getattr(lib(), "orc_get_event_flow_" + _sfx(dtype))(_ptr(mx), _ptr(my), _ptr(loc), _ptr(out), B, N, H, W)
    return out


def event_propagation(events_ts, events_idx, flow, tref, dtype=np.float32):
    ts, loc, fl = _np(events_ts, dtype), _np(events_idx, dtype), _np(flow, dtype)
    out = np.zeros_like(loc)
    ct = ctypes.c_float if dtype == np.float32 else ctypes.c_double
    getattr(lib(), "orc_event_propagation_" + _sfx(dtype))(_ptr(ts), _ptr(loc), _ptr(fl), ct(tref), _ptr(out), ctypes.c_long(ts.size))
    return out


def purge_unfeasible(event_loc, event_pol_mask, res, dtype=np.float32):
    loc, mk = _np(event_loc, dtype).copy(), _np(event_pol_mask, dtype).copy()
    getattr(lib(), "orc_purge_unfeasible_" + _sfx(dtype))(_ptr(loc), _ptr(mk), ctypes.c_long(loc.size // 2), int(res[0]), int(res[1]))
    return loc, mk


def get_interpolation(warped_events, res, round_idx=False, dtype=np.float32):
    w = _np(warped_events, dtype)
    B, N, _ = w.shape
    M = N if round_idx else 4 * N
    idx, wt = np.zeros((B, M, 1), dtype), np.zeros((B, M, 1), dtype)
    getattr(lib(), "orc_get_interpolation_" + _sfx(dtype))(_ptr(w), _ptr(idx), _ptr(wt), B, N, int(res[0]), int(res[1]), int(round_idx))
    return idx, wt


def interpolate(idx, weights, res, polarity_mask=None, zeros=None, dtype=np.float32):
    idx, wt = _np(idx, dtype), _np(weights, dtype)
    B, M = idx.shape[0], idx.shape[1]
    pol = _np(polarity_mask, dtype) if polarity_mask is not None else None
    iwe = np.zeros((B, res[0] * res[1]), dtype) if zeros is None else _np(zeros, dtype).reshape(B, -1).copy()
    getattr(lib(), "orc_interpolate_" + _sfx(dtype))(_ptr(idx), _ptr(wt), _ptr(pol), _ptr(iwe), B, ctypes.c_long(M), int(res[0]), int(res[1]))
    return iwe.reshape(B, 1, res[0], res[1])


def deblur_events(flow, event_list, res, round_idx=True, polarity_mask=None, round_flow=True, dtype=np.float32):
    fl, ev = _np(flow, dtype), _np(event_list, dtype)
    B, N = ev.shape[0], ev.shape[1]
    pol = _np(polarity_mask, dtype) if polarity_mask is not None else None
    iwe = np.zeros((B, res[0] * res[1]), dtype)
    getattr(lib(), "orc_deblur_events_" + _sfx(dtype))(_ptr(fl), _ptr(ev), _ptr(pol), _ptr(iwe), B, N, int(res[0]), int(res[1]), int(round_idx), int(round_flow))
    return iwe.reshape(B, 1, res[0], res[1])


def compute_pol_iwe(flow, event_list, res, pol_mask, round_idx=True, round_flow=True, dtype=np.float32):
    pm = _np(pol_mask, dtype)
    pos = deblur_events(flow, event_list, res, round_idx, pm[:, :, 0:1], round_flow, dtype)
    neg = deblur_events(flow, event_list, res, round_idx, pm[:, :, 1:2], round_flow, dtype)
    return np.concatenate([pos, neg], 1)


def _enc_check(rc, sensor_size):
    if rc != 0:
        raise IndexError("event coordinates out of bounds for a sensor of size %s (reference: index_put_ raises IndexError)" % (tuple(sensor_size),))


def events_to_image(xs, ys, ps, sensor_size=(180, 240), accumulate=True, dtype=np.float32):
    xs, ys, ps = _np(xs, dtype), _np(ys, dtype), _np(ps, dtype)
    img = np.zeros(tuple(sensor_size), dtype)
    _enc_check(getattr(lib(), "orc_events_to_image_" + _sfx(dtype))(_ptr(xs), _ptr(ys), _ptr(ps), _ptr(img), ctypes.c_long(xs.size), int(sensor_size[0]),
                                                                     int(sensor_size[1]), int(bool(accumulate))), sensor_size)
    return img


def events_to_channels(xs, ys, ps, sensor_size=(180, 240), dtype=np.float32):
    xs, ys, ps = _np(xs, dtype), _np(ys, dtype), _np(ps, dtype)
    out = np.zeros((2,) + tuple(sensor_size), dtype)
    _enc_check(getattr(lib(), "orc_events_to_channels_" + _sfx(dtype))(_ptr(xs), _ptr(ys), _ptr(ps), _ptr(out), ctypes.c_long(xs.size), int(sensor_size[0]),
                                                                        int(sensor_size[1])), sensor_size)
    return out


def events_to_voxel(xs, ys, ts, ps, num_bins, sensor_size=(180, 240), dtype=np.float32):
    xs, ys, ts, ps = _np(xs, dtype), _np(ys, dtype), _np(ts, dtype), _np(ps, dtype)
    out = np.zeros((num_bins,) + tuple(sensor_size), dtype)
    _enc_check(getattr(lib(), "orc_events_to_voxel_" + _sfx(dtype))(_ptr(xs), _ptr(ys), _ptr(ts), _ptr(ps), _ptr(out), ctypes.c_long(xs.size), int(num_bins),
                                                                     int(sensor_size[0]), int(sensor_size[1])), sensor_size)
    return out


def get_hot_event_mask(event_rate, idx, max_px=100, min_obvs=5, max_rate=0.8, dtype=np.float32):
    """PARITY UNPINNED: not in the reference; restates tudelft/event_flow's published routine.  Returns (mask, updated rate)."""
    rate = _np(event_rate, dtype).copy()
    mask = np.zeros_like(rate)
    ct = ctypes.c_float if dtype == np.float32 else ctypes.c_double
    getattr(lib(), "orc_get_hot_event_mask_" + _sfx(dtype))(_ptr(rate), _ptr(mask), int(rate.size), int(idx), int(max_px), int(min_obvs), ct(max_rate))
    return mask, rate
