"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def loss_case_names(kind=None):
    names = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "loss_*.npz")))
    if kind == "iterative":
        names = [n for n in names if n.startswith("iter")]
    elif kind == "linear":
        names = [n for n in names if n.startswith("lin")]
    elif kind == "cm_only":      # cases the oracle covers (it restates the CM terms, not the smoothness priors)
        names = [n for n in names if "smooth" not in n]
    return names


def load_loss_case(name):
    """Golden loss case -> dict with python lists indexed by pass (the layout `update` receives)."""
    z = np.load(os.path.join(GOLDEN, "loss_%s.npz" % name))
    c = {k: z[k].item() if z[k].ndim == 0 else z[k] for k in z.files if not k[:2] in ("ev", "mk") and not k[:3] in ("dev", "dmk")}
    P, F = int(c["P"]), int(c["F"])
    c.setdefault("smooth_spat", -1.0)
    c.setdefault("smooth_temp", -1.0)
    c["flow_list"] = [[c["flows"][f, t] for f in range(F)] for t in range(P)]
    c["events"] = [z["ev%d" % t] for t in range(P)]
    c["masks"] = [z["mk%d" % t] for t in range(P)]
    c["d_events"] = [z["dev%d" % t] for t in range(P)]
    c["d_masks"] = [z["dmk%d" % t] for t in range(P)]
    return c


def rel_err(a, b):
    """(L-inf, L2) norm-relative error of a against b (element-wise relative error is meaningless here)."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    if a.size == 0:
        return 0.0, 0.0
    return (np.abs(a - b).max() / max(np.abs(b).max(), 1e-300), np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
