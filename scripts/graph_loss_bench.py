"""Loss window through the eager Python API vs one CUDA-graph replay (taming_event_flow_b200.graphs.GraphedLossWindow):
host-bound small windows (BASELINE.json configs[0]).  Checks that both give the same loss and gradients."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from taming_event_flow_b200 import synthetic as syn  # noqa: E402
from taming_event_flow_b200.graphs import GraphedLossWindow  # noqa: E402
from taming_event_flow_b200.loss import flow as tef_flow  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=50)
a = ap.parse_args()
dev = torch.device("cuda", 0)
for name in ("iterative_128x128_b8_f1", "iterative_128x128_b8_f4", "linear_128x128_b8_f4", "iterative_480x640_100kev"):
    wl = dict(bench.WORKLOADS[name], name=name)
    seq = bench.fast_sequence(100, wl)
    P = wl["P"]
    cfg = syn.loss_config(wl["H"], wl["W"], wl["B"], P, wl["S"], wl["mode"], warping=wl["warping"])
    module = getattr(tef_flow, wl["warping"])(cfg, dev)
    flows = [[f.to(dev).requires_grad_(True) for f in per] for per in seq["flows"]]
    ev, mk, dv, dm = ([x.to(dev) for x in seq[k]] for k in ("events", "masks", "d_events", "d_masks"))

    def eager():
        module.reset()
        for t in range(P):
            module.update(flows[t], ev[t].clone(), mk[t], dv[t].clone(), dm[t])
        loss = module()
        loss.backward()
        g = [f.grad for per in flows for f in per]
        for per in flows:
            for f in per:
                f.grad = None
        return loss, g

    for _ in range(3):
        l_e, g_e = eager()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        eager()
    torch.cuda.synchronize()
    t_e = (time.perf_counter() - t0) / a.steps
    gw = GraphedLossWindow(getattr(tef_flow, wl["warping"])(cfg, dev), flows, ev, mk, dv, dm)
    l_g, g_g = gw.replay()
    torch.cuda.synchronize()
    err = max(float((a_ - b_).abs().max() / (a_.abs().max() + 1e-30)) for a_, b_ in zip(g_e, [f for per in g_g for f in per]))
    t0 = time.perf_counter()
    for _ in range(a.steps):
        gw.replay()
    torch.cuda.synchronize()
    t_g = (time.perf_counter() - t0) / a.steps
    E = bench.events_per_step(wl)
    print("%-28s eager %.3f ms (%6.0f Mev/s)   graph %.3f ms (%6.0f Mev/s)   launches/replay %d   loss %.7g vs %.7g   grad Linf-rel diff %.2e"
          % (name, t_e * 1e3, E / t_e / 1e6, t_g * 1e3, E / t_g / 1e6, gw.launches_per_replay, l_e.item(), l_g.item(), err), flush=True)
