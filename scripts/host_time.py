"""Host-side time of update() / forward / backward for a small workload (no profiler, perf_counter around un-synchronised calls)."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="iterative_128x128_b8_f1")
ap.add_argument("--steps", type=int, default=200)
args = ap.parse_args()
from taming_event_flow_b200 import synthetic as syn
from taming_event_flow_b200.loss import flow as tef_flow
wl = dict(bench.WORKLOADS[args.workload], name=args.workload)
seq = bench.fast_sequence(100, wl)
dev = torch.device("cuda", 0)
cfg = syn.loss_config(wl["H"], wl["W"], wl["B"], wl["P"], wl["S"], wl["mode"], warping=wl["warping"])
module = getattr(tef_flow, wl["warping"])(cfg, dev)
flows = [[f.to(dev).requires_grad_(True) for f in per] for per in seq["flows"]]
masks = [m.to(dev) for m in seq["masks"]]; dmasks = [m.to(dev) for m in seq["d_masks"]]
evs = [e.to(dev) for e in seq["events"]]; devs = [e.to(dev) for e in seq["d_events"]]
tu = tf = tb = 0.0
for i in range(args.steps + 5):
    if i == 5:
        tu = tf = tb = 0.0
    torch.cuda.synchronize()                 # host time only: the GPU is idle when we start
    t0 = time.perf_counter()
    module.reset()
    for t in range(wl["P"]):
        module.update(flows[t], evs[t], masks[t], devs[t], dmasks[t])
    t1 = time.perf_counter()
    loss = module()
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    tu += t1 - t0; tf += t2 - t1; tb += t3 - t2
    for per in flows:
        for f in per:
            f.grad = None
n = args.steps
print("%s host us/step: update x%d %.1f (%.1f each), forward %.1f, backward %.1f, total %.1f" % (
    args.workload, wl["P"], tu / n * 1e6, tu / n / wl["P"] * 1e6, tf / n * 1e6, tb / n * 1e6, (tu + tf + tb) / n * 1e6))
