"""cProfile of the host side of a small workload (where Python + launch overhead dominates)."""
import argparse
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="iterative_128x128_b8_f1")
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    from taming_event_flow_b200 import synthetic as syn
    from taming_event_flow_b200.loss import flow as tef_flow

    wl = dict(bench.WORKLOADS[args.workload], name=args.workload)
    seq = bench.fast_sequence(100, wl)
    dev = torch.device("cuda", 0)
    cfg = syn.loss_config(wl["H"], wl["W"], wl["B"], wl["P"], wl["S"], wl["mode"], warping=wl["warping"])
    module = getattr(tef_flow, wl["warping"])(cfg, dev)
    flows = [[f.to(dev).requires_grad_(True) for f in per] for per in seq["flows"]]
    masks = [m.to(dev) for m in seq["masks"]]
    dmasks = [m.to(dev) for m in seq["d_masks"]]
    evs = [[e.to(dev) for e in seq["events"]] for _ in range(args.steps + 3)]
    devs = [[e.to(dev) for e in seq["d_events"]] for _ in range(args.steps + 3)]

    def step(i):
        module.reset()
        for t in range(wl["P"]):
            module.update(flows[t], evs[i][t], masks[t], devs[i][t], dmasks[t])
        loss = module()
        loss.backward()
        for per in flows:
            for f in per:
                f.grad = None

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    for i in range(args.steps):
        step(3 + i)
    pr.disable()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("host time per step %.3f ms, with final sync %.3f ms" % ((t1 - t0) / args.steps * 1e3, (t2 - t0) / args.steps * 1e3))
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()
