"""Run a few plain steps of one bench.py workload (update x P -> forward -> backward) for ncu / compute-sanitizer.

    ncu --set full --clock-control none --import-source on -k regex:iter_ -s 6 -c 2 -o gpurun_out/prof \\
        python scripts/profile_step.py --workload iterative_480x640_1Mev --steps 5
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--events", type=int, default=None, help="override events per window")
    args = ap.parse_args()
    from taming_event_flow_b200 import synthetic as syn
    from taming_event_flow_b200.loss import flow as tef_flow

    wl = dict(bench.WORKLOADS[args.workload], name=args.workload)
    seq = bench.fast_sequence(100, wl, n_override=args.events)
    dev = torch.device("cuda", 0)
    cfg = syn.loss_config(wl["H"], wl["W"], wl["B"], wl["P"], wl["S"], wl["mode"], warping=wl["warping"])
    module = getattr(tef_flow, wl["warping"])(cfg, dev)
    flows = [[f.to(dev).requires_grad_(True) for f in per] for per in seq["flows"]]
    masks = [m.to(dev) for m in seq["masks"]]
    dmasks = [m.to(dev) for m in seq["d_masks"]]
    for step in range(args.steps):
        module.reset()
        for t in range(wl["P"]):
            module.update(flows[t], seq["events"][t].to(dev), masks[t], seq["d_events"][t].to(dev), dmasks[t])
        loss = module()
        loss.backward()
        torch.cuda.synchronize()
        print("step", step, "loss", loss.item())


if __name__ == "__main__":
    main()
