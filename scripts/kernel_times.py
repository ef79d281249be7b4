"""Per-kernel device times (library ProfScope events) and step time for one workload; used to compare kernel variants:
    TEF_B200_LIB=build_variants/lib_x.so python scripts/kernel_times.py --workload iterative_480x640_1Mev"""
import argparse
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--events", type=int, default=None)
    ap.add_argument("--no-prof", action="store_true", help="time the steps without the library's per-kernel CUDA events")
    args = ap.parse_args()
    from taming_event_flow_b200 import _lib, synthetic as syn
    from taming_event_flow_b200.loss import flow as tef_flow

    wl = dict(bench.WORKLOADS[args.workload], name=args.workload)
    seq = bench.fast_sequence(100, wl, n_override=args.events)
    dev = torch.device("cuda", 0)
    cfg = syn.loss_config(wl["H"], wl["W"], wl["B"], wl["P"], wl["S"], wl["mode"], warping=wl["warping"])
    module = getattr(tef_flow, wl["warping"])(cfg, dev)
    flows = [[f.to(dev).requires_grad_(True) for f in per] for per in seq["flows"]]
    masks = [m.to(dev) for m in seq["masks"]]
    dmasks = [m.to(dev) for m in seq["d_masks"]]
    n = args.steps + 2
    evs = [[e.to(dev) for e in seq["events"]] for _ in range(n)]
    devs = [[e.to(dev) for e in seq["d_events"]] for _ in range(n)]
    L = _lib.lib()

    def step(i):
        module.reset()
        for t in range(wl["P"]):
            module.update(flows[t], evs[i][t], masks[t], devs[i][t], dmasks[t])
        loss = module()
        loss.backward()
        for per in flows:
            for f in per:
                f.grad = None
        return loss

    for i in range(2):
        step(i)
    torch.cuda.synchronize()
    L.tef_prof_reset()
    L.tef_prof_enable(0 if args.no_prof else 1)
    t0 = time.perf_counter()
    for i in range(args.steps):
        loss = step(2 + i)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    L.tef_prof_enable(0)
    L.tef_prof_name.restype = ctypes.c_char_p
    out = []
    tot_all = 0.0
    for k in range(L.tef_prof_num_kernels()):
        tot, timed, cnt = ctypes.c_double(), ctypes.c_long(), ctypes.c_long()
        L.tef_prof_read(k, ctypes.byref(tot), ctypes.byref(timed), ctypes.byref(cnt))
        if timed.value:
            out.append("%s=%.4f" % (L.tef_prof_name(k).decode().replace("_kernel", ""), tot.value / timed.value))
            tot_all += tot.value
    ev = bench.events_per_step(wl, args.events)
    print("%s lib=%s step=%.3fms (%.0f Mev/s) kernels_sum=%.3fms loss=%.7g | %s" % (
        args.workload, os.path.basename(_lib.LIB_PATH), dt * 1e3, ev / dt / 1e6, tot_all / args.steps, loss.item(), " ".join(out)))


if __name__ == "__main__":
    main()
