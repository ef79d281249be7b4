"""Kernel breakdown of the GRAPHED training step (train_128x128_b8): which kernels the replayed graph spends its time in.

    python scripts/train_kernels.py [--dtype f32|bf16] [--top 45]

torch.profiler (CUPTI) sees the kernels of a graph replay; busy time and launch counts are per optimizer step."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--workload", default="train_128x128_b8")
    a = ap.parse_args()
    wl = dict(bench.TRAIN_WORKLOADS[a.workload], name=a.workload)
    base = bench.run_train(argparse.Namespace(steps=5, warmup=3, train_mode="graph", train_dtype=a.dtype), wl, quiet=True)
    print("graph %s: %.2f ms/step" % (a.dtype, base["ms_per_step"]))
    from torch.profiler import ProfilerActivity, profile

    steps = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        bench.run_train(argparse.Namespace(steps=steps, warmup=0, train_mode="graph", train_dtype=a.dtype), wl, quiet=True)
    # the capture's warm-up steps run eagerly inside run_train as well: count only what repeats per replay by dividing the
    # totals of the profile by the number of executed steps (2 eager warm-up steps + `steps` replays)
    ka = [k for k in prof.key_averages() if k.device_type == torch.autograd.DeviceType.CUDA]
    n_exec = steps + 2
    tot = sum(k.device_time_total for k in ka)
    cnt = sum(k.count for k in ka)
    print("%d launches, %.2f ms GPU-busy per executed step (%d steps: 2 eager warm-up + %d replays)" % (cnt // n_exec, tot / n_exec / 1e3, n_exec, steps))
    for k in sorted(ka, key=lambda k: -k.device_time_total)[:a.top]:
        print("%8.3f ms/step %6d x %7.2f us  %s" % (k.device_time_total / n_exec / 1e3, k.count // n_exec, k.device_time_total / max(k.count, 1), k.key[:110]))


if __name__ == "__main__":
    main()
