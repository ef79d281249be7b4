"""How much of the training step is GPU-busy? (kernel time summed by torch.profiler vs wall time per step)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import argparse
import torch
import bench

args = argparse.Namespace(steps=5, warmup=3)
wl = dict(bench.TRAIN_WORKLOADS["train_128x128_b8"], name="train_128x128_b8")
res = bench.run_train(args, wl, quiet=True)
print("baseline ms/step", res["ms_per_step"])

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    res = bench.run_train(argparse.Namespace(steps=2, warmup=1), wl, quiet=True)
ka = prof.key_averages()
tot = sum(k.device_time_total for k in ka if k.device_type == torch.autograd.DeviceType.CUDA)
n = sum(k.count for k in ka if k.device_type == torch.autograd.DeviceType.CUDA)
print("cuda kernels: %d launches over 3 steps, %.2f ms busy per step" % (n, tot / 3 / 1e3))
rows = sorted([k for k in ka if k.device_type == torch.autograd.DeviceType.CUDA], key=lambda k: -k.device_time_total)[:18]
for k in rows:
    print("%9.3f ms/step %6d  %s" % (k.device_time_total / 3 / 1e3, k.count // 3, k.key[:90]))
