#!/usr/bin/env python
"""Benchmark of the contrast-maximization hot path (BASELINE.json: "CM loss fwd+bwd Mevents/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One step = one full loss window through the reference-facing API of
taming_event_flow_b200.loss.flow: `update` x P passes -> `forward` -> `backward` (flow gradients
delivered to every flow map).  Mevents/s = input events of the window (each counted once,
with-gradient + detached) / step time.

* `value`   inputs already resident in HBM when the timed region starts (a fresh device copy of the
            event tensors per step, because `update` mutates the caller's timestamps in place).
* `e2e`     same metric with every input (events, masks, flow maps) in pinned HOST memory and the
            loss + flow gradients read back to the host, all copies inside the timed region.
* `roofline` for the dominant kernel, timed live with CUDA events on the launching stream
            (library-side ProfScope), against MEASURED_PEAKS.json.
* `cpu_baseline` the CPU oracle (oracle/, an OpenMP C port of the reference algorithm) on a bounded
            sample of the same workload, on the host cores of this box.
* `--impl reference` times that CPU port alone (the reference itself is Python/PyTorch and cannot
            travel to the GPU box).

Multi-GPU (torchrun): the batch of independent event-window sequences is sharded, one process per GPU,
no data-path collective (the loss sums over independent samples) -> weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: B, P, N, Nd, H, W, F, S, mode, sigma, distribution, warping
    # north_star target: 1M-event 480x640 windows (BASELINE.json configs[4] at 1M events/window)
    "iterative_480x640_1Mev": dict(B=1, P=10, N=1_000_000, Nd=0, H=480, W=640, F=1, S=1, mode="two", sigma=3.0, dist="uniform", warping="Iterative"),
    "iterative_480x640_100kev": dict(B=1, P=10, N=100_000, Nd=0, H=480, W=640, F=1, S=1, mode="two", sigma=3.0, dist="uniform", warping="Iterative"),
    "iterative_480x640_4Mev": dict(B=1, P=10, N=4_000_000, Nd=0, H=480, W=640, F=1, S=1, mode="two", sigma=3.0, dist="uniform", warping="Iterative"),
    "iterative_480x640_250kev": dict(B=1, P=10, N=250_000, Nd=0, H=480, W=640, F=1, S=1, mode="two", sigma=3.0, dist="uniform", warping="Iterative"),
    "iterative_480x640_500kev": dict(B=1, P=10, N=500_000, Nd=0, H=480, W=640, F=1, S=1, mode="two", sigma=3.0, dist="uniform", warping="Iterative"),
    "iterative_480x640_2Mev": dict(B=1, P=10, N=2_000_000, Nd=0, H=480, W=640, F=1, S=1, mode="two", sigma=3.0, dist="uniform", warping="Iterative"),
    "iterative_240x320_1Mev": dict(B=1, P=10, N=1_000_000, Nd=0, H=240, W=320, F=1, S=1, mode="two", sigma=1.5, dist="uniform", warping="Iterative"),
    "iterative_480x640_1Mev_edges": dict(B=1, P=10, N=1_000_000, Nd=0, H=480, W=640, F=1, S=1, mode="two", sigma=3.0, dist="edges", warping="Iterative"),
    # SURVEY 8d's other input variants: small flow (sigma 0.5 px/window) and ragged batches (per-sample counts U[0.3 N, N], zero-padded)
    "iterative_480x640_1Mev_sigma05": dict(B=1, P=10, N=1_000_000, Nd=0, H=480, W=640, F=1, S=1, mode="two", sigma=0.5, dist="uniform", warping="Iterative"),
    "iterative_128x128_b8_f4_ragged": dict(B=8, P=10, N=10_000, Nd=10_000, H=128, W=128, F=4, S=1, mode="two", sigma=3.0, dist="uniform", warping="Iterative",
                                           ragged=True),
    # BASELINE.json configs[0]: 128x128 crops, batch 8, 10 passes x (10k grad + 10k detached)
    "iterative_128x128_b8_f1": dict(B=8, P=10, N=10_000, Nd=10_000, H=128, W=128, F=1, S=1, mode="two", sigma=3.0, dist="uniform", warping="Iterative"),
    "iterative_128x128_b8_f4": dict(B=8, P=10, N=10_000, Nd=10_000, H=128, W=128, F=4, S=1, mode="two", sigma=3.0, dist="uniform", warping="Iterative"),
    "linear_128x128_b8_f4": dict(B=8, P=10, N=10_000, Nd=10_000, H=128, W=128, F=4, S=1, mode="two", sigma=3.0, dist="uniform", warping="Linear"),
    "linear_480x640_1Mev": dict(B=1, P=10, N=1_000_000, Nd=0, H=480, W=640, F=1, S=1, mode="two", sigma=3.0, dist="uniform", warping="Linear"),
}
# BASELINE.json configs[1] / configs[3]: recurrent EV-FlowNet (PyTorch) + CM loss training step; metric = windows/s
TRAIN_WORKLOADS = {
    "train_128x128_b8": dict(B=8, P=10, N=10_000, Nd=10_000, H=128, W=128, F=4, S=1, mode="two", sigma=3.0, dist="uniform", warping="Iterative", scaling="weak"),
    "train_128x128_gb64": dict(B=64, P=10, N=10_000, Nd=10_000, H=128, W=128, F=4, S=1, mode="two", sigma=3.0, dist="uniform", warping="Iterative", scaling="strong"),
}
# BASELINE.json configs[2]: sequential inference at DSEC resolution with the events_to_voxel encoding (5 bins)
INFER_WORKLOADS = {
    "inference_480x640_100kev": dict(N=100_000, H=480, W=640, bins=5),
    "inference_480x640_500kev": dict(N=500_000, H=480, W=640, bins=5),
    "inference_480x640_1Mev": dict(N=1_000_000, H=480, W=640, bins=5),
}
# the validation criteria of loss/flow_val.py on the CUDA primitives (Iterative flavour, 10-window interval)
VAL_WORKLOADS = {
    "validation_480x640_100kev": dict(N=100_000, H=480, W=640, P=10),
    "validation_480x640_500kev": dict(N=500_000, H=480, W=640, P=10),
}
DEFAULT_WORKLOAD = "iterative_480x640_1Mev"
# training-step settings (measured on B200): "eager" | "graph", "f32" | "bf16"
TRAIN_MODE_DEFAULT = "graph"
TRAIN_DTYPE_DEFAULT = "f32"


def fast_sequence(seed, wl, device="cpu", n_override=None):
    """Synthetic loss window with the statistics of taming_event_flow_b200.synthetic.make_sequence, vectorised
    over the batch so that 10M events are generated in seconds."""
    from taming_event_flow_b200 import synthetic as syn

    B, P, H, W, F = wl["B"], wl["P"], wl["H"], wl["W"], wl["F"]
    N = wl["N"] if n_override is None else n_override
    Nd = wl["Nd"] if n_override is None else (n_override if wl["Nd"] > 0 else 0)
    gen = torch.Generator().manual_seed(seed)
    seq = {"flows": [], "events": [], "masks": [], "d_events": [], "d_masks": []}

    def window(n, t):
        if wl["dist"] != "uniform":
            return syn.make_window(gen, B, n, H, W, False, wl["dist"], t)
        ev = torch.zeros(B, n, 4)
        if n > 0:
            ts, _ = torch.sort(torch.rand(B, n, generator=gen), dim=1)
            if n > 1:
                ts = (ts - ts[:, :1]) / (ts[:, -1:] - ts[:, :1])
            else:
                ts = torch.zeros(B, n)
            ev[:, :, 0] = ts
            ev[:, :, 1] = torch.randint(0, H, (B, n), generator=gen).float()
            ev[:, :, 2] = torch.randint(0, W, (B, n), generator=gen).float()
            ev[:, :, 3] = (torch.randint(0, 2, (B, n), generator=gen) * 2 - 1).float()
            if wl.get("ragged"):                       # zero rows beyond each sample's own count, like custom_collate
                counts = (torch.rand(B, generator=gen) * 0.7 + 0.3) * n
                ev[torch.arange(n)[None, :] >= counts[:, None]] = 0.0
        mk = torch.stack([(ev[:, :, 3] > 0).float(), (ev[:, :, 3] < 0).float()], -1)
        return ev, mk

    for t in range(P):
        seq["flows"].append([syn.make_flow(gen, B, H, W, wl["sigma"]) for _ in range(F)])
        ev, mk = window(N, t)
        dev, dmk = window(Nd, t)
        seq["events"].append(ev)
        seq["masks"].append(mk)
        seq["d_events"].append(dev)
        seq["d_masks"].append(dmk)
    return seq


def events_per_step(wl, n_override=None):
    N = wl["N"] if n_override is None else n_override
    Nd = wl["Nd"] if n_override is None else (n_override if wl["Nd"] > 0 else 0)
    return wl["B"] * wl["P"] * (N + Nd)


# ----------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_port_step(seq, wl):
    from oracle import cm_oracle as orc

    cfg = orc.make_cfg(wl["B"], wl["H"], wl["W"], wl["P"], wl["F"], wl["S"], wl["mode"])
    fn = orc.iterative if wl["warping"] == "Iterative" else orc.linear
    t0 = time.perf_counter()
    out = fn(cfg, seq["flows"], seq["events"], seq["masks"], seq["d_events"], seq["d_masks"], np.float32, want_grad=True)
    return time.perf_counter() - t0, out


def cpu_threads():
    """All host threads this process may use, set explicitly (torchrun exports OMP_NUM_THREADS=1); returns the team size in effect."""
    from oracle import cm_oracle as orc

    try:
        want = len(os.sched_getaffinity(0))
    except AttributeError:
        want = os.cpu_count()
    return orc.set_threads(want)


def cpu_sample_size(wl):
    # bounded sample: same resolution / passes / batch, fewer events per window
    return min(wl["N"], 200_000)


def run_cpu_baseline(wl, repeats=2):
    n = cpu_sample_size(wl)
    seq = fast_sequence(1234, wl, n_override=n)
    threads = cpu_threads()
    cpu_port_step(seq, wl)  # warm-up (page faults, OpenMP pool)
    best = min(cpu_port_step(seq, wl)[0] for _ in range(repeats))
    ev = events_per_step(wl, n)
    return {
        "value": ev / best / 1e6, "unit": "Mevents/s", "cores": threads, "kind": "port",
        "sample": "%s with %d events/window (%d events/step), oracle/cm_oracle.c fwd+bwd, OpenMP on %d threads, best of %d"
                  % (wl["name"], n, ev, threads, repeats),
    }


def reference_root():
    """Where the UNMODIFIED reference can be imported from: the build container mounts it at /root/reference; a pip --target
    install would sit in baseline/_ref.  Neither exists on the GPU box (the reference is plain Python without a setup.py, so
    there is nothing to install), and then the stock arms below are reported as unavailable."""
    for cand in (os.environ.get("TEF_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.exists(os.path.join(cand, "loss", "flow.py")):
            return cand
    return None


def stock_reference_step(ref_mods, wl, seq, device):
    """One loss window through the reference's own classes (upstream loss/flow.py:415-746): update x P, forward, backward."""
    import copy

    from taming_event_flow_b200 import synthetic as syn

    cfg = syn.loss_config(wl["H"], wl["W"], wl["B"], wl["P"], wl["S"], wl["mode"], warping=wl["warping"])
    m = getattr(ref_mods, wl["warping"])(copy.deepcopy(cfg), device)
    flows = [[f.to(device).clone().requires_grad_(True) for f in per] for per in seq["flows"]]
    t0 = time.perf_counter()
    for t in range(wl["P"]):
        m.update(flows[t], seq["events"][t].to(device).clone(), seq["masks"][t].to(device).clone(), seq["d_events"][t].to(device).clone(),
                 seq["d_masks"][t].to(device).clone())
    loss = m()
    loss.backward()
    if device != "cpu":
        torch.cuda.synchronize()
    return time.perf_counter() - t0, float(loss.item())


def run_stock_reference(wl, n, repeats=1):
    """The reference itself (not the port), on the host cores and -- when a GPU is visible -- on CUDA in eager mode
    (SURVEY.md 8d last row).  Returns None where the reference cannot be imported."""
    root = reference_root()
    if root is None:
        return None
    import importlib

    sys.path.insert(0, root)
    try:
        ref_mods = importlib.import_module("loss.flow")
    except Exception as exc:
        return {"unavailable": repr(exc)[:160]}
    finally:
        sys.path.remove(root)
    seq = fast_sequence(1234, wl, n_override=n)
    ev = events_per_step(wl, n)
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count()
    torch.set_num_threads(threads)
    out = {"source": root, "events_per_window": n, "events_per_step": ev}
    for device in ["cpu"] + (["cuda"] if torch.cuda.is_available() else []):
        stock_reference_step(ref_mods, wl, seq, device)             # warm-up
        best = min(stock_reference_step(ref_mods, wl, seq, device)[0] for _ in range(repeats))
        out[device] = {"value": ev / best / 1e6, "unit": "Mevents/s", "s_per_step": best, "threads": threads if device == "cpu" else None,
                       "note": "unmodified reference classes, eager PyTorch %s" % torch.__version__}
    return out


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = cpu_sample_size(wl)
    seq = fast_sequence(1234, wl, n_override=n)
    from oracle import cm_oracle as orc

    orc.lib()                      # compile / load the port before anything is timed
    threads = cpu_threads()
    for _ in range(args.warmup):
        cpu_port_step(seq, wl)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_step(seq, wl)
    dt = time.perf_counter() - t0
    ev = events_per_step(wl, n)
    v = ev * args.steps / dt / 1e6
    sample = "%s with %d events/window (%d events/step), CPU port of the reference algorithm (oracle/cm_oracle.c), OpenMP on %d threads" % (
        wl["name"], n, ev, threads)
    # the reference itself, where it can be imported (build container): a smaller sample, it is ~10x slower than the port
    stock = None
    if os.environ.get("TEF_STOCK_REFERENCE", "1") != "0":
        try:
            stock = run_stock_reference(wl, min(n, 50_000 if wl["H"] * wl["W"] > 128 * 128 else n))
        except Exception as exc:
            stock = {"unavailable": repr(exc)[:160]}
    cfg = workload_config(wl)
    cfg["events_per_window"] = n                       # what this arm actually times: a bounded sample of the workload (a rate, so the ratio holds)
    cfg["events_per_window_of_workload"] = wl["N"]
    line = {
        "impl": "reference", "metric": "cm_loss_fwd_bwd_throughput", "value": v, "unit": "Mevents/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": v, "unit": "Mevents/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mevents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "stock_reference": stock if stock is not None else {"unavailable": "the reference sources are not on this machine (/root/reference exists in the build container only; "
                                                                            "it is plain Python without a setup.py, so nothing can be pip-installed into baseline/_ref)"},
    }
    print(json.dumps(line))


def workload_config(wl):
    return {
        "workload": wl["name"], "warping": wl["warping"], "iterative_mode": wl["mode"], "batch": wl["B"], "passes_loss": wl["P"],
        "events_per_window": wl["N"], "detached_events_per_window": wl["Nd"], "resolution": [wl["H"], wl["W"]],
        "flow_scales": wl["F"], "scales_loss": wl["S"], "event_distribution": wl["dist"], "flow_sigma_px": wl["sigma"],
        "cache": "inputs_larger_than_L2 (fresh event tensors every step)", "ragged": bool(wl.get("ragged", False)),
    }


def with_affinity(cfg, affinity):
    cfg = dict(cfg)
    cfg["cpu_affinity"] = affinity
    return cfg


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML in a thread, every 4 ms).
    Started before the warm-up so that NVML initialisation does not land inside the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.on, self.h, self.nv = index, [], False, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while self.on:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nv is None:
            return
        self.rows, self.on = [], True
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.on = False
        self.th.join()
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        reasons = sorted(k for k, bit in names.items() if any(r & bit for _, r in self.rows))
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


def bind_to_gpu_cpus(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index` (its NUMA node) BEFORE any pinned host memory is
    allocated, so that the staging buffers of the end-to-end arms live next to the GPU's PCIe root.  Returns a description."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and w * 64 + b < ncpu}
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or allowed
        os.sched_setaffinity(0, cpus)
        return "%d CPUs local to GPU %d (of %d allowed)" % (len(cpus), index, len(allowed))
    except Exception as exc:
        return "unchanged (%s)" % type(exc).__name__


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def kernel_bytes(wl, kernel):
    """Algorithmic (compulsory) HBM bytes of one launch of the two event kernels."""
    E = events_per_step(wl)
    Eg = wl["B"] * wl["P"] * wl["N"]
    maps = 8 * wl["F"] * wl["P"] * wl["B"] * wl["H"] * wl["W"]     # packed float2 flow maps
    if kernel in ("iter_fwd_kernel", "linear_fwd_kernel"):
        return 24 * E + maps
    if kernel in ("iter_bwd_kernel", "linear_bwd_kernel"):
        return 24 * Eg + 2 * maps
    raise KeyError(kernel)


def measure_l2_rates(L, dev):
    """Peak lane-op rates of the two operations that bound the CM kernels, measured on this GPU (outside the timed region)."""
    import ctypes

    buf = torch.zeros(64 << 20, dtype=torch.uint8, device=dev)     # L2-resident (126 MB L2)
    out = {}
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for name, kind, mode in (("red_v4_spread", 0, 0), ("red_v4_local", 0, 1), ("red_v4_tile_sorted", 0, 10), ("gather8_spread", 1, 0), ("gather8_local", 1, 1),
                             ("gather16_spread", 2, 0), ("gather16_local", 2, 1)):
        ops = ctypes.c_long()
        best = 1e30
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = L.tef_microbench(kind, mode, ctypes.c_void_p(buf.data_ptr()), ctypes.c_long(buf.numel()), 256, ctypes.byref(ops), st)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0
            best = min(best, e0.elapsed_time(e1))
        out[name] = ops.value / (best * 1e-3) / 1e9     # G lane-ops / s
    return out


def encoding_roofline(L, dev, hbm, N=1_000_000, H=480, W=640, bins=5, reps=20):
    """HBM roofline of the event encodings (dataloader/encodings.py) on one DSEC-resolution window: kernel-only times from the
    library's CUDA events.  Algorithmic bytes: events_to_voxel 16 N + 4 bins H W, events_to_channels 12 N + 8 H W."""
    import ctypes

    from taming_event_flow_b200.dataloader import encodings as enc

    g = torch.Generator().manual_seed(3)
    xs = torch.randint(0, W, (N,), generator=g).float().to(dev)
    ys = torch.randint(0, H, (N,), generator=g).float().to(dev)
    ts = torch.sort(torch.rand(N, generator=g))[0].to(dev)
    ps = (torch.randint(0, 2, (N,), generator=g) * 2 - 1).float().to(dev)
    L.tef_prof_name.restype = ctypes.c_char_p
    kid = [k for k in range(L.tef_prof_num_kernels()) if L.tef_prof_name(k) == b"encoding_kernels"][0]
    out = {}
    for name, call, nbytes in (("to_voxel_kernel", lambda: enc.events_to_voxel(xs, ys, ts, ps, bins, (H, W)), 16 * N + 4 * bins * H * W),
                               ("to_channels_kernel", lambda: enc.events_to_channels(xs, ys, ps, (H, W)), 12 * N + 8 * H * W)):
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        L.tef_prof_reset()
        L.tef_prof_enable(1)
        for _ in range(reps):
            call()
        torch.cuda.synchronize()
        L.tef_prof_enable(0)
        tot, timed, cnt = ctypes.c_double(), ctypes.c_long(), ctypes.c_long()
        L.tef_prof_read(kid, ctypes.byref(tot), ctypes.byref(timed), ctypes.byref(cnt))
        ms = tot.value / max(timed.value, 1)
        out[name] = {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / hbm,
                     "kernel_ms_avg": ms, "algorithmic_bytes_per_launch": nbytes, "events": N, "resolution": [H, W]}
    return out


def cm_lane_ops(wl, kernel):
    """Lane-op counts of the two Iterative kernels as issued (mode two, S = 1), per event and flow scale: P+1 chain steps of
    two 16-byte tap-row gathers and ~0.8 P reference times of 2 red.v4 (forward); ~0.8 P nodes of 2 + 2 16-byte gathers
    (gradient-image corner pairs, tap rows) and 2 red.v4 (backward)."""
    E = events_per_step(wl) * wl["F"]
    Eg = wl["B"] * wl["P"] * wl["N"] * wl["F"]
    P = wl["P"]
    pairs = sum(min(P, tr + P // 2) - max(0, tr - P // 2) for tr in range(P + 1)) / P     # (event, tref) pairs per event: 8 at P = 10
    if kernel == "iter_fwd_kernel":
        return {"gathers": 2 * (P + 1) * E, "reds": 2 * pairs * E}
    if kernel == "iter_bwd_kernel":
        return {"gathers": 4 * pairs * Eg, "reds": 2 * pairs * Eg}
    return None


def run_ours(args, wl):
    import ctypes

    import torch.distributed as dist

    from taming_event_flow_b200 import _lib, synthetic as syn
    from taming_event_flow_b200.loss import flow as tef_flow

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_cpus(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    L.tef_launch_count.restype = ctypes.c_long

    seq = fast_sequence(100 + rank, wl)
    P, F = wl["P"], wl["F"]
    cfg = syn.loss_config(wl["H"], wl["W"], wl["B"], P, wl["S"], wl["mode"], warping=wl["warping"])
    module = getattr(tef_flow, wl["warping"])(cfg, dev)
    E = events_per_step(wl)
    if wl.get("ragged"):                               # padding rows are not events
        E = int(sum((m.sum(-1) > 0).sum().item() for key in ("masks", "d_masks") for m in seq[key]))
    nsteps = args.warmup + args.steps

    # ---- device-resident arm: fresh event tensors for every step (update() mutates ts in place)
    d_flows = [[f.to(dev).requires_grad_(True) for f in per] for per in seq["flows"]]
    d_masks = [m.to(dev) for m in seq["masks"]]
    d_dmasks = [m.to(dev) for m in seq["d_masks"]]
    ev_src = [e.to(dev) for e in seq["events"]]
    dev_src = [e.to(dev) for e in seq["d_events"]]
    ev_steps = [[e.clone() for e in ev_src] for _ in range(nsteps)]
    dev_steps = [[e.clone() for e in dev_src] for _ in range(nsteps)]

    def step_resident(i):
        module.reset()
        for t in range(P):
            module.update(d_flows[t], ev_steps[i][t], d_masks[t], dev_steps[i][t], d_dmasks[t])
        loss = module()
        loss.backward()
        for per in d_flows:
            for f in per:
                f.grad = None
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local) if rank == 0 else None      # NVML initialised before the warm-up
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    L.tef_prof_enable(0)
    if rank == 0:
        clocks.start()
    launches0 = L.tef_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        loss = step_resident(args.warmup + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.tef_launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    loss_value = float(loss.item())

    # Second timed pass over the same workload with the library's per-launch CUDA events switched on (they cost ~6 % of the
    # step: an event pair around each of the ~50 launches), for the per-kernel durations of the roofline.
    prof_steps = min(args.steps, 10)
    del ev_steps[args.warmup + prof_steps:], dev_steps[args.warmup + prof_steps:]
    for i in range(args.warmup, args.warmup + prof_steps):
        for t in range(P):
            ev_steps[i][t].copy_(ev_src[t])
            dev_steps[i][t].copy_(dev_src[t])
    L.tef_prof_reset()
    L.tef_prof_enable(1)
    barrier()
    e0.record()
    for i in range(prof_steps):
        step_resident(args.warmup + i)
    e1.record()
    barrier()
    ms_prof = e0.elapsed_time(e1) / prof_steps
    L.tef_prof_enable(0)

    # per-kernel device time (CUDA events on the launching stream, recorded inside the library)
    kern = {}
    L.tef_prof_name.restype = ctypes.c_char_p
    for k in range(L.tef_prof_num_kernels()):
        tot, timed, cnt = ctypes.c_double(), ctypes.c_long(), ctypes.c_long()
        L.tef_prof_read(k, ctypes.byref(tot), ctypes.byref(timed), ctypes.byref(cnt))
        if timed.value:
            kern[L.tef_prof_name(k).decode()] = {"ms_total": tot.value, "launches": timed.value, "ms_avg": tot.value / timed.value}
    del ev_steps, dev_steps
    torch.cuda.empty_cache()

    # ---- end-to-end arms: every input in pinned host memory, loss + flow gradients read back into pinned memory.
    # Pipelined like a prefetching loader: the H2D copies of step i+1 run on a copy stream while step i computes and the
    # D2H read-back of step i runs on a third stream; every copy of every step is inside the timed region.
    from taming_event_flow_b200.dataloader import base as tef_base

    B, H, W = wl["B"], wl["H"], wl["W"]
    h_flow_all = torch.stack([torch.stack(per) for per in seq["flows"]]).pin_memory()               # [P,F,B,2,H,W]
    h_ev_all, h_mk_all = torch.stack(seq["events"]).pin_memory(), torch.stack(seq["masks"]).pin_memory()
    h_dev_all, h_dmk_all = torch.stack(seq["d_events"]).pin_memory(), torch.stack(seq["d_masks"]).pin_memory()
    h_grads = torch.empty((P, F, B, 2, H, W), dtype=torch.float32).pin_memory()
    h_loss = torch.empty((), dtype=torch.float32).pin_memory()
    d_grads, d_loss = torch.empty(h_grads.shape, dtype=torch.float32, device=dev), torch.empty((), dtype=torch.float32, device=dev)
    d2h = h_grads.numel() * 4 + 4
    copy_stream, d2h_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    e2e_steps = max(3, min(args.steps, 40))          # the same K steps as the resident loop (the first one pays its un-overlapped upload)

    def run_pipeline(host_tensors, windows_of, full):
        """host_tensors: pinned inputs of one step; windows_of(slot, t) -> the four event tensors of pass t.
        full = True: the flow maps are inputs too (host_tensors[0], uploaded every step) and every flow gradient is read back;
        full = False: the step a training loop runs -- only the EVENTS cross PCIe, the flow maps are device tensors (a network's
        output), their gradients stay on the device (the network's backward consumes them) and the loss is read back.
        Two device slots are allocated once (a prefetching loader's staging buffers): no allocator traffic when timed."""
        slots = [tuple(torch.empty_like(h, device=dev) for h in host_tensors) for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        read_back = torch.cuda.Event()
        copy_marks = []

        def prefetch(k):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[k])      # the step that used this slot has finished with it
                m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                m0.record(copy_stream)
                for d, h in zip(slots[k], host_tensors):
                    d.copy_(h, non_blocking=True)
                m1.record(copy_stream)
                copy_marks.append((m0, m1))
                ready[k].record(copy_stream)

        def step(i, last):
            k = i % 2
            if not last:
                prefetch((i + 1) % 2)
            main_stream.wait_event(ready[k])
            module.reset()
            flows = []
            for t in range(P):
                fl = [slots[k][0][t, f].requires_grad_(True) for f in range(F)] if full else d_flows[t]
                flows.append(fl)
                module.update(fl, *windows_of(slots[k], t))
            consumed[k].record(main_stream)              # update() has staged the events and packed the flow maps
            loss = module()
            loss.backward()
            main_stream.wait_event(read_back)            # the previous step's read-back has left d_grads / d_loss
            if full:
                torch.stack([f.grad for per in flows for f in per], out=d_grads.view(P * F, B, 2, H, W))
            else:
                for per in flows:
                    for f in per:
                        f.grad = None
            d_loss.copy_(loss.detach())
            done = torch.cuda.Event()
            done.record(main_stream)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                if full:
                    h_grads.copy_(d_grads, non_blocking=True)
                h_loss.copy_(d_loss, non_blocking=True)
                read_back.record(d2h_stream)

        for k in range(2):
            consumed[k].record(main_stream)
        read_back.record(d2h_stream)
        prefetch(0)
        for i in range(2):
            step(i, False)
        barrier()
        loss_seen = float(h_loss.item())
        del copy_marks[:]
        # restart the pipeline so that the first timed step pays its own (un-overlapped) upload
        e0.record()
        prefetch(0)
        for i in range(e2e_steps):
            step(i, i == e2e_steps - 1)
        main_stream.wait_stream(d2h_stream)              # the last read-back is inside the timed region
        e1.record()
        barrier()
        copy_ms = sum(a.elapsed_time(b) for a, b in copy_marks) / max(len(copy_marks), 1)
        return e0.elapsed_time(e1), loss_seen, copy_ms

    # (1) everything from the host, in the reference's own tensors: fp32 event lists + polarity masks (24 B per event) and the
    # flow maps up, every flow gradient down
    lists = (h_flow_all, h_ev_all, h_mk_all, h_dev_all, h_dmk_all)
    h2d_full = sum(x.numel() * 4 for x in lists)
    ms_full, loss_full, cp_full = run_pipeline(lists, lambda sl, t: (sl[1][t], sl[2][t], sl[3][t], sl[4][t]), True)
    torch.cuda.empty_cache()

    # the packed loader contract: events cross PCIe once, 8 B each, and are formatted (and split into
    # gradient / detached lists) on the device by dataloader/base.py format_windows
    batches = []
    for t in range(P):
        wins = []
        for b in range(B):
            ev = torch.cat([seq["events"][t][b], seq["d_events"][t][b]]).numpy()
            wins.append(tef_base.pack_events(ev[:, 2].astype(np.int64), ev[:, 1].astype(np.int64), ev[:, 0], (ev[:, 3] > 0).astype(np.int64)))
        batches.append(tef_base.PackedBatch(wins))
    k_grad = wl["N"] if wl["Nd"] > 0 else None
    packed_bytes = sum(bt.nbytes for bt in batches)

    def packed_windows(off):
        def fn(sl, t):
            w = tef_base.format_windows(batches[t], (H, W), dev, max_num_grad_events=k_grad, with_cnt=False, uploaded=(sl[off + 2 * t], sl[off + 1 + 2 * t]))
            return w["event_list"], w["event_list_pol_mask"], w["d_event_list"], w["d_event_list_pol_mask"]
        return fn

    # (2) packed events + flow maps up, gradients down
    packed_host = [h_flow_all]
    for bt in batches:
        packed_host.extend([bt.host, bt.offsets_host])
    ms_packed_full, loss_packed_full, cp_packed_full = run_pipeline(tuple(packed_host), packed_windows(1), True)
    torch.cuda.empty_cache()

    # (3) THE HEADLINE e2e: what crosses PCIe in a training step -- the packed events up, the loss down; flow maps and their
    # gradients are device tensors on both sides of the loss (network output / network backward)
    ms_e2e, loss_e2e, cp_e2e = run_pipeline(tuple(packed_host[1:]), packed_windows(0), False)
    h2d, d2h_e2e = packed_bytes, 4

    # ---- aggregate over ranks (max time), whole-job throughput
    # ---- aggregate over ranks (max time), whole-job throughput
    times = torch.tensor([ms, ms_e2e, ms_packed_full, ms_full, cp_e2e, cp_packed_full, cp_full], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_packed_full, ms_full, cp_e2e, cp_packed_full, cp_full = times.tolist()
    value = world * E * args.steps / (ms * 1e-3) / 1e6
    e2e_value = world * E * e2e_steps / (ms_e2e * 1e-3) / 1e6

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak_src = "fallback (B200_PROFILING.md)"
        hbm = 6650.0
        if os.path.exists(pk):
            peaks = json.load(open(pk))
            hbm = float(peaks.get("hbm_gbs", hbm))
            peak_src = "measured (MEASURED_PEAKS.json)"
        cand = [k for k in ("iter_fwd_kernel", "iter_bwd_kernel", "linear_fwd_kernel", "linear_bwd_kernel") if k in kern]
        dom = max(cand, key=lambda k: kern[k]["ms_total"])
        nbytes = kernel_bytes(wl, dom)
        achieved = nbytes / (kern[dom]["ms_avg"] * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(wl["name"], {}).get(dom)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                    "traffic": traffic, "algorithmic_bytes_per_launch": nbytes, "kernel_ms_avg": kern[dom]["ms_avg"], "peak_source": peak_src,
                    "kernel_share_of_step": kern[dom]["ms_avg"] * kern[dom]["launches"] / prof_steps / ms_prof}
        # the bound that actually bites: L2 gather / reduction lane-op rates, measured on this GPU
        rates = measure_l2_rates(L, dev)
        l2 = {"peaks_Gops": {k: round(v, 1) for k, v in rates.items()}, "kernels": {}}
        for k in ("iter_fwd_kernel", "iter_bwd_kernel"):
            ops = cm_lane_ops(wl, k) if k in kern else None
            if ops:
                t_g = ops["gathers"] / (rates["gather16_local"] * 1e9)
                t_r = ops["reds"] / (rates["red_v4_local"] * 1e9)
                t_min = max(t_g, t_r)                   # roofline: the slower of the two resources, each at its measured peak
                l2["kernels"][k] = {"gathers16": int(ops["gathers"]), "red_v4": int(ops["reds"]), "ms_gathers_at_peak": round(t_g * 1e3, 4),
                                    "ms_reds_at_peak": round(t_r * 1e3, 4), "bound": "gather" if t_g >= t_r else "red",
                                    "frac": round(t_min * 1e3 / kern[k]["ms_avg"], 4)}
        # The unit both event kernels run closest to: the SM's L1TEX path moves about one scattered
        # 16-byte lane access per cycle -- every lane of a red.v4, every distinct sector of a gather or store.  Sector counts per
        # launch come from the committed ncu capture of this workload (profiles/traffic.json), the time is measured live, the
        # peak is the red.v4 lane rate of the tile-sorted micro-benchmark on this GPU.
        l1 = None
        tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
        sect = tj.get(wl["name"], {}).get("l1tex_lane_sectors")
        if sect:
            # peak: one sector per SM and cycle through the L1TEX t-stage (what ncu's l1tex__throughput is a percentage of), at the SM
            # clock sampled under load; the tile-sorted red.v4 micro-benchmark (all lanes reductions, three lanes per pixel) is listed beside it
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            mhz = float((clk or {}).get("sm_mhz") or 0.0) or 1965.0
            peak_l1 = sms * mhz * 1e-3
            l1 = {"peak_G_per_s": round(peak_l1, 1), "peak_source": "%d SMs x %.0f MHz x 1 sector per cycle" % (sms, mhz), "unit": "G lane-sectors/s",
                  "red_v4_tile_sorted_microbench_G_per_s": round(rates["red_v4_tile_sorted"], 1), "source": tj.get("_l1tex_source"), "kernels": {}}
            for k, n in sect.items():
                if k in kern:
                    a = n / (kern[k]["ms_avg"] * 1e-3) / 1e9
                    l1["kernels"][k] = {"lane_sectors_per_launch": int(n), "achieved": round(a, 1), "frac": round(a / peak_l1, 4)}
        try:
            enc_roof = encoding_roofline(L, dev, hbm)
        except Exception as exc:
            enc_roof = {"error": repr(exc)[:200]}
        cpu = run_cpu_baseline(wl)
    # the second half of BASELINE.json's metric ("train windows/s"): a short run of the training-step workload
    # (PyTorch network + CM loss + SUM all-reduce), reported as an extra key of the same line
    train = None
    if os.environ.get("TEF_BENCH_TRAIN", "1") != "0":
        try:
            targs = argparse.Namespace(steps=min(args.steps, 5), warmup=3)
            train = run_train(targs, dict(TRAIN_WORKLOADS["train_128x128_b8"], name="train_128x128_b8"), quiet=True)
            train = {k: train[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "scaling", "dtype", "config")}
            targs.train_dtype = "bf16"            # same step with the network under bf16 autocast (fp32 flow heads, fp32 CM loss)
            t16 = run_train(targs, dict(TRAIN_WORKLOADS["train_128x128_b8"], name="train_128x128_b8"), quiet=True)
            train["bf16_network"] = {k: t16[k] for k in ("value", "unit", "ms_per_step", "dtype")}
            targs.fp32_heads = False
            t16 = run_train(targs, dict(TRAIN_WORKLOADS["train_128x128_b8"], name="train_128x128_b8"), quiet=True)
            train["bf16_network_bf16_heads"] = {k: t16[k] for k in ("value", "unit", "ms_per_step", "dtype")}
        except Exception as exc:      # the extra must never take the headline number down
            train = {"error": repr(exc)[:200]}
    if rank == 0:
        line = {
            "metric": "cm_loss_fwd_bwd_throughput", "value": value, "unit": "Mevents/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": with_affinity(workload_config(wl), affinity), "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "Mevents/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_e2e, "steps": e2e_steps,
                    "ms_per_step": ms_e2e / e2e_steps, "loss": loss_e2e, "h2d_gbps_per_gpu_while_copying": round(h2d / (cp_e2e * 1e-3) / 1e9, 2),
                    "note": "what crosses PCIe in a training step: events uploaded from pinned host memory as 8-byte packed records (dataloader/base.py "
                            "pack_events / format_windows) and the loss read back; flow maps and flow gradients are device tensors (network output / network "
                            "backward).  e2e_full_* upload the flow maps and read every gradient back as well"},
            "e2e_full_packed": {"value": world * E * e2e_steps / (ms_packed_full * 1e-3) / 1e6, "unit": "Mevents/s", "h2d_bytes_per_step": h_flow_all.numel() * 4 + packed_bytes,
                                "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": ms_packed_full / e2e_steps, "loss": loss_packed_full,
                                "h2d_gbps_per_gpu_while_copying": round((h_flow_all.numel() * 4 + packed_bytes) / (cp_packed_full * 1e-3) / 1e9, 2)},
            "e2e_full_fp32_lists": {"value": world * E * e2e_steps / (ms_full * 1e-3) / 1e6, "unit": "Mevents/s", "h2d_bytes_per_step": h2d_full, "d2h_bytes_per_step": d2h,
                                    "steps": e2e_steps, "ms_per_step": ms_full / e2e_steps, "loss": loss_full,
                                    "h2d_gbps_per_gpu_while_copying": round(h2d_full / (cp_full * 1e-3) / 1e9, 2),
                                    "note": "round 1's e2e: the reference's own fp32 tensors (24 B per event) and the flow maps up, all gradients down"},
            "gpu_launches": int(launches), "roofline": roofline, "roofline_l2_ops": l2, "roofline_l1tex": l1, "roofline_encodings": enc_roof, "cpu_baseline": cpu,
            "kernels": {k: {"ms_avg": round(v["ms_avg"], 5), "launches": v["launches"], "share_of_step": round(v["ms_total"] / prof_steps / ms_prof, 4)}
                        for k, v in kern.items()},
            "kernels_note": "per-kernel CUDA events in a second pass of %d steps at %.4f ms/step (the events themselves cost the difference to ms_per_step)" % (prof_steps, ms_prof),
            "loss": loss_value, "events_per_step_per_gpu": E, "train_step": train,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train(args, wl, quiet=False):
    """Training step of the recurrent EV-FlowNet (PyTorch/cuDNN) with the CM loss: P x (encode -> network -> update),
    loss, backward through the loss kernels and the network, SUM all-reduce of the gradients, clip, Adam.
    windows/s = global batch x P / step time."""
    import torch.distributed as dist

    from taming_event_flow_b200 import synthetic as syn
    from taming_event_flow_b200.flownet import RecEVFlowNet
    from taming_event_flow_b200.loss import flow as tef_flow
    from taming_event_flow_b200.training import GradReducer, GraphedTrainStep, shard_range, train_step

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    own_pg = False
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
        own_pg = True
    if wl["scaling"] == "strong":
        a, b = shard_range(wl["B"], world, rank)
        B_local, B_global = b - a, wl["B"]
    else:
        B_local, B_global = wl["B"], wl["B"] * world
    wl_local = dict(wl, B=B_local)
    seq = fast_sequence(500 + rank, wl_local)
    P = wl["P"]
    cfg = syn.loss_config(wl["H"], wl["W"], B_local, P, wl["S"], wl["mode"], warping=wl["warping"])
    loss_fn = getattr(tef_flow, wl["warping"])(cfg, dev)
    torch.manual_seed(0)
    # The network's convolutions stay on PyTorch/cuDNN (north_star).  What is applied around them: channels_last,
    # the element-wise stages as fused CUDA kernels with one weight gradient per layer and loss window (netops, csrc/tef_net.cu),
    # one CUDA graph over forward + loss + backward, a flat gradient buffer with a bucketed SUM all-reduce issued under the
    # backward pass, fused Adam; optional bf16 autocast (plain modules; the CM loss stays fp32).
    mode = getattr(args, "train_mode", None) or os.environ.get("TEF_TRAIN_MODE", TRAIN_MODE_DEFAULT)
    dtype = getattr(args, "train_dtype", None) or os.environ.get("TEF_TRAIN_DTYPE", TRAIN_DTYPE_DEFAULT)
    autocast = torch.bfloat16 if dtype == "bf16" else None
    torch.backends.cudnn.benchmark = True
    fp32_heads = os.environ.get("TEF_TRAIN_FP32_HEADS", "1") != "0" and getattr(args, "fp32_heads", True)
    fused = os.environ.get("TEF_TRAIN_FUSED", "1") != "0"          # 0: the plain PyTorch modules (A/B of the fused element-wise kernels)
    model = RecEVFlowNet(num_bins=2, fp32_heads=fp32_heads, fused=fused).to(dev).to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(model.parameters(), lr=1e-5, fused=True)      # the optimizer step stays outside the captured graph
    reducer = GradReducer(list(model.parameters()), world)
    from taming_event_flow_b200.dataloader.encodings import events_to_channels_batched

    def encode(ev, dv):
        x = events_to_channels_batched(torch.cat([ev, dv], 1) if dv.shape[1] else ev, (wl["H"], wl["W"]))
        return x.contiguous(memory_format=torch.channels_last)

    nsteps = args.warmup + args.steps
    masks = [(seq["masks"][t].to(dev), seq["d_masks"][t].to(dev)) for t in range(P)]
    src = [(seq["events"][t].to(dev), seq["d_events"][t].to(dev)) for t in range(P)]

    if mode == "graph":
        # static inputs of the captured step; a step's events are copied in before the replay (inside the timed region)
        static = [(src[t][0].clone(), masks[t][0], src[t][1].clone(), masks[t][1]) for t in range(P)]
        graphed = GraphedTrainStep(model, loss_fn, opt, static, flow_scaling=32.0, clip_grad=100.0, encode=encode, reducer=reducer, autocast=autocast,
                                   capture_collectives=False)       # capturing the all-reduces in the graph hung at 8 GPUs with the final build: not offered here

        def step(i):
            for t in range(P):
                static[t][0].copy_(src[t][0])
                static[t][2].copy_(src[t][1])
            return graphed.step()
    else:
        evs = [[(src[t][0].clone(), src[t][1].clone()) for t in range(P)] for _ in range(nsteps)]

        def step(i):
            windows = [(evs[i][t][0], masks[t][0], evs[i][t][1], masks[t][1]) for t in range(P)]
            return train_step(model, loss_fn, opt, windows, flow_scaling=32.0, clip_grad=100.0, world_size=world, encode=encode, reducer=reducer,
                              autocast=autocast)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        loss = step(args.warmup + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    res = {"metric": "train_throughput", "value": B_global * P * args.steps / (ms * 1e-3), "unit": "windows/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": wl["scaling"],
           "vs_baseline": None, "dtype": dtype, "data": "synthetic",
           "config": dict(workload_config(wl), batch_global=B_global, batch_per_gpu=B_local, network="RecEVFlowNet (cuDNN convolutions, channels_last, %s, 31.4M params)" % (("bf16 autocast on the plain PyTorch modules (the fused element-wise kernels are fp32), %s flow heads, fp32 CM loss" % ("fp32" if fp32_heads else "bf16")) if autocast else "fp32/TF32, plain PyTorch modules" if not fused else "fp32/TF32; ConvGRU gates, bias + activation, decoder inputs and flow-head up-sampling as fused CUDA kernels (netops), one weight gradient per layer and loss window"),
                          step_mode=(("one CUDA graph over forward + CM loss + backward%s; clip, Adam eager" % (" + the bucketed all-reduces" if graphed.comm_captured else "; all-reduce after the replay"))
                                     if mode == "graph" else "eager"),
                          optimizer="fused Adam lr 1e-5, clip 100, flat gradient buffer, bucketed SUM all-reduce (%d buckets; issued from gradient hooks in eager mode, after the replay in graph mode)" % len(reducer.buckets)),
           "loss": float(loss.item()), "events_per_step": B_global * P * (wl["N"] + wl["Nd"])}
    if own_pg:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and not quiet:
        print(json.dumps(res))
    return res


def run_inference(args, wl):
    """Sequential inference (eval_flow.py:70-90): per window, events_to_voxel (our kernel) + recurrent network forward
    (PyTorch), batch 1, no gradients.  Reports windows/s, the per-window latency split and the encoding kernel's HBM roofline."""
    import ctypes

    from taming_event_flow_b200 import _lib
    from taming_event_flow_b200.dataloader.encodings import events_to_voxel
    from taming_event_flow_b200.flownet import RecEVFlowNet

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    H, W, N, bins = wl["H"], wl["W"], wl["N"], wl["bins"]
    g = torch.Generator().manual_seed(3)
    nwin = 16
    wins = []
    for _ in range(nwin):
        ts, _ = torch.sort(torch.rand(N, generator=g))
        ts = (ts - ts[0]) / (ts[-1] - ts[0])
        wins.append(tuple(t.to(dev) for t in (torch.randint(0, W, (N,), generator=g).float(), torch.randint(0, H, (N,), generator=g).float(), ts,
                                              (torch.randint(0, 2, (N,), generator=g) * 2 - 1).float())))
    torch.manual_seed(0)
    model = RecEVFlowNet(num_bins=bins).to(dev).eval()
    L = _lib.lib()

    def window(i, net=True):
        xs, ys, ts, ps = wins[i % nwin]
        vox = events_to_voxel(xs, ys, ts, ps, bins, (H, W))
        return model(vox.unsqueeze(0))["flow"][-1] if net else vox

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.no_grad():
        for i in range(args.warmup):
            window(i)
        torch.cuda.synchronize()
        ev[0].record()
        for i in range(args.steps):
            window(i, net=False)
        ev[1].record()
        torch.cuda.synchronize()
        L.tef_prof_reset(); L.tef_prof_enable(1)
        ev[2].record()
        for i in range(args.steps):
            window(i)
        ev[3].record()
        torch.cuda.synchronize()
        L.tef_prof_enable(0)
    tot, timed, cnt = ctypes.c_double(), ctypes.c_long(), ctypes.c_long()
    L.tef_prof_name.restype = ctypes.c_char_p
    kid = [k for k in range(L.tef_prof_num_kernels()) if L.tef_prof_name(k) == b"encoding_kernels"][0]
    L.tef_prof_read(kid, ctypes.byref(tot), ctypes.byref(timed), ctypes.byref(cnt))
    enc_ms = ev[0].elapsed_time(ev[1]) / args.steps
    tot_ms = ev[2].elapsed_time(ev[3]) / args.steps
    k_ms = tot.value / max(timed.value, 1)
    nbytes = 16 * N + 4 * bins * H * W
    hbm = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        hbm = float(json.load(open(pk)).get("hbm_gbs", hbm))
    print(json.dumps({
        "metric": "inference_throughput", "value": 1e3 / tot_ms, "unit": "windows/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tot_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "events_per_window": N, "resolution": [H, W], "num_bins": bins, "network": "RecEVFlowNet (PyTorch fp32 eager)"},
        "latency_ms": {"encode_call": enc_ms, "encode_kernel": k_ms, "network_forward": tot_ms - enc_ms, "window": tot_ms},
        "encode_Mevents_per_s": N / (enc_ms * 1e-3) / 1e6,
        "roofline": {"bound": "hbm", "kernel": "to_voxel_kernel", "achieved": nbytes / (k_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": nbytes / (k_ms * 1e-3) / 1e9 / hbm, "traffic": None, "algorithmic_bytes_per_launch": nbytes}}))


def run_validation(args, wl):
    """eval_flow.py:114-176 without the network: per window `criteria.update(...)`, and FWL + RSAT every P windows."""
    from taming_event_flow_b200 import synthetic as syn
    from taming_event_flow_b200.loss import flow_val as fv

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    H, W, N, P = wl["H"], wl["W"], wl["N"], wl["P"]
    g = torch.Generator().manual_seed(9)
    wins = []
    for t in range(P):
        ev, mk = syn.make_window(g, 1, N, H, W)
        wins.append((ev.to(dev), mk.to(dev), syn.make_flow(g, 1, H, W, 3.0).to(dev), torch.ones(1, 1, H, W, device=dev)))
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"round_ts": False}, "vis": {"mask_output": True}, "metrics": {"name": ["FWL", "RSAT"]}}
    crit = fv.Iterative(cfg, dev)

    def interval():
        crit.reset()
        for ev, mk, flow, em in wins:
            crit.update([flow], ev.clone(), mk, em)
        return crit.fwl(), crit.rsat()

    with torch.no_grad():
        for _ in range(max(2, args.warmup // 2)):
            interval()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            fwl, rsat = interval()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({
        "metric": "validation_throughput", "value": P / (ms * 1e-3), "unit": "windows/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "events_per_window": N, "resolution": [H, W], "windows_per_interval": P,
                   "criteria": "flow_val.Iterative update x P + FWL + RSAT"},
        "fwl": float(fwl.item()), "rsat": float(rsat.item()), "Mevents_per_s": P * N / (ms * 1e-3) / 1e6}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--train-mode", default=None, choices=["eager", "graph"], help="training-step workloads: eager launches or one CUDA graph")
    ap.add_argument("--train-dtype", default=None, choices=["f32", "bf16"], help="training-step workloads: network compute type (the CM loss is fp32)")
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD,
                    choices=sorted(WORKLOADS) + sorted(TRAIN_WORKLOADS) + sorted(INFER_WORKLOADS) + sorted(VAL_WORKLOADS))
    args = ap.parse_args()
    if args.workload in VAL_WORKLOADS:
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the CPU arm covers the CM-loss workloads only"}))
            return
        run_validation(args, dict(VAL_WORKLOADS[args.workload], name=args.workload))
        return
    if args.workload in INFER_WORKLOADS:
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the CPU arm covers the CM-loss workloads only"}))
            return
        args.warmup = max(args.warmup, 3)
        run_inference(args, dict(INFER_WORKLOADS[args.workload], name=args.workload))
        return
    if args.workload in TRAIN_WORKLOADS:
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the CPU arm covers the CM-loss workloads only"}))
            return
        args.warmup = max(args.warmup, 3)
        run_train(args, dict(TRAIN_WORKLOADS[args.workload], name=args.workload))
        return
    wl = dict(WORKLOADS[args.workload], name=args.workload)
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args, wl)


if __name__ == "__main__":
    main()
