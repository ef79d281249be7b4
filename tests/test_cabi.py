"""The C-ABI library loads and exports every symbol include/tef_b200.h declares (no compute, runs without a GPU),
and the host mirrors expose the reference's names and signatures."""
import ctypes
import inspect
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from taming_event_flow_b200 import _lib

    _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "tef_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(tef_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    lib.tef_strerror.restype = ctypes.c_char_p
    assert lib.tef_strerror(0) == b"success"
    assert b"TypeError" in lib.tef_strerror(-4)
    assert lib.tef_version() >= 100


def test_descriptor_layout_matches_header():
    from taming_event_flow_b200 import _lib

    # 10 ints, then 2x31 pointers x2, 2x31 ints, 14 pointers, 2 ints, 2 pointers
    expect = 10 * 4 + 2 * 31 * 8 * 2 + 2 * 31 * 4 + 14 * 8 + 2 * 4 + 2 * 8
    assert ctypes.sizeof(_lib.CmDesc) == expect
    lib = ctypes.CDLL(_lib.LIB_PATH)
    d = _lib.CmDesc()
    d.B, d.H, d.W, d.P, d.F, d.S, d.mode, d.border_comp = 1, 32, 32, 10, 1, 1, 2, 1
    assert lib.tef_cm_num_slots(ctypes.byref(d), 0) == 11           # trefs 0..10
    assert lib.tef_cm_num_slots(ctypes.byref(d), 1) == 2            # Linear: both ends
    d.S = 2
    assert lib.tef_cm_num_slots(ctypes.byref(d), 0) == 11 + 2 * 6
    d.S, d.P = 1, 1
    assert lib.tef_cm_num_slots(ctypes.byref(d), 0) == -3           # delta 0 -> empty torch.cat upstream
    d.P, d.mode = 8, 4
    assert lib.tef_cm_num_slots(ctypes.byref(d), 0) == -4           # mode four + border compensation


def test_host_mirror_signatures():
    from taming_event_flow_b200.dataloader import encodings as enc
    from taming_event_flow_b200.loss import flow as lf
    from taming_event_flow_b200.utils import iwe

    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(iwe.event_propagation) == ["events_ts", "events_idx", "flow", "tref"]
    assert sig(iwe.get_event_flow) == ["flow_map_x", "flow_map_y", "event_loc"]
    assert sig(iwe.purge_unfeasible) == ["event_loc", "event_pol_mask", "res"]
    assert sig(iwe.get_interpolation) == ["warped_events", "res", "round_idx", "zeros"]
    assert sig(iwe.interpolate) == ["idx", "weights", "res", "polarity_mask", "zeros"]
    assert sig(iwe.deblur_events) == ["flow", "event_list", "res", "round_idx", "polarity_mask", "round_flow"]
    assert sig(iwe.compute_pol_iwe) == ["flow", "event_list", "res", "pol_mask", "round_idx", "round_flow"]
    assert sig(enc.events_to_image) == ["xs", "ys", "ps", "sensor_size", "accumulate"]
    assert sig(enc.events_to_voxel) == ["xs", "ys", "ts", "ps", "num_bins", "sensor_size"]
    assert sig(enc.events_to_channels) == ["xs", "ys", "ps", "sensor_size"]
    for cls in (lf.Linear, lf.Iterative):
        assert sig(cls.__init__) == ["self", "config", "device", "loss_scaling"]
        assert sig(cls.update) == ["self", "flow_list", "event_list", "pol_mask", "d_event_list", "d_pol_mask"]
        assert hasattr(cls, "reset") and hasattr(cls, "num_passes")
    assert issubclass(lf.Iterative, lf.BaseEventWarping) and lf.EventWarping is lf.Iterative


def test_product_code_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load anything under oracle/."""
    pat = re.compile(r"(^|\n)\s*(from|import)\s+oracle|libcm_oracle|#include\s+\"[^\"]*oracle|CDLL\([^)]*oracle")
    pkg = os.path.join(ROOT, "taming_event_flow_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert not pat.search(src), os.path.join(dp, f)


def test_no_contracted_packed_multiply_adds_in_the_event_kernels():
    """ptxas 12.9 contracts `mul.rn.f32x2` feeding `add.rn.f32x2` into one FFMA2 even with --fmad=false, which would
    break the bit-faithful per-event arithmetic (csrc/tef_device.cuh).  The only FFMA2 the event kernels may contain are
    the explicit fma2() calls of sample_flow_inside_xy: 2 (division by a constant) + 3 (tap accumulation, dead in the
    backward kernel, which only needs the taps) per inlined copy, two copies per kernel."""
    import shutil
    import subprocess

    import pytest

    from taming_event_flow_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    _lib.build()
    # <deterministic, quad-cell copies>: every instantiation has the same per-event arithmetic
    expect = {"iter_fwd_kernelILb0ELb0E": 10, "iter_fwd_kernelILb0ELb1E": 10, "iter_fwd_kernelILb1ELb0E": 10,
              "iter_bwd_kernelILb0ELb0E": 4, "iter_bwd_kernelILb0ELb1E": 4, "iter_bwd_kernelILb1ELb0E": 4}
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    counts, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next((k for k in expect if k in m.group(1)), None)
        elif cur and "FFMA2" in line:
            counts[cur] = counts.get(cur, 0) + 1
    assert counts == expect, counts

