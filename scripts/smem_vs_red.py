"""Shared-memory privatised accumulation vs global red.v4 (DESIGN.md decision 14).

Prints, for an L2-resident 64 MB image buffer:
  * the red.global.add.v4.f32 lane rate for the access patterns of csrc/tef_microbench.cu, including the tile-sorted pattern in the
    single-plane pair layout (mode 9) and in the dual-phase layout the CM kernels use (mode 10);
  * the rate of ORIGINAL 16-byte updates when they are first accumulated in a CTA-private shared-memory patch with shared-memory
    atomics (fp32 compare-and-swap loop / native u32 / u64 fixed point) and flushed with coalesced red.v4, for several patch sizes
    and update densities (updates per patch slot between flushes).
Both in G updates/s, so the columns compare directly."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from taming_event_flow_b200 import _lib  # noqa: E402

L = _lib.lib()
buf = torch.zeros(64 << 20, dtype=torch.uint8, device="cuda")
st = _lib.stream()


def timed(call):
    ops = ctypes.c_long()
    best = 1e30
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = call(ops)
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0, rc
        best = min(best, e0.elapsed_time(e1))
    return ops.value / (best * 1e-3) / 1e9


names = {0: "random over 64 MB", 1: "random in a 4 KB window per warp", 2: "32 consecutive slots per warp (512 B)",
         9: "tile-sorted-like, single-plane pair layout", 10: "tile-sorted-like, dual-phase layout (shipped)"}
for mode, name in names.items():
    r = timed(lambda ops: L.tef_microbench(0, mode, ctypes.c_void_p(buf.data_ptr()), ctypes.c_long(buf.numel()), 256, ctypes.byref(ops), st))
    print("global red.v4   mode %-2d %-52s %8.1f G updates/s" % (mode, name, r))

for mode, name in ((0, "2x2 fetch: two 16-byte gathers (dual-phase rows)"), (1, "2x2 fetch: one 32-byte gather (quad-phase cell)")):
    r = timed(lambda ops: L.tef_microbench(4, mode, ctypes.c_void_p(buf.data_ptr()), ctypes.c_long(buf.numel()), 256, ctypes.byref(ops), st))
    print("gather          mode %-2d %-52s %8.1f G fetches/s" % (mode, name, r))

atoms = {0: "fp32 atomicAdd (ATOMS.CAST.SPIN)", 1: "u32 fixed point (ATOMS.ADD, native)", 2: "u64 fixed point (ATOMS.CAST.SPIN.64)"}
for pattern, pname in ((0, "random in patch"), (1, "tile-sorted-like")):
    for patch in (256, 512, 2048):
        for k in (1, 2, 8, 32):
            row = []
            for atom in (0, 1, 2):
                r = timed(lambda ops: L.tef_microbench_smem(atom, patch, k, pattern, ctypes.c_void_p(buf.data_ptr()), ctypes.c_long(buf.numel()), 256,
                                                            ctypes.byref(ops), st))
                row.append(r)
            print("smem patch %4d slots (%5.1f KB fp32), %2d updates/thread/flush (%.2f per slot), %-16s  fp32-CAS %7.1f   u32 %7.1f   u64-CAS %7.1f  G updates/s"
                  % (patch, patch * 16 / 1024, k, 256 * k / patch, pname, row[0], row[1], row[2]))
