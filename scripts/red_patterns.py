"""red.global.add.v4.f32 lane rate on an L2-resident buffer as a function of the access pattern inside a warp
(csrc/tef_microbench.cu): which patterns the L1/L2 reduction path merges, if any."""
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
names = {0: "random over 64 MB", 1: "random in a 4 KB window per warp", 2: "32 consecutive slots per warp (512 B)", 3: "one slot for the whole warp",
         4: "lane pairs share a slot", 5: "lane pairs share a 32-byte sector", 6: "lane quads share a slot",
         7: "pairs share a slot, merged in software", 8: "random in 4 KB, merge attempted (cost)"}
for mode, name in names.items():
    ops = ctypes.c_long()
    best = 1e30
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = L.tef_microbench(0, mode, ctypes.c_void_p(buf.data_ptr()), ctypes.c_long(buf.numel()), 256, ctypes.byref(ops), st)
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0
        best = min(best, e0.elapsed_time(e1))
    print("red.v4 mode %d  %-42s %8.1f G lane-ops/s" % (mode, name, ops.value / (best * 1e-3) / 1e9))
