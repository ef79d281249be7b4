"""Soak: many loss windows with varying (ragged) event counts; device memory and step time must stay flat."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from taming_event_flow_b200 import synthetic as syn
from taming_event_flow_b200.loss import flow as tef_flow

B, P, H, W, F = 4, 10, 128, 128, 2
dev = torch.device("cuda")
cfg = syn.loss_config(H, W, B, P, 1, "two")
m = tef_flow.Iterative(cfg, dev)
rng = np.random.default_rng(0)
gen = torch.Generator().manual_seed(0)
flows = [[(torch.randn(B, 2, H, W, generator=gen) * 2).to(dev).requires_grad_(True) for _ in range(F)] for _ in range(P)]
base = None
t0 = time.perf_counter()
for step in range(3000):
    m.reset()
    for t in range(P):
        n, nd = int(rng.integers(100, 6000)), int(rng.integers(0, 3000))
        ev, mk = syn.make_window(gen, B, n, H, W) if step < 3 else (ev_pool[n % 8][:, :n], mk_pool[n % 8][:, :n])
        dv, dm = (ev_pool[nd % 8][:, :nd], mk_pool[nd % 8][:, :nd]) if step >= 3 else syn.make_window(gen, B, max(nd, 1), H, W)
        if step < 3:
            ev_pool = [syn.make_window(gen, B, 6000, H, W)[0].to(dev) for _ in range(8)] if step == 0 and t == 0 else ev_pool
            mk_pool = [torch.stack([(e[:, :, 3] > 0).float(), (e[:, :, 3] < 0).float()], -1) for e in ev_pool] if step == 0 and t == 0 else mk_pool
            ev, mk, dv, dm = ev.to(dev), mk.to(dev), dv.to(dev), dm.to(dev)
        m.update(flows[t], ev.clone(), mk, dv.clone(), dm)
    loss = m()
    loss.backward()
    for per in flows:
        for f in per:
            f.grad = None
    if step in (100, 1000, 2000, 2999):
        torch.cuda.synchronize()
        alloc, reserved = torch.cuda.memory_allocated() >> 20, torch.cuda.memory_reserved() >> 20
        print("step %d: %.3f ms/step so far, allocated %d MB, reserved %d MB, loss %.5f" % (step, (time.perf_counter() - t0) / (step + 1) * 1e3, alloc, reserved, loss.item()))
        if base is None:
            base = reserved
assert torch.cuda.memory_reserved() >> 20 <= base * 1.5 + 64, "device memory grows"
print("soak ok")
