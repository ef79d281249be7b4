"""Try PyTorch-side settings for the (out-of-scope) network of the training-step workload."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from taming_event_flow_b200 import synthetic as syn
from taming_event_flow_b200.flownet import RecEVFlowNet
from taming_event_flow_b200.loss import flow as tef_flow
from taming_event_flow_b200.training import train_step

wl = dict(bench.TRAIN_WORKLOADS["train_128x128_b8"], name="train_128x128_b8")
dev = torch.device("cuda", 0)
seq = bench.fast_sequence(500, wl)
P = wl["P"]
masks = [(seq["masks"][t].to(dev), seq["d_masks"][t].to(dev)) for t in range(P)]


def run(tag, cl=False, bench_flag=False, steps=5):
    torch.backends.cudnn.benchmark = bench_flag
    cfg = syn.loss_config(wl["H"], wl["W"], wl["B"], P, wl["S"], wl["mode"])
    loss_fn = tef_flow.Iterative(cfg, dev)
    torch.manual_seed(0)
    model = RecEVFlowNet(2).to(dev)
    if cl:
        model = model.to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(model.parameters(), lr=1e-5)
    enc = None
    if cl:
        from taming_event_flow_b200.dataloader.encodings import events_to_channels_batched
        enc = lambda ev, dv: events_to_channels_batched(torch.cat([ev, dv], 1), (wl["H"], wl["W"])).contiguous(memory_format=torch.channels_last)
    def step():
        windows = [(seq["events"][t].to(dev), masks[t][0], seq["d_events"][t].to(dev), masks[t][1]) for t in range(P)]
        return train_step(model, loss_fn, opt, windows, encode=enc)
    first = step().item()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        l = step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print("%-28s %.2f ms/step  %.0f windows/s  first loss %.5f  loss after 8 steps %.5f" % (tag, dt * 1e3, wl["B"] * P / dt, first, l.item()))

run("baseline")
run("cudnn.benchmark", bench_flag=True)
run("channels_last+benchmark", cl=True, bench_flag=True)
