"""Diagnostic: eager train_step vs GraphedTrainStep, loss and state checksums per step, fused / plain network."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from taming_event_flow_b200 import synthetic as syn
from taming_event_flow_b200.flownet import RecEVFlowNet
from taming_event_flow_b200.loss.flow import Iterative
from taming_event_flow_b200.training import GradReducer, GraphedTrainStep, train_step

B, P, N, H, W = 2, 4, 1500, 64, 64
seq = syn.make_sequence(21, B, P, N, 500, H, W, 1, 1.0)
wins = [(seq["events"][t].cuda(), seq["masks"][t].cuda(), seq["d_events"][t].cuda(), seq["d_masks"][t].cuda()) for t in range(P)]


def chk(model):
    return [float(s.double().sum()) for s in model.states] + [float(sum(p.double().sum() for p in model.parameters()))]


for fused in (False, True):
    for mode in ("eager", "graph", "eager"):
        torch.manual_seed(0)
        model = RecEVFlowNet(num_bins=2, base_channels=8, fused=fused).cuda()
        opt = torch.optim.SGD(model.parameters(), lr=3e-3)
        red = GradReducer(list(model.parameters()), world_size=1)
        loss_fn = Iterative(syn.loss_config(H, W, B, P), "cuda")
        if mode == "eager":
            for i in range(5):
                l = train_step(model, loss_fn, opt, [(e.clone(), m, d.clone(), dm) for e, m, d, dm in wins], reducer=red).item()
                print(fused, mode, i, "%.7f" % l, ["%.6f" % c for c in chk(model)])
        else:
            static = [(e.clone(), m, d.clone(), dm) for e, m, d, dm in wins]
            g = GraphedTrainStep(model, loss_fn, opt, static, reducer=red, warmup=2)
            print(fused, mode, 1, "-", ["%.6f" % c for c in chk(model)])
            for i in range(2, 5):
                l = g.step().item()
                print(fused, mode, i, "%.7f" % l, ["%.6f" % c for c in chk(model)])
