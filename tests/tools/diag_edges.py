import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, copy
from oracle import cm_oracle as orc
from taming_event_flow_b200 import synthetic as syn
from taming_event_flow_b200.loss import flow as tef_flow
from util import rel_err
B,P,N,Nd,H,W,F,S,mode,sigma,ragged,border,dist = 1,10,100000,20000,480,640,1,1,"two",1.0,True,True,"edges"
seq = syn.make_sequence(11, B, P, N, Nd, H, W, F, sigma, ragged, dist)
oc = orc.make_cfg(B, H, W, P, F, S, mode, border)
o32 = orc.iterative(oc, seq["flows"], seq["events"], seq["masks"], seq["d_events"], seq["d_masks"], np.float32, want_grad=True, want_iwe=True)
o64 = orc.iterative(oc, seq["flows"], seq["events"], seq["masks"], seq["d_events"], seq["d_masks"], np.float64, want_grad=True, want_iwe=True)
res = {}
for det in (False, True):
    cfg = syn.loss_config(H, W, B, P, S, mode); cfg["loss"]["deterministic"] = det
    m = tef_flow.Iterative(cfg, torch.device("cuda"))
    fl = [[f.cuda().requires_grad_(True) for f in per] for per in seq["flows"]]
    for t in range(P):
        m.update(fl[t], seq["events"][t].cuda().clone(), seq["masks"][t].cuda(), seq["d_events"][t].cuda().clone(), seq["d_masks"][t].cuda())
    loss = m(); iwe = m.images().cpu().numpy(); loss.backward()
    g = np.stack([np.stack([fl[t][f].grad.cpu().numpy() for t in range(P)]) for f in range(F)])
    res[det] = (loss.item(), iwe, g)
print("max events per pixel (count image max):", o64["iwe"][..., :2, :, :].max())
for name,(l,iwe,g) in (("gpu", res[False]), ("gpu_det", res[True])):
    print(name, "loss rel vs o32 %.2e vs o64 %.2e" % (abs(l-o32["loss"])/abs(o32["loss"]), abs(l-o64["loss"])/abs(o64["loss"])))
    print("   iwe  vs o32", rel_err(iwe, o32["iwe"]), "vs o64", rel_err(iwe, o64["iwe"]))
    print("   grad vs o32", rel_err(g, o32["gflow"]), "vs o64", rel_err(g, o64["gflow"]))
print("o32 vs o64: loss %.2e" % (abs(o32["loss"]-o64["loss"])/abs(o64["loss"])), "iwe", rel_err(o32["iwe"], o64["iwe"]), "grad", rel_err(o32["gflow"], o64["gflow"]))
