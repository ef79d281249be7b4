"""Golden vectors of the recurrent flow network from the UNMODIFIED reference (build container only; needs /root/reference):

    python tests/golden/make_golden_flownet.py

Upstream's ``RecEVFlowNet`` (``models/model.py:6-85``) with ``base_channels = 4`` (the topology of the trained model at 1/16 of
its width, so that the fixture stays small), random biases, three recurrent passes on an input whose size is not a multiple
of 16 (upstream pads it) -- state dict, inputs, the four flow maps of every pass and the gradient of a fixed linear loss with
respect to every parameter go to ``tests/golden/flownet.npz``."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TEF_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from models.model import RecEVFlowNet  # noqa: E402  (reference)


def main():
    torch.manual_seed(7)
    net = RecEVFlowNet({"base_channels": 4}, num_bins=2)
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.2)
    B, H, W, T = 2, 40, 56, 3
    xs = torch.rand(T, B, 2, H, W) * 3.0
    gs = torch.randn(T, 4, B, 2, H, W)
    out = {"x": xs.numpy(), "g": gs.numpy()}
    for k, v in net.state_dict().items():
        out["sd/" + k] = v.numpy().copy()
    net.reset_states()
    loss = 0.0
    for t in range(T):
        flows = net(xs[t])["flow"]
        for i, f in enumerate(flows):
            out["flow/%d/%d" % (t, i)] = f.detach().numpy().copy()
            loss = loss + (f * gs[t, i]).sum()
    loss.backward()
    for k, p in net.named_parameters():
        out["grad/" + k] = p.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "flownet.npz"), **out)
    print("wrote flownet.npz:", sum(v.size for v in out.values()), "values")


if __name__ == "__main__":
    main()
