"""The network restatement (taming_event_flow_b200/flownet.py, plain PyTorch formulation) against upstream's RecEVFlowNet
(models/model.py:6-85, models/arch.py:197-242, models/submodules.py): golden vectors generated from the unmodified reference
(tests/golden/make_golden_flownet.py) and, in the build container, the reference imported live.  CPU.  The fused CUDA operators are
checked against this plain formulation on the GPU (tests/test_netops_gpu.py), which closes the chain to the reference."""
import os
import sys

import numpy as np
import pytest
import torch

from taming_event_flow_b200.flownet import RecEVFlowNet, count_parameters, from_upstream_state_dict, to_upstream_state_dict
from util import GOLDEN, rel_err

REF = os.environ.get("TEF_REFERENCE", "/root/reference")


def test_plain_network_reproduces_the_reference_golden_vectors():
    z = np.load(os.path.join(GOLDEN, "flownet.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    net = RecEVFlowNet(num_bins=2, base_channels=4)
    net.load_state_dict(from_upstream_state_dict(sd))
    xs, gs = torch.from_numpy(z["x"]), torch.from_numpy(z["g"])
    loss = 0.0
    for t in range(xs.shape[0]):
        flows = net(xs[t])["flow"]                      # 40 x 56 input: padded to 48 x 64 and cropped like upstream's ImagePadder
        assert len(flows) == 4
        for i, f in enumerate(flows):
            ref = z["flow/%d/%d" % (t, i)]
            assert tuple(f.shape) == ref.shape
            assert rel_err(f.detach().numpy(), ref)[0] < 1e-6, (t, i)
            loss = loss + (f * gs[t, i]).sum()
    loss.backward()
    grads = to_upstream_state_dict({k: p.grad for k, p in net.named_parameters()})
    for k in z.files:
        if k.startswith("grad/"):
            assert rel_err(grads[k[5:]].numpy(), z[k])[0] < 1e-5, k


def test_state_dict_mapping_round_trips_and_keeps_the_parameter_count():
    net = RecEVFlowNet(num_bins=2, base_channels=8)
    sd = net.state_dict()
    up = to_upstream_state_dict(sd)
    assert sum(v.numel() for v in up.values()) == count_parameters(net)
    assert any(k.endswith("recurrent_block.update_gate.weight") for k in up) and any(k.startswith("arch.preds.3.conv2d") for k in up)
    back = from_upstream_state_dict(up)
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    assert count_parameters(RecEVFlowNet(2)) == 31365352            # upstream RecEVFlowNet with 2 input channels


def test_flow_scaling_is_folded_into_the_flow_maps():
    torch.manual_seed(0)
    net = RecEVFlowNet(num_bins=2, base_channels=4)
    x = torch.rand(1, 2, 32, 32)
    a = net(x)["flow"]
    net.reset_states()
    b = net(x, flow_scaling=32.0)["flow"]
    for fa, fb in zip(a, b):
        assert torch.allclose(fa * 32.0, fb, rtol=1e-6, atol=0)


def test_deferred_weight_gradient_window_is_inert_on_the_plain_path():
    """begin_window only concerns the fused CUDA operators: on the CPU the modules run their plain formulation and gradients are
    the per-pass ones."""
    torch.manual_seed(1)
    net = RecEVFlowNet(num_bins=2, base_channels=4)
    xs = [torch.rand(1, 2, 32, 32) for _ in range(2)]
    out = []
    for window in (0, 2):
        net.reset_states()
        net.zero_grad(set_to_none=True)
        net.begin_window(window)
        sum(f.sum() for x in xs for f in net(x)["flow"]).backward()
        out.append([p.grad.clone() for p in net.parameters()])
    assert all(torch.equal(a, b) for a, b in zip(*out))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference not mounted")
@pytest.mark.parametrize("shape", [(2, 2, 64, 64), (1, 2, 33, 72)])
def test_plain_network_is_bit_identical_to_the_live_reference(shape):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from models.model import RecEVFlowNet as Upstream      # reference

    torch.manual_seed(3)
    up = Upstream({"base_channels": 8}, num_bins=2)
    net = RecEVFlowNet(num_bins=2, base_channels=8)
    net.load_state_dict(from_upstream_state_dict(up.state_dict()))
    for t in range(3):
        x = torch.rand(*shape) * 2.0
        fu, fo = up(x)["flow"], net(x)["flow"]
        for a, b in zip(fo, fu):
            assert torch.equal(a, b), (t, float((a - b).abs().max()))
    up.detach_states()
    net.detach_states()
    x = torch.rand(*shape)
    for a, b in zip(net(x)["flow"], up(x)["flow"]):          # states carried across a detach
        assert torch.equal(a, b)
