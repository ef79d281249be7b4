"""Fused element-wise stages of the flow network's training step (netops / csrc/tef_net.cu) against their plain
PyTorch fp32 formulation -- the restatement of upstream's module code in flownet.py (models/submodules.py:111-152, models/model.py:65-85).
Tolerance 1e-5 norm-relative on values and on every gradient (fp32 convolutions, TF32 off, so both routes run the same arithmetic
up to the order of the bias-gradient sums)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


@pytest.fixture(autouse=True)
def _fp32_convs():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("deferred", [False, True])
@pytest.mark.parametrize("B,C,H,W", [(2, 8, 12, 20), (3, 64, 16, 16), (1, 512, 4, 4), (2, 12, 5, 7)])
def test_conv_gru_matches_the_module_formulation(B, C, H, W, deferred):
    from taming_event_flow_b200.flownet import ConvGRUCell

    torch.manual_seed(C + H)
    cell = ConvGRUCell(C).cuda().to(memory_format=torch.channels_last)
    with torch.no_grad():
        cell.gate_zr.bias.normal_(0, 0.3)
        cell.gate_c.bias.normal_(0, 0.3)
    x0 = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    h0 = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    gout = torch.randn(B, C, H, W, device="cuda")
    res = {}
    for fused in (False, True):
        cell.fused = fused
        cell.stacks.begin(2 if (fused and deferred) else 0)          # one weight gradient for the two steps (netops.WindowStacks)
        x, h = x0.clone().requires_grad_(True), h0.clone().requires_grad_(True)
        for p in cell.parameters():
            p.grad = None
        # two steps of the recurrence, so the state gradient flows through a cell as well
        s1 = cell(x, h)
        s2 = cell(x * 0.5, s1)
        (s2 * gout).sum().backward()
        res[fused] = [s1, s2, x.grad, h.grad] + [p.grad.clone() for p in cell.parameters()]
    for a, b in zip(res[True], res[False]):
        assert rel(a, b) < TOL


@pytest.mark.parametrize("act", ["relu", "tanh", "none"])
@pytest.mark.parametrize("stride,residual", [(1, False), (2, False), (1, True)])
def test_conv_bias_act_matches_pytorch(act, stride, residual):
    from taming_event_flow_b200 import netops

    torch.manual_seed(3)
    B, Ci, Co, H, W = 2, 8 if residual else 6, 8, 10, 14
    conv = torch.nn.Conv2d(Ci, Co, 3, stride, 1).cuda()
    x0 = torch.randn(B, Ci, H, W, device="cuda")
    out = {}
    for fused in (False, True):
        x = x0.clone().requires_grad_(True)
        conv.zero_grad(set_to_none=True)
        r = x * 0.7 if residual else None
        if fused:
            y = netops.conv_bias_act(x, conv.weight, conv.bias, r, act, stride, 1)
        else:
            y = conv(x) + (r if residual else 0)
            y = torch.relu(y) if act == "relu" else (torch.tanh(y) if act == "tanh" else y)
        (y * torch.linspace(-1, 1, y.numel(), device="cuda").view_as(y)).sum().backward()
        out[fused] = [y, x.grad, conv.weight.grad.clone(), conv.bias.grad.clone()]
    for a, b in zip(out[True], out[False]):
        assert rel(a, b) < TOL


@pytest.mark.parametrize("h,w,H,W", [(16, 16, 128, 128), (32, 32, 128, 128), (128, 128, 128, 128), (15, 20, 120, 160), (7, 9, 20, 31)])
@pytest.mark.parametrize("channels_last", [False, True])
def test_upsample_scale_matches_interpolate(h, w, H, W, channels_last):
    from taming_event_flow_b200 import netops

    torch.manual_seed(h)
    p0 = torch.randn(3, 2, h, w, device="cuda")
    if channels_last:
        p0 = p0.contiguous(memory_format=torch.channels_last)
    g = torch.randn(3, 2, H, W, device="cuda")
    out = {}
    for fused in (False, True):
        p = p0.clone(memory_format=torch.preserve_format).requires_grad_(True)
        y = netops.upsample_scale(p, (H, W), 64.0) if fused else F.interpolate(p, size=(H, W), mode="bilinear", align_corners=False) * 64.0
        (y * g).sum().backward()
        out[fused] = [y, p.grad]
    assert out[True][0].is_contiguous()
    for a, b in zip(out[True], out[False]):
        assert rel(a, b) < TOL


@pytest.mark.parametrize("scale", [2, 1])        # 2: the cell-wise x2 kernel every decoder stage uses; 1: the generic kernel
@pytest.mark.parametrize("B,C,h,w,with_pred,with_skip", [(2, 8, 6, 9, True, True), (1, 64, 16, 16, True, True), (2, 512, 4, 4, False, True), (2, 32, 5, 3, True, False),
                                                         (1, 6, 1, 1, True, True)])
def test_decoder_up_matches_add_cat_interpolate(B, C, h, w, with_pred, with_skip, scale):
    from taming_event_flow_b200 import netops

    torch.manual_seed(C)
    cl = lambda t: t.contiguous(memory_format=torch.channels_last)
    x0, s0, p0 = cl(torch.randn(B, C, h, w, device="cuda")), cl(torch.randn(B, C, h, w, device="cuda")), cl(torch.randn(B, 2, h, w, device="cuda"))
    g = torch.randn(B, C + (2 if with_pred else 0), scale * h, scale * w, device="cuda")
    out = {}
    for fused in (False, True):
        x, sk, pr = x0.clone().requires_grad_(True), s0.clone().requires_grad_(True), p0.clone().requires_grad_(True)
        if fused:
            y = netops.decoder_up(x, sk if with_skip else None, pr if with_pred else None, scale)
        else:
            y = x + sk if with_skip else x
            if with_pred:
                y = torch.cat([pr, y], 1)
            y = F.interpolate(y, scale_factor=scale, mode="bilinear", align_corners=False)
        (y * g).sum().backward()
        out[fused] = [y, x.grad] + ([sk.grad] if with_skip else []) + ([pr.grad] if with_pred else [])
    for a, b in zip(out[True], out[False]):
        assert rel(a, b) < TOL


@pytest.mark.parametrize("window", [0, 3, 2])
def test_fused_network_matches_the_plain_network_over_a_recurrent_window(window):
    """RecEVFlowNet(fused=True) against fused=False from the same weights: flow maps of three recurrent passes and every parameter
    gradient of a loss on them."""
    from taming_event_flow_b200.flownet import RecEVFlowNet

    torch.manual_seed(0)
    net = RecEVFlowNet(num_bins=2, base_channels=8).cuda().to(memory_format=torch.channels_last)
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.1)
    xs = [torch.rand(2, 2, 64, 64, device="cuda").contiguous(memory_format=torch.channels_last) for _ in range(3)]
    gs = [[torch.randn(2, 2, 64, 64, device="cuda") for _ in range(4)] for _ in range(3)]
    out = {}
    for fused in (False, True):
        net.fused = fused
        for c in net.enc_gru:
            c.fused = fused
        net.reset_states()
        net.zero_grad(set_to_none=True)
        net.begin_window(window)             # 3: deferred weight gradients over the window; 2: the third pass falls back to its own; 0: off
        loss, flows = 0.0, []
        for x, g in zip(xs, gs):
            fl = net(x, flow_scaling=32.0)["flow"]
            flows += fl
            loss = loss + sum((f * gg).sum() for f, gg in zip(fl, g))
        loss.backward()
        out[fused] = flows + [p.grad.clone() for p in net.parameters()]
    worst = max(rel(a, b) for a, b in zip(out[True], out[False]))
    assert worst < 5 * TOL, worst          # 60 convolutions deep: rounding of the two routes' sums accumulates


def test_netops_reject_cpu_tensors():
    from taming_event_flow_b200 import _lib, netops

    with pytest.raises(_lib.TefError):
        netops.upsample_scale(torch.zeros(1, 2, 4, 4), (8, 8), 1.0)
