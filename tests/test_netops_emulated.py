"""Host-side wiring of the fused network operators (netops.py: autograd Functions, gradient routing, the deferred weight gradient
of a loss window, the decoder writing into its convolution's slot) on the CPU, with the C entry points of csrc/tef_net.cu EMULATED
by a few lines of torch each.  The emulation restates what every kernel computes (it is test infrastructure, like the oracle); the
kernels themselves are checked on the GPU by tests/test_netops_gpu.py.  What this pins without a GPU: that the Python side hands
the right tensors, in the right layout and order, to the right entry point, and that a whole recurrent window through the fused
operators gives the plain modules' flow maps and parameter gradients."""
import pytest
import torch
import torch.nn.functional as F

from taming_event_flow_b200 import _lib, netops
from taming_event_flow_b200.flownet import ConvGRUCell, RecEVFlowNet


def rows(t):
    """[B, C, H, W] channels_last tensor (or a 1-D vector) as the [M, C] rows the kernels see; a view, so writes land in `t`."""
    if t.dim() == 1:
        return t
    assert t.is_contiguous(memory_format=torch.channels_last), "the kernels expect NHWC-dense tensors"
    return t.permute(0, 2, 3, 1).reshape(-1, t.shape[1])


def val(x):
    return x.value if hasattr(x, "value") else x


class EmulatedLib:
    """The entry points of csrc/tef_net.cu on CPU tensors (pointers are the tensors themselves)."""

    calls = []

    def tef_gru_gates(self, zr, bias, xh, xrh, M, Cx, C, st):
        self.calls.append("gru_gates")
        z, x, o = rows(zr), rows(xh), rows(xrh)
        assert z.shape == (val(M), 2 * C) and x.shape == (val(M), Cx + C)
        if bias is not None:
            z += bias
        z.sigmoid_()
        o[:, :Cx] = x[:, :Cx]
        o[:, Cx:] = x[:, Cx:] * z[:, C:]
        return 0

    def tef_gru_output(self, c, bias, xh, zr, out, M, Cx, C, st):
        self.calls.append("gru_output")
        a, x, z = rows(c), rows(xh), rows(zr)
        if bias is not None:
            a += bias
        a.tanh_()
        rows(out)[:] = x[:, Cx:] * (1 - z[:, :C]) + a * z[:, :C]
        return 0

    def tef_gru_output_bwd(self, gout, cand, xh, zr, gc, gzr, gh, gbias_c, gbias_zr, M, Cx, C, st):
        self.calls.append("gru_output_bwd")
        g, a, h, z = rows(gout), rows(cand), rows(xh)[:, Cx:], rows(zr)[:, :C]
        rows(gc)[:] = g * z * (1 - a * a)
        rows(gzr)[:, :C] = g * (a - h) * (z * (1 - z))
        rows(gh)[:] = g * (1 - z)
        if gbias_c is not None:
            gbias_c += rows(gc).sum(0)
        if gbias_zr is not None:
            gbias_zr[:C] += rows(gzr)[:, :C].sum(0)
        return 0

    def tef_gru_gates_bwd(self, gxrh, xh, zr, gzr, gh, gbias_zr, M, Cx, C, st):
        self.calls.append("gru_gates_bwd")
        ghr, h, r = rows(gxrh)[:, Cx:], rows(xh)[:, Cx:], rows(zr)[:, C:]
        rows(gh)[:] += ghr * r
        rows(gzr)[:, C:] = ghr * h * (r * (1 - r))
        if gbias_zr is not None:
            gbias_zr[C:] += rows(gzr)[:, C:].sum(0)
        return 0

    def tef_gru_input_grads(self, gxrh, gxh, gx, gh, M, Cx, C, st):
        self.calls.append("gru_input_grads")
        rows(gx)[:] = rows(gxrh)[:, :Cx] + rows(gxh)[:, :Cx]
        rows(gh)[:] += rows(gxh)[:, Cx:]
        return 0

    def tef_bias_act(self, y, bias, res, act, M, C, st):
        self.calls.append("bias_act")
        v = rows(y)
        assert v.shape == (val(M), C)
        if bias is not None:
            v += bias
        if res is not None:
            v += rows(res)
        if act == 1:
            v.relu_()
        elif act == 2:
            v.tanh_()
        return 0

    def tef_bias_act_bwd(self, gy, y, gpre, gbias, act, M, C, st):
        self.calls.append("bias_act_bwd")
        g = rows(gy).clone()
        if act == 1:
            g = g * (rows(y) > 0)
        elif act == 2:
            g = g * (1 - rows(y) ** 2)
        if gpre is not gy or act:
            rows(gpre)[:] = g
        if gbias is not None:
            gbias += g.sum(0)
        return 0

    def tef_upsample_scale(self, pred, strides, h, w, out, B, H, W, scale, st):
        self.calls.append("upsample_scale")
        assert tuple(strides) == tuple(pred.stride()) and out.is_contiguous() and tuple(out.shape) == (B, 2, H, W)
        out.copy_(F.interpolate(pred, size=(H, W), mode="bilinear", align_corners=False) * val(scale))
        return 0

    def tef_upsample_scale_bwd(self, g, B, H, W, scale, gpred, strides, h, w, st):
        self.calls.append("upsample_scale_bwd")
        assert g.is_contiguous() and tuple(strides) == tuple(gpred.stride())
        with torch.enable_grad():
            p = torch.zeros(B, 2, h, w, requires_grad=True)
            y = F.interpolate(p, size=(H, W), mode="bilinear", align_corners=False) * val(scale)
        gpred.copy_(torch.autograd.grad(y, p, g)[0])
        return 0

    def tef_decoder_up(self, x, skip, pred, ps, out, B, h, w, C, H, W, st):
        self.calls.append("decoder_up")
        y = x + skip if skip is not None else x
        if pred is not None:
            assert tuple(ps) == tuple(pred.stride())
            y = torch.cat([pred, y], 1)
        assert out.is_contiguous(memory_format=torch.channels_last) and tuple(out.shape) == (B, y.shape[1], H, W)
        out.copy_(F.interpolate(y, size=(H, W), mode="bilinear", align_corners=False))
        return 0

    def tef_decoder_up_bwd(self, g, gx, gpred, ps, B, h, w, C, H, W, st):
        self.calls.append("decoder_up_bwd")
        Cp = C + (2 if gpred is not None else 0)
        with torch.enable_grad():
            p = torch.zeros(B, Cp, h, w, requires_grad=True)
            y = F.interpolate(p, size=(H, W), mode="bilinear", align_corners=False)
        gin = torch.autograd.grad(y, p, g)[0]
        if gpred is not None:
            gpred.copy_(gin[:, :2])
        rows(gx)[:] = rows(gin[:, Cp - C:].contiguous(memory_format=torch.channels_last))
        return 0


@pytest.fixture
def emulated(monkeypatch):
    lib = EmulatedLib()
    EmulatedLib.calls = []
    monkeypatch.setattr(_lib, "lib", lambda: lib)
    monkeypatch.setattr(_lib, "ptr", lambda t: t)
    monkeypatch.setattr(_lib, "stream", lambda: None)
    monkeypatch.setattr(_lib, "require_cuda", lambda *a: None)
    cpu_ok = lambda *ts: all(t is None or (t.dtype == torch.float32 and t.dim() == 4 and t.shape[1] % 4 == 0) for t in ts)
    monkeypatch.setattr(netops, "usable", cpu_ok)
    monkeypatch.setattr(netops, "usable_input", lambda x: x.dtype == torch.float32 and x.dim() == 4)
    return lib


def rel(a, b):
    return float((a.detach() - b.detach()).abs().max() / max(float(b.detach().abs().max()), 1e-30))


@pytest.mark.parametrize("deferred", [False, True])
def test_conv_gru_wiring(emulated, deferred):
    torch.manual_seed(0)
    cell = ConvGRUCell(8)
    with torch.no_grad():
        cell.gate_zr.bias.normal_(0, 0.3)
        cell.gate_c.bias.normal_(0, 0.3)
    x0, h0, g = torch.randn(2, 8, 6, 10), torch.randn(2, 8, 6, 10), torch.randn(2, 8, 6, 10)
    res = {}
    for fused in (False, True):
        cell.fused = fused
        cell.stacks.begin(2 if (fused and deferred) else 0)
        cell.zero_grad(set_to_none=True)
        x, h = x0.clone().requires_grad_(True), h0.clone().requires_grad_(True)
        s1 = cell(x, h)
        s2 = cell(x * 0.5, s1)
        (s2 * g).sum().backward()
        res[fused] = [s1, s2, x.grad, h.grad] + [p.grad.clone() for p in cell.parameters()]
    assert "gru_gates" in emulated.calls and "gru_input_grads" in emulated.calls
    for a, b in zip(res[True], res[False]):
        assert rel(a, b) < 1e-5


@pytest.mark.parametrize("window", [0, 3, 2])
def test_fused_network_window_wiring(emulated, window):
    """RecEVFlowNet through the (emulated) fused operators against its plain modules: flow maps of three recurrent passes and every
    parameter gradient; window = 3 defers every weight gradient to the first pass's backward, 2 leaves the third pass on its own."""
    torch.manual_seed(1)
    net = RecEVFlowNet(num_bins=2, base_channels=4)
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.1)
    xs = [torch.rand(2, 2, 32, 48) for _ in range(3)]
    gs = [[torch.randn(2, 2, 32, 48) for _ in range(4)] for _ in range(3)]
    out = {}
    for fused in (False, True):
        net.fused = fused
        for c in net.enc_gru:
            c.fused = fused
        net.reset_states()
        net.zero_grad(set_to_none=True)
        net.begin_window(window)
        loss, flows = 0.0, []
        for x, g in zip(xs, gs):
            fl = net(x, flow_scaling=32.0)["flow"]
            flows += fl
            loss = loss + sum((f * gg).sum() for f, gg in zip(fl, g))
        loss.backward()
        out[fused] = flows + [p.grad.clone() for p in net.parameters()]
    used = set(emulated.calls)
    assert {"gru_gates", "gru_output", "gru_output_bwd", "gru_gates_bwd", "gru_input_grads", "bias_act", "bias_act_bwd", "decoder_up", "decoder_up_bwd",
            "upsample_scale", "upsample_scale_bwd"} <= used
    worst = max(rel(a, b) for a, b in zip(out[True], out[False]))
    assert worst < 2e-5, worst
    if window == 3:
        # the decoder's up-sampled input was written straight into its convolution's slot: no copy in between
        st = net._stacks[id(net.dec[0])]
        assert st.used == 3 and st.x["x"].shape[0] == 3 * 2


def test_a_fused_operator_back_propagates_once(emulated):
    conv = torch.nn.Conv2d(4, 4, 3, padding=1)
    x = torch.randn(1, 4, 5, 5, requires_grad=True)
    y = netops.conv_bias_act(x, conv.weight, conv.bias, None, "relu", 1, 1)
    y.sum().backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="once"):
        y.sum().backward()


def test_decoder_up_rejects_scales_its_adjoint_cannot_take(emulated):
    with pytest.raises(_lib.TefError):
        netops.decoder_up(torch.zeros(1, 4, 4, 4), None, None, 4)


def _window_grads(net, xs, use, window):
    """Gradients of a loss that only looks at the flow maps of the passes in `use`."""
    net.reset_states()
    net.zero_grad(set_to_none=True)
    net.begin_window(window)
    loss = 0.0
    for t, x in enumerate(xs):
        fl = net(x)["flow"]
        if t in use:
            loss = loss + sum((f ** 2).sum() for f in fl)
    loss.backward()
    return [None if p.grad is None else p.grad.clone() for p in net.parameters()]


def _set_fused(net, fused):
    net.fused = fused
    for c in net.enc_gru:
        c.fused = fused


def test_passes_that_do_not_reach_the_loss_contribute_zero_to_the_window_gradient(emulated):
    """The decoder of a pass the loss ignores never runs its backward: its rows of the gradient stacks (stale from the window before)
    must not leak into the deferred weight gradient."""
    torch.manual_seed(2)
    net = RecEVFlowNet(num_bins=2, base_channels=4)
    xs = [torch.rand(1, 2, 32, 32) for _ in range(3)]
    _set_fused(net, True)
    _window_grads(net, xs, {0, 1, 2}, 3)                      # fills every slot of the gradient stacks
    got = _window_grads(net, xs, {0, 2}, 3)                   # the middle pass reaches the loss through the recurrent states only
    _set_fused(net, False)
    want = _window_grads(net, xs, {0, 2}, 0)
    assert max(rel(a, b) for a, b in zip(got, want)) < 2e-5


def test_a_window_whose_first_pass_never_back_propagates_is_reported(emulated):
    torch.manual_seed(3)
    net = RecEVFlowNet(num_bins=2, base_channels=4)
    xs = [torch.rand(1, 2, 32, 32) for _ in range(2)]
    _set_fused(net, True)
    _window_grads(net, xs, {1}, 2)                            # decoders and heads of pass 0 never run their backward
    with pytest.raises(RuntimeError, match="deferred weight gradient lost"):
        net.begin_window(2)
    net.begin_window(2)                                       # reported once; the next window starts clean
    _window_grads(net, xs, {0, 1}, 2)
    net.begin_window(0)
