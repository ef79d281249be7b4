"""Recurrent EV-FlowNet in plain PyTorch, for the training-step workloads of ``bench.py`` / ``train_synthetic.py``.

``north_star`` keeps the flow networks on PyTorch (cuDNN convolutions); this module is NOT part of the accelerated path.
It re-states the topology of upstream ``RecEVFlowNet`` (``models/model.py:6-85``, ``models/arch.py:177-242``,
``models/submodules.py``) so that the synthetic training step has the same cost and parameter count
(31,365,352 parameters with 2 input channels):

    4 x [3x3 stride-2 conv + ReLU -> ConvGRU]   channels 64, 128, 256, 512
    2 x residual block at 512 channels
    4 x [skip (sum) -> (concat previous 2-ch prediction) -> bilinear x2 -> 3x3 conv + ReLU]   256, 128, 64, 32
    a 1x1 tanh flow head after every decoder, each prediction up-sampled to the input size and scaled by 2^(3-i)
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import netops


def _conv(cin, cout, k, stride=1, w_scale=None):
    conv = nn.Conv2d(cin, cout, k, stride, k // 2)
    s = math.sqrt(1.0 / cin) if w_scale is None else w_scale
    nn.init.uniform_(conv.weight, -s, s)
    nn.init.zeros_(conv.bias)
    return conv


class ConvGRUCell(nn.Module):
    """Upstream's ConvGRU (``models/submodules.py:111-152``) with the update and the reset gate in ONE convolution
    (``gate_zr``: filters of the update gate stacked over those of the reset gate, each block initialised orthogonally like
    upstream's separate gates) -- same function, same parameter count.  On fp32 CUDA tensors the cell runs through
    ``netops.conv_gru``: two convolutions and two fused kernels forward, three fused kernels backward."""

    def __init__(self, channels, fused=True):
        super().__init__()
        self.channels, self.fused = channels, fused
        self.stacks = netops.WindowStacks()       # deferred weight gradient over a loss window (RecEVFlowNet.begin_window)
        self.gate_zr = nn.Conv2d(2 * channels, 2 * channels, 3, padding=1)
        self.gate_c = nn.Conv2d(2 * channels, channels, 3, padding=1)
        with torch.no_grad():
            nn.init.orthogonal_(self.gate_zr.weight[:channels])
            nn.init.orthogonal_(self.gate_zr.weight[channels:])
            nn.init.orthogonal_(self.gate_c.weight)
            nn.init.zeros_(self.gate_zr.bias)
            nn.init.zeros_(self.gate_c.bias)

    def forward(self, x, h):
        if h is None:
            h = torch.zeros_like(x)
        if self.fused and netops.usable(x, h):
            return netops.conv_gru(x, h, self.gate_zr.weight, self.gate_zr.bias, self.gate_c.weight, self.gate_c.bias, self.stacks)
        z, r = torch.sigmoid(self.gate_zr(torch.cat([x, h], 1))).chunk(2, 1)
        cand = torch.tanh(self.gate_c(torch.cat([x, h * r], 1)))
        return h * (1 - z) + cand * z


class RecEVFlowNet(nn.Module):
    folds_flow_scaling = True         # forward(x, flow_scaling=...) multiplies the flow maps itself

    def __init__(self, num_bins=2, base_channels=64, num_encoders=4, final_w_scale=0.01, fp32_heads=True, fused=True):
        super().__init__()
        self.fp32_heads = fp32_heads      # under autocast: 1x1 flow heads, tanh and up-sampling in fp32 (bf16 would quantise the flow to 0.4 %)
        self.fused = fused                # fp32 CUDA tensors go through the fused element-wise kernels of netops (csrc/tef_net.cu)
        chans = [base_channels * 2 ** i for i in range(num_encoders)]            # 64 128 256 512
        ins = [num_bins] + chans[:-1]
        self.enc_conv = nn.ModuleList([_conv(i, o, 3, stride=2) for i, o in zip(ins, chans)])
        self.enc_gru = nn.ModuleList([ConvGRUCell(o, fused) for o in chans])
        self.res = nn.ModuleList([nn.ModuleList([_default_conv(chans[-1]), _default_conv(chans[-1])]) for _ in range(2)])
        dec_in = list(reversed(chans))                                           # 512 256 128 64
        dec_out = [c // 2 for c in dec_in]                                       # 256 128 64 32
        self.dec = nn.ModuleList([_conv(i + (0 if k == 0 else 2), o, 3) for k, (i, o) in enumerate(zip(dec_in, dec_out))])
        self.heads = nn.ModuleList([_conv(o, 2, 1, w_scale=final_w_scale) for o in dec_out])
        self.num_encoders = num_encoders
        self.states = [None] * num_encoders
        # deferred weight gradients over a loss window (begin_window): one WindowStacks per convolution the fused path runs
        self._stacks = {id(c): netops.WindowStacks() for c in list(self.enc_conv) + [c for pair in self.res for c in pair] + list(self.dec)}

    def begin_window(self, passes):
        """Open a loss window of `passes` forward calls whose loss back-propagates through all of them (upstream's training loop,
        train_flow.py:106-137): the recurrent layers then compute their weight gradients once per window instead of once per
        pass (netops.WindowStacks).  `passes` = 0 switches back to per-pass gradients."""
        lost = None
        for st in [cell.stacks for cell in self.enc_gru] + list(self._stacks.values()):
            try:
                st.begin(passes if self.fused else 0)
            except RuntimeError as exc:          # the window before lost a layer's weight gradient: re-arm every layer, then report
                lost = lost or exc
                st.begin(passes if self.fused else 0)
        if lost is not None:
            raise lost

    def reset_states(self):
        self.states = [None] * self.num_encoders

    def detach_states(self):
        self.states = [None if s is None else s.detach() for s in self.states]

    def _conv_act(self, conv, x, act="relu", residual=None, slot=None):
        """act(conv(x) + residual): one cuDNN convolution + one fused kernel (bias, residual, activation) where netops applies."""
        if self.fused and conv.out_channels % 4 == 0 and netops.usable(residual) and netops.usable_input(x):
            return netops.conv_bias_act(x, conv.weight, conv.bias, residual, act, conv.stride[0], conv.padding[0], self._stacks.get(id(conv)), slot)
        y = conv(x)
        if residual is not None:
            y = y + residual
        return torch.relu(y) if act == "relu" else (torch.tanh(y) if act == "tanh" else y)

    def _head(self, i, x, size, scale):
        """1x1 flow head + tanh, up-sampled to the input size and scaled (models/model.py:65-85): returns (prediction, flow map)."""
        if self.fp32_heads and torch.is_autocast_enabled(x.device.type):
            with torch.autocast(x.device.type, enabled=False):
                pred = torch.tanh(self.heads[i](x.float()))
                return pred, F.interpolate(pred, size=size, mode="bilinear", align_corners=False) * scale
        pred = torch.tanh(self.heads[i](x))
        if self.fused and netops.usable_input(pred):
            return pred, netops.upsample_scale(pred, size, scale)
        return pred, F.interpolate(pred, size=size, mode="bilinear", align_corners=False) * scale

    def forward(self, x, flow_scaling=1.0):
        """`flow_scaling`: extra factor folded into the flow maps (upstream multiplies the network output by
        config["loss"]["flow_scaling"] in the training loop, train_flow.py:106-108)."""
        # upstream pads the input on the top / left to a multiple of 16 and crops the flow maps again (models/model_util.py:29-71,
        # models/model.py:66-85); a no-op for the benchmark resolutions
        ph, pw = (16 - x.shape[2] % 16) % 16, (16 - x.shape[3] % 16) % 16
        if ph or pw:
            x = F.pad(x, (pw, 0, ph, 0))
        H, W = x.shape[2], x.shape[3]
        skips = []
        for i in range(self.num_encoders):
            x = self._conv_act(self.enc_conv[i], x)
            x = self.enc_gru[i](x, self.states[i])
            self.states[i] = x
            skips.append(x)
        for c1, c2 in self.res:
            x = self._conv_act(c2, self._conv_act(c1, x), residual=x)
        flows, pred = [], None
        for i in range(self.num_encoders):
            skip = skips[self.num_encoders - 1 - i]
            slot = None
            if self.fused and netops.usable(x, skip) and (pred is None or pred.dtype == torch.float32):
                # skip sum, concatenation and x2 up-sampling in one kernel, written straight into the decoder convolution's slot
                conv, st = self.dec[i], self._stacks[id(self.dec[i])]
                slot = st.take_slot() if conv.out_channels % 4 == 0 else None        # (else the convolution runs unfused)
                out = None
                if slot is not None:
                    shp = (x.shape[0], x.shape[1] + (2 if pred is not None else 0), 2 * x.shape[2], 2 * x.shape[3])
                    out = netops.conv_input_slot(st, slot, shp, conv.weight, conv.stride[0], conv.padding[0], x.device)
                x = netops.decoder_up(x, skip, pred, 2, out=out)
            else:
                x = x + skip
                if pred is not None:
                    x = torch.cat([pred, x], 1)
                x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
            x = self._conv_act(self.dec[i], x, slot=slot)
            pred, flow = self._head(i, x, (H, W), float(2 ** (self.num_encoders - 1 - i)) * float(flow_scaling))
            flows.append(flow[..., ph:, pw:].contiguous() if (ph or pw) else flow)
        return {"flow": flows}


def _default_conv(c):
    return nn.Conv2d(c, c, 3, padding=1)


def count_parameters(m):
    return sum(p.numel() for p in m.parameters())


def from_upstream_state_dict(sd):
    """State dict of upstream's ``RecEVFlowNet`` (``models/model.py``; keys ``arch.encoders.i.conv.conv2d.*``,
    ``arch.encoders.i.recurrent_block.{update,reset,out}_gate.*``, ``arch.resblocks.j.conv{1,2}.*``, ``arch.decoders.i.conv2d.*``,
    ``arch.preds.i.conv2d.*``) -> the names of this module, so that a checkpoint trained with the reference loads into
    ``RecEVFlowNet.load_state_dict``.  The update and the reset gate are stacked into ``gate_zr`` (update first)."""
    out = {}
    n_enc = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("arch.encoders."))
    n_res = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("arch.resblocks."))
    for i in range(n_enc):
        for wb in ("weight", "bias"):
            out["enc_conv.%d.%s" % (i, wb)] = sd["arch.encoders.%d.conv.conv2d.%s" % (i, wb)]
            g = "arch.encoders.%d.recurrent_block." % i
            out["enc_gru.%d.gate_zr.%s" % (i, wb)] = torch.cat([sd[g + "update_gate." + wb], sd[g + "reset_gate." + wb]], 0)
            out["enc_gru.%d.gate_c.%s" % (i, wb)] = sd[g + "out_gate." + wb]
            out["dec.%d.%s" % (i, wb)] = sd["arch.decoders.%d.conv2d.%s" % (i, wb)]
            out["heads.%d.%s" % (i, wb)] = sd["arch.preds.%d.conv2d.%s" % (i, wb)]
    for j in range(n_res):
        for wb in ("weight", "bias"):
            out["res.%d.0.%s" % (j, wb)] = sd["arch.resblocks.%d.conv1.%s" % (j, wb)]
            out["res.%d.1.%s" % (j, wb)] = sd["arch.resblocks.%d.conv2.%s" % (j, wb)]
    return out


def to_upstream_state_dict(sd):
    """The inverse of ``from_upstream_state_dict``: this module's state dict under upstream's parameter names."""
    out = {}
    n_enc = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("enc_conv."))
    n_res = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("res."))
    for i in range(n_enc):
        for wb in ("weight", "bias"):
            out["arch.encoders.%d.conv.conv2d.%s" % (i, wb)] = sd["enc_conv.%d.%s" % (i, wb)]
            zr = sd["enc_gru.%d.gate_zr.%s" % (i, wb)]
            g = "arch.encoders.%d.recurrent_block." % i
            out[g + "update_gate." + wb], out[g + "reset_gate." + wb] = zr[:zr.shape[0] // 2], zr[zr.shape[0] // 2:]
            out[g + "out_gate." + wb] = sd["enc_gru.%d.gate_c.%s" % (i, wb)]
            out["arch.decoders.%d.conv2d.%s" % (i, wb)] = sd["dec.%d.%s" % (i, wb)]
            out["arch.preds.%d.conv2d.%s" % (i, wb)] = sd["heads.%d.%s" % (i, wb)]
    for j in range(n_res):
        for wb in ("weight", "bias"):
            out["arch.resblocks.%d.conv1.%s" % (j, wb)] = sd["res.%d.0.%s" % (j, wb)]
            out["arch.resblocks.%d.conv2.%s" % (j, wb)] = sd["res.%d.1.%s" % (j, wb)]
    return out
