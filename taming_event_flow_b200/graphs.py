"""CUDA-graph replay of a whole CM-loss window (``update`` x P -> ``forward`` -> ``backward``).

Small windows are bound by the host: config 1 (128x128, batch 8, 10 x (10 k + 10 k) events) spends ~0.4 ms of Python and
ctypes on its 20 launches, more than the kernels take.  Every C entry point of ``libtef_b200.so`` only enqueues work on the
stream it is given (no allocation, no synchronisation), so the launch sequence of a window with fixed shapes can be
captured once and replayed as ONE graph launch.

    gw = GraphedLossWindow(module, flows, events, masks, d_events, d_masks)     # static inputs: filled by the caller
    loss, grads = gw.replay()                                                    # static outputs

The tensors handed to the constructor ARE the static inputs: write the next window into them (``copy_``) and replay.
`update` adds the pass index to the caller's timestamps in place (upstream ``loss/flow.py:457-458``); the graph therefore
starts with a copy of the event tensors into private working buffers and leaves the static inputs untouched.
"""
import torch


class GraphedLossWindow:
    def __init__(self, module, flows, events, masks, d_events, d_masks, warmup=2):
        P = len(flows)
        self.module = module
        self.flows = [[f.detach().requires_grad_(True) for f in per] for per in flows]
        self.events, self.masks, self.d_events, self.d_masks = list(events), list(masks), list(d_events), list(d_masks)
        self._work = [(e.clone(), d.clone()) for e, d in zip(self.events, self.d_events)]
        self.P = P
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up off the capturing stream: grows the workspace to its final size
            for _ in range(max(1, warmup)):
                self._body()
                self._clear_grads()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._clear_grads()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()
        self.grads = [[f.grad for f in per] for per in self.flows]
        self.launches_per_replay = self._launches

    def _clear_grads(self):
        for per in self.flows:
            for f in per:
                f.grad = None

    def _body(self):
        import ctypes

        from . import _lib

        L = _lib.lib()
        L.tef_launch_count.restype = ctypes.c_long
        n0 = L.tef_launch_count()
        m = self.module
        m.reset()
        for t in range(self.P):
            ev, dv = self._work[t]
            ev.copy_(self.events[t])
            dv.copy_(self.d_events[t])
            m.update(self.flows[t], ev, self.masks[t], dv, self.d_masks[t])
        loss = m()
        loss.backward()
        self._launches = int(L.tef_launch_count() - n0)
        return loss.detach()

    def replay(self):
        """Run the captured window on whatever the static inputs hold now.  Returns the static loss tensor and the static
        gradient tensors ``grads[t][f]`` (overwritten by the next replay)."""
        self.graph.replay()
        return self.loss, self.grads
