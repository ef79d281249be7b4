"""Data-parallel training step around the CM loss (upstream loop: ``train_flow.py:81-137``).

One process per GPU; the batch of independent event-window sequences is sharded across ranks.  The CM loss needs no
collective (it is a sum over samples, upstream ``loss/flow.py:122-129``); the only exchange is one all-reduce of the
network gradients per optimizer step.  Because the loss SUMS over the batch, gradients are reduced with SUM (not
DistributedDataParallel's mean), which makes N ranks x B/N sequences identical to one rank x B sequences before
``clip_grad_norm_`` (upstream ``train_flow.py:127-128``).
"""
import torch
import torch.distributed as dist


def shard_range(global_batch, world_size, rank):
    """Contiguous shard [start, stop) of `global_batch` sequences for `rank`; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(global_batch, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_gradients_sum(params, world_size=None, bucket_bytes=64 << 20):
    """SUM all-reduce of `.grad` over the default process group in flat buckets (NCCL over NVLink on GPUs, gloo in
    the CPU tests).  No-op for a single process."""
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if world_size <= 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    nbytes, bucket, size = 0, [], 0

    def flush():
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        nbytes += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
            bucket, size = [], 0
    flush()
    return nbytes


class GradReducer:
    """Gradients of all parameters in ONE flat fp32 buffer (``p.grad`` are views with the parameter's own strides), reduced
    with SUM in buckets that start as soon as the backward pass has finished the parameters of a bucket.

    * no ``torch.cat`` / copy-back round trip (two extra passes over 125 MB per step);
    * buckets follow the REVERSE registration order (decoder first, encoder last), the order in which autograd finishes
      them; each bucket's all-reduce is issued from ``register_post_accumulate_grad_hook`` and runs on NCCL's stream under
      the rest of the backward pass.  With back-propagation through time every pass re-uses every parameter, so a
      parameter's gradient is final only during the backward of the FIRST pass of the window: that is the overlap window;
    * ``zero()`` replaces ``zero_grad(set_to_none=True)``: the views stay in place, so the buffer is also a static address
      for CUDA-graph capture of the step.
    """

    def __init__(self, params, world_size=None, bucket_bytes=32 << 20):
        self.params = [p for p in params if p.requires_grad]
        if world_size is None:
            world_size = dist.get_world_size() if dist.is_initialized() else 1
        self.world_size = world_size
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.buckets, self._bucket_of, off, start, members = [], {}, 0, 0, 0
        for p in reversed(self.params):
            if not (p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))):
                raise ValueError("GradReducer needs dense parameters (contiguous or channels_last)")
            p.grad = torch.as_strided(self.flat, p.size(), p.stride(), off)      # same memory format as the parameter
            self._bucket_of[p] = len(self.buckets)
            off += p.numel()
            members += 1
            if (off - start) * 4 >= bucket_bytes:
                self.buckets.append((start, off, members))
                start, members = off, 0
        if off > start:
            self.buckets.append((start, off, members))
        self._pending = [m for _, _, m in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._works = []
        self._hooks = []
        self.overlap = True
        if world_size > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def zero(self):
        self.flat.zero_()

    def _launch(self, b):
        lo, hi, _ = self.buckets[b]
        self._launched[b] = True
        self._works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

    def _on_grad(self, p):
        if not self.overlap:
            return
        b = self._bucket_of[p]
        self._pending[b] -= 1
        if self._pending[b] == 0 and not self._launched[b]:
            self._launch(b)

    def finish(self):
        """Issue whatever has not been issued (parameters without a gradient this step, or overlap switched off), wait for all
        buckets and re-arm.  Returns the bytes reduced."""
        if self.world_size <= 1:
            return 0
        for b in range(len(self.buckets)):
            if not self._launched[b]:
                self._launch(b)
        for w in self._works:
            w.wait()
        self._works = []
        self._pending = [m for _, _, m in self.buckets]
        self._launched = [False] * len(self.buckets)
        return self.nbytes


def max_over_ranks(value, device):
    """Max of a python float over all ranks (device timing is reported as the slowest rank)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _forward_window(model, loss_fn, windows, flow_scaling, encode, autocast):
    from .dataloader.encodings import events_to_channels_batched

    res = loss_fn.res
    loss_fn.reset()
    if hasattr(model, "begin_window"):
        model.begin_window(len(windows))         # one weight gradient per recurrent layer and window (netops.WindowStacks)
    for ev, mk, dev, dmk in windows:
        if encode is not None:
            x = encode(ev, dev)
        else:
            x = events_to_channels_batched(torch.cat([ev, dev], 1) if dev.shape[1] else ev, res)
        # px / input window (train_flow.py:107-108); RecEVFlowNet folds the factor into its flow heads
        folds = getattr(model, "folds_flow_scaling", False)
        if autocast is not None:
            with torch.autocast("cuda", dtype=autocast):
                out = (model(x, flow_scaling=flow_scaling) if folds else model(x))["flow"]
            flows = [f.float() if folds else f.float() * flow_scaling for f in out]    # the CM loss is fp32 whatever the network computes in
        else:
            flows = model(x, flow_scaling=flow_scaling)["flow"] if folds else [f * flow_scaling for f in model(x)["flow"]]
        loss_fn.update(flows, ev, mk, dev, dmk)
    return loss_fn()


def train_step(model, loss_fn, optimizer, windows, flow_scaling=32.0, clip_grad=100.0, world_size=1, encode=None, reducer=None, autocast=None):
    """One optimizer step over a loss window of P passes (upstream ``train_flow.py:106-137``).

    `windows[t]` = (event_list [B,N,4], pol_mask [B,N,2], d_event_list, d_pol_mask) on the device; `encode` maps an
    event list to the network input (default: batched per-polarity event counts of grad + detached events).
    `reducer`: a `GradReducer` over the model's parameters (flat gradient buffer, bucketed SUM all-reduce overlapped with
    the backward pass); without one the gradients are reduced after the backward pass in copied buckets.
    `autocast`: e.g. torch.bfloat16 for the network (the loss stays fp32).  Returns the (un-synchronised) loss tensor.
    """
    if reducer is not None:
        reducer.zero()
    loss = _forward_window(model, loss_fn, windows, flow_scaling, encode, autocast)
    loss.backward()
    if reducer is not None:
        reducer.finish()
    else:
        allreduce_gradients_sum(list(model.parameters()), world_size)
    if clip_grad is not None:
        torch.nn.utils.clip_grad_norm_(model.parameters(), clip_grad)
    optimizer.step()
    if reducer is None:
        optimizer.zero_grad(set_to_none=True)
    model.detach_states()
    return loss.detach()


class GraphedTrainStep:
    """The forward pass over the P windows, the CM loss and the whole backward pass (through the loss kernels and through
    time) captured in ONE CUDA graph; gradient all-reduce, clipping and the optimizer step run eagerly after the replay.

    The eager step issues ~4 900 small kernels; replaying them removes the launch gaps between them.
    Requirements of a capture: fixed shapes (the synthetic workloads; a real loader would pad to a fixed N), the event
    tensors of a step are copied into the static `windows` before `step()`, gradients live in the `GradReducer`'s flat
    buffer, and the recurrent states are static tensors that each replay reads and writes back (upstream carries the
    states from one loss window to the next and detaches them, ``train_flow.py:136``)."""

    def __init__(self, model, loss_fn, optimizer, windows, flow_scaling=32.0, clip_grad=100.0, encode=None, reducer=None, autocast=None,
                 warmup=2, capture_collectives=False):
        self.model, self.loss_fn, self.opt, self.windows = model, loss_fn, optimizer, windows
        self.flow_scaling, self.clip_grad, self.encode, self.autocast = flow_scaling, clip_grad, encode, autocast
        self.reducer = reducer if reducer is not None else GradReducer(list(model.parameters()), world_size=1)
        self._work = [(ev.clone(), dv.clone()) for ev, _, dv, _ in windows]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):                 # eager steps: cuDNN picks its algorithms, every workspace reaches its size
                self.reducer.zero()
                self._fwd_bwd()
                self._tail()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # the recurrent states become static tensors: read at the start of every replay, written back at its end
        self.states = [s.detach().clone() for s in model.states]
        self.graph = torch.cuda.CUDAGraph()
        # capture_collectives (EXPERIMENTAL, off everywhere: with the fused network step a run at 8 GPUs never finished its first
        # replay -- gpurun call 148 of round 2 -- and it has not been debugged since): the bucketed all-reduces the gradient hooks issue (on NCCL's stream, forked from the capturing
        # stream) become branches of the graph and overlap the tail of the captured backward pass; otherwise they are issued
        # after the replay
        self.comm_captured = bool(capture_collectives and self.reducer.world_size > 1)
        overlap, self.reducer.overlap = self.reducer.overlap, self.comm_captured
        try:
            with torch.cuda.graph(self.graph, **({"capture_error_mode": "thread_local"} if self.comm_captured else {})):
                self.reducer.zero()
                model.states = list(self.states)
                self.loss = self._fwd_bwd()
                if self.comm_captured:
                    self.reducer.finish()
                for dst, src in zip(self.states, model.states):
                    dst.copy_(src.detach())
        finally:
            self.reducer.overlap = overlap
        model.states = list(self.states)

    def _fwd_bwd(self):
        wins = []
        for (ev, mk, dv, dmk), (wev, wdv) in zip(self.windows, self._work):
            wev.copy_(ev)                                   # update() adds the pass index in place: work on copies
            wdv.copy_(dv)
            wins.append((wev, mk, wdv, dmk))
        loss = _forward_window(self.model, self.loss_fn, wins, self.flow_scaling, self.encode, self.autocast)
        loss.backward()
        self.model.detach_states()
        return loss.detach()

    def _tail(self, reduce=True):
        if reduce:
            self.reducer.finish()
        if self.clip_grad is not None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip_grad)
        self.opt.step()

    def step(self):
        self.graph.replay()
        self._tail(reduce=not self.comm_captured)
        return self.loss
