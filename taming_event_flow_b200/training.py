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


def max_over_ranks(value, device):
    """Max of a python float over all ranks (device timing is reported as the slowest rank)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def train_step(model, loss_fn, optimizer, windows, flow_scaling=32.0, clip_grad=100.0, world_size=1, encode=None):
    """One optimizer step over a loss window of P passes (upstream ``train_flow.py:106-137``).

    `windows[t]` = (event_list [B,N,4], pol_mask [B,N,2], d_event_list, d_pol_mask) on the device; `encode` maps an
    event list to the network input (default: batched per-polarity event counts of grad + detached events).
    Returns the (un-synchronised) loss tensor.
    """
    from .dataloader.encodings import events_to_channels_batched

    res = loss_fn.res
    loss_fn.reset()
    for ev, mk, dev, dmk in windows:
        if encode is not None:
            x = encode(ev, dev)
        else:
            x = events_to_channels_batched(torch.cat([ev, dev], 1) if dev.shape[1] else ev, res)
        flows = [f * flow_scaling for f in model(x)["flow"]]         # px / input window (train_flow.py:107-108)
        loss_fn.update(flows, ev, mk, dev, dmk)
    loss = loss_fn()
    loss.backward()
    allreduce_gradients_sum(list(model.parameters()), world_size)
    if clip_grad is not None:
        torch.nn.utils.clip_grad_norm_(model.parameters(), clip_grad)
    optimizer.step()
    optimizer.zero_grad(set_to_none=True)
    model.detach_states()
    return loss.detach()
