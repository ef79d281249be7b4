"""Host-side multi-process logic on CPU (gloo, world_size 2): batch sharding, SUM gradient all-reduce equals the
single-process global batch, max-over-ranks timing.  The CM kernels themselves have no CPU path; a stand-in loss with the
same reduction semantics (a SUM over samples) is used here."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from taming_event_flow_b200.training import allreduce_gradients_sum, max_over_ranks, shard_range


def test_shard_range_partitions_the_batch():
    for gb in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(gb, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(2, 4, 3, padding=1), torch.nn.Tanh(), torch.nn.Conv2d(4, 2, 1))


def _loss(model, x):
    # sum over samples of a per-sample normalised quantity: the reduction structure of focus_loss (loss/flow.py:122-129)
    y = model(x)
    per_sample = (y ** 2).flatten(1).sum(1) / (x.flatten(1).abs().sum(1) + 1e-9)
    return per_sample.sum()


def _worker(rank, world, port, gb, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(123)
    data = torch.randn(gb, 2, 8, 8)
    a, b = shard_range(gb, world, rank)
    model = _model()
    _loss(model, data[a:b]).backward()
    nbytes = allreduce_gradients_sum(list(model.parameters()), world, bucket_bytes=256)   # several buckets
    slow = max_over_ranks(1.0 + rank, torch.device("cpu"))
    if rank == 0:
        torch.save({"grads": [p.grad.clone() for p in model.parameters()], "nbytes": nbytes, "slow": slow}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sum_allreduce_matches_single_process(tmp_path):
    gb, world = 6, 2
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), gb, out), nprocs=world, join=True)
    got = torch.load(out)
    torch.manual_seed(123)
    data = torch.randn(gb, 2, 8, 8)
    model = _model()
    _loss(model, data).backward()
    for g, p in zip(got["grads"], model.parameters()):
        assert torch.allclose(g, p.grad, rtol=1e-5, atol=1e-7)
    assert got["nbytes"] == sum(p.numel() * 4 for p in model.parameters())
    assert got["slow"] == 2.0


def _reducer_worker(rank, world, port, gb, out):
    from taming_event_flow_b200.training import GradReducer

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(123)
    data = torch.randn(gb, 2, 8, 8)
    a, b = shard_range(gb, world, rank)
    model = _model().to(memory_format=torch.channels_last)
    red = GradReducer(list(model.parameters()), world, bucket_bytes=16)          # several buckets, issued from the grad hooks
    res = []
    for step in range(2):                                                        # the second step checks the re-arming
        red.zero()
        # a weight-shared "unrolled" loss, like back-propagation through time: every parameter is used twice
        (_loss(model, data[a:b]) + 0.5 * _loss(model, data[a:b] * (1.0 + step))).backward()
        launched_in_backward = sum(red._launched)
        nbytes = red.finish()
        res.append({"grads": [p.grad.clone() for p in model.parameters()], "nbytes": nbytes, "in_backward": launched_in_backward,
                    "views": all(p.grad.untyped_storage().data_ptr() == red.flat.untyped_storage().data_ptr() for p in model.parameters()),
                    "strides": all(p.grad.stride() == p.stride() for p in model.parameters()), "buckets": len(red.buckets)})
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


def test_grad_reducer_overlapped_buckets_match_single_process(tmp_path):
    """GradReducer: flat gradient buffer (views with the parameters' strides), bucketed SUM all-reduce issued from
    post-accumulate-grad hooks while the backward pass is still running; equals the single-process global batch."""
    gb, world = 6, 2
    out = str(tmp_path / "r0.pt")
    mp.spawn(_reducer_worker, args=(world, _free_port(), gb, out), nprocs=world, join=True)
    got = torch.load(out)
    torch.manual_seed(123)
    data = torch.randn(gb, 2, 8, 8)
    for step, r in enumerate(got):
        model = _model()
        (_loss(model, data) + 0.5 * _loss(model, data * (1.0 + step))).backward()
        for g, p in zip(r["grads"], model.parameters()):
            assert torch.allclose(g, p.grad, rtol=1e-5, atol=1e-7)
        assert r["nbytes"] == sum(p.numel() * 4 for p in model.parameters())
        assert r["views"] and r["strides"] and r["buckets"] >= 2
        assert r["in_backward"] == r["buckets"]            # every bucket was issued from a hook, before backward() returned


def _net_window_loss(net, xs):
    """Stand-in for the CM loss over a recurrent window: a SUM over samples of a per-sample quantity of every pass's flow maps."""
    net.reset_states()
    net.begin_window(len(xs))
    loss = 0.0
    for x in xs:
        for f in net(x, flow_scaling=32.0)["flow"]:
            loss = loss + (f ** 2).flatten(1).mean(1).sum()
    return loss


def _net_worker(rank, world, port, gb, out):
    from taming_event_flow_b200.flownet import RecEVFlowNet
    from taming_event_flow_b200.training import GradReducer

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(5)
    net = RecEVFlowNet(num_bins=2, base_channels=4)
    xs = [torch.rand(gb, 2, 32, 32) for _ in range(2)]
    a, b = shard_range(gb, world, rank)
    red = GradReducer(list(net.parameters()), world, bucket_bytes=4096)
    red.zero()
    _net_window_loss(net, [x[a:b] for x in xs]).backward()
    red.finish()
    if rank == 0:
        torch.save([p.grad.clone() for p in net.parameters()], out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_recurrent_window_matches_the_global_batch(tmp_path):
    """The recurrent network over a two-pass window (back-propagation through time, states per shard), batch sharded over two
    ranks, SUM all-reduce from the flat gradient buffer: equals one process on the global batch."""
    from taming_event_flow_b200.flownet import RecEVFlowNet

    gb, world = 4, 2
    out = str(tmp_path / "r0.pt")
    mp.spawn(_net_worker, args=(world, _free_port(), gb, out), nprocs=world, join=True)
    got = torch.load(out)
    torch.manual_seed(5)
    net = RecEVFlowNet(num_bins=2, base_channels=4)
    xs = [torch.rand(gb, 2, 32, 32) for _ in range(2)]
    _net_window_loss(net, xs).backward()
    for g, p in zip(got, net.parameters()):
        assert torch.allclose(g, p.grad, rtol=1e-4, atol=1e-6)
