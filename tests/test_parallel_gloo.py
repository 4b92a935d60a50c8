"""Batch-sharded data parallelism on CPU (gloo, world_size 2): the flat per-block gradient buckets,
the overlapped all-reduce hooks and the dead-parameter rule.  The averaged gradients of two ranks
that each saw half of a batch must equal the single-process gradients of the whole batch (mean
loss => average).  The model here is the CPU oracle net -- the bucket
code only sees nn.Parameters."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dstagnn_drought_b200.parallel import GradBuckets, bucket_of, live_parameters


class TinyNet(torch.nn.Module):
    """Same parameter naming scheme as the real net (BlockList.i.* + head), incl. dead params."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.BlockList = torch.nn.ModuleList()
        for _ in range(3):
            blk = torch.nn.Module()
            blk.lin = torch.nn.Linear(6, 6)
            blk.EmbedT = torch.nn.Linear(6, 6)            # dead in blocks >= 1
            blk.residual_conv = torch.nn.Linear(6, 6)     # dead in blocks >= 1
            self.BlockList.append(blk)
        self.final_fc = torch.nn.Linear(6, 2)

    def forward(self, x):
        for i, b in enumerate(self.BlockList):
            h = torch.tanh(b.lin(x))
            if i == 0:
                h = h + b.EmbedT(x) + b.residual_conv(x)
            x = h
        return self.final_fc(x)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, overlap, ret, zero_mode="buckets"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = TinyNet()
    buckets = GradBuckets(net, overlap=overlap)
    opt = torch.optim.SGD([p for p in net.parameters()], lr=0.0)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(8, 6, generator=g)
    y = torch.randn(8, 2, generator=g)
    xs, ys = x.chunk(world)[rank], y.chunk(world)[rank]
    for _ in range(3):                                   # later steps check that the hooks are re-armed
        if zero_mode == "buckets":
            buckets.zero()
        else:
            # the reference loop's own call (train_DSTAGNN_my.py:154): set_to_none=True detaches the views
            opt.zero_grad(set_to_none=(zero_mode == "none"))
        torch.nn.functional.smooth_l1_loss(net(xs), ys).backward()
        buckets.finish()
    for b, flat in buckets.buckets.items():              # every live gradient lives in its bucket again
        for p, v in zip(buckets._members[b], buckets._views[b]):
            assert p.grad is not None and p.grad.data_ptr() == v.data_ptr()
    if rank == 0:
        ret["grads"] = {k: (p.grad.clone() if p.grad is not None else None) for k, p in net.named_parameters()}
        ret["numel"] = buckets.numel
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap,zero_mode", [(True, "buckets"), (False, "buckets"), (True, "none"), (False, "none"),
                                               (True, "zeros")])
def test_two_rank_average_equals_full_batch(overlap, zero_mode):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, overlap, ret, zero_mode), nprocs=2, join=True)
    net = TinyNet()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(8, 6, generator=g)
    y = torch.randn(8, 2, generator=g)
    torch.nn.functional.smooth_l1_loss(net(x), y).backward()
    live = live_parameters(net)
    for k, p in net.named_parameters():
        got = ret["grads"][k]
        if k in live:
            assert got is not None and torch.allclose(got, p.grad, atol=1e-6), k
        else:
            assert got is None, f"{k} is a dead parameter: its grad must stay None"
    assert ret["numel"] == sum(p.numel() for p in live.values())


def test_bucket_assignment_and_dead_parameter_rule():
    net = TinyNet()
    live = live_parameters(net)
    assert "BlockList.0.EmbedT.weight" in live and "BlockList.1.EmbedT.weight" not in live
    assert "BlockList.2.residual_conv.bias" not in live and "final_fc.weight" in live
    assert bucket_of("BlockList.2.lin.weight") == "BlockList.2" and bucket_of("final_fc.bias") == "head"
