"""The step either side of the block stack on the GPU: SmoothL1 kernels, flat Adam, the (adj_pa . mask)^T
kernel, and the CUDA-graph-captured training step against the eager one."""
import numpy as np
import pytest
import torch

from helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu


def test_smooth_l1_matches_torch():
    from dstagnn_drought_b200.step import smooth_l1_loss
    dev = torch.device("cuda:0")
    for n in ((3, 5, 7), (64, 883, 12), (1,)):
        g = torch.Generator().manual_seed(0)
        pred = (torch.randn(*n, generator=g) * 1.5).to(dev).requires_grad_(True)
        tgt = torch.randn(*n, generator=g).to(dev)
        loss = smooth_l1_loss(pred, tgt)
        (loss * 3.0).backward()
        p2 = pred.detach().double().requires_grad_(True)
        ref = torch.nn.SmoothL1Loss()(p2, tgt.double())
        (ref * 3.0).backward()
        assert abs(loss.item() - ref.item()) < 1e-6 * max(1.0, abs(ref.item()))
        assert rel_err(pred.grad, p2.grad) < 1e-6
        assert torch.equal(smooth_l1_loss(pred, tgt), loss)              # reproducible reduction


def test_flat_adam_matches_torch_adam():
    from dstagnn_drought_b200.parallel import GradBuckets
    from dstagnn_drought_b200.step import FlatAdam, StepCounters
    dev = torch.device("cuda:0")

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(2)
            self.BlockList = torch.nn.ModuleList([torch.nn.Linear(37, 19), torch.nn.Linear(19, 5)])
            self.final_fc = torch.nn.Linear(5, 3)

        def forward(self, x):
            return self.final_fc(torch.tanh(self.BlockList[1](torch.tanh(self.BlockList[0](x)))))

    a, b = Net().to(dev), Net().to(dev)
    buckets = GradBuckets(a)
    counters = StepCounters(dev)
    try:
        opt_a = FlatAdam(a, buckets, 1e-2, counters=counters)
        opt_b = torch.optim.Adam(b.parameters(), lr=1e-2)
        x = torch.randn(11, 37, device=dev)
        for _ in range(25):
            buckets.zero()
            a(x).square().mean().backward()
            buckets.finish()
            opt_a.step()
            counters.tick()
            opt_b.zero_grad()
            b(x).square().mean().backward()
            opt_b.step()
        for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            assert rel_err(pa, pb) < 2e-5, k
        assert a.state_dict().keys() == b.state_dict().keys()
    finally:
        counters.close()


def test_bias_transpose_kernel_fwd_bwd():
    from dstagnn_drought_b200.ops import BiasT
    dev = torch.device("cuda:0")
    for N, K in ((5, 1), (33, 3), (307, 2)):
        ld = (N + 31) // 32 * 32
        pa = (torch.rand(N, N, device=dev) < 0.3).float()
        masks = [torch.randn(N, N, device=dev, requires_grad=True) for _ in range(K)]
        out = BiasT.apply(pa, ld, *masks)
        want = torch.stack([(pa * m).t() for m in masks])
        assert torch.equal(out[..., :N], want) and bool((out[..., N:] == 0).all())
        d = torch.randn(K, N, ld, device=dev)
        out.backward(d)
        for k in range(K):
            assert torch.equal(masks[k].grad, pa * d[k, :, :N].t())


def _net(dev, dropout):
    from dstagnn_drought_b200 import make_model
    d, cfg, w, g, cheb = load_golden("g_n33")
    net = make_model(dev, 1, cfg["nb"], 1, cfg["K"], 32, 32, 1, torch.from_numpy(d["adj_tmd"]),
                     torch.from_numpy(d["adj_pa"]), torch.from_numpy(d["adj_tmd"]), cfg["P"], cfg["T"],
                     cfg["N"], cfg["D"], 32, 32, cfg["H"])
    net.load_state_dict(w)
    net.train()
    if not dropout:
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    return net, torch.from_numpy(d["x"]).to(dev), torch.from_numpy(d["target"]).to(dev)


def test_graph_captured_step_equals_eager_step_and_torch_adam():
    """Without dropout the captured step, the eager step and a plain torch loop (nn.SmoothL1Loss + optim.Adam
    over the live parameters) walk the same trajectory."""
    from dstagnn_drought_b200.parallel import live_parameters
    from dstagnn_drought_b200.step import TrainStep
    dev = torch.device("cuda:0")
    losses = {}
    for mode in ("graph", "eager"):
        net, x, y = _net(dev, dropout=False)
        ts = TrainStep(net, 1e-3, graph=(mode == "graph"), warmup=2)
        try:
            losses[mode] = [float(ts.step(x, y)) for _ in range(8)]
            if mode == "graph":
                assert ts.replays == 6 and ts.kernels_per_replay > 50
        finally:
            ts.close()
    net, x, y = _net(dev, dropout=False)
    opt = torch.optim.Adam(list(live_parameters(net).values()), lr=1e-3)
    ref = []
    for _ in range(8):
        opt.zero_grad()
        loss = torch.nn.SmoothL1Loss()(net(x), y)
        loss.backward()
        opt.step()
        ref.append(float(loss))
    assert np.allclose(losses["graph"], losses["eager"], rtol=2e-5)
    assert np.allclose(losses["eager"], ref, rtol=2e-4)
    assert losses["eager"][-1] < losses["eager"][0]


def test_graph_replays_draw_fresh_dropout_masks():
    """The dropout key of a captured step comes from the device-resident step counter: identical inputs give
    different losses on consecutive replays (p = 0.05 masks differ), while re-running from the same counter value
    reproduces the loss bit for bit."""
    from dstagnn_drought_b200.step import TrainStep
    dev = torch.device("cuda:0")
    net, x, y = _net(dev, dropout=True)
    ts = TrainStep(net, 0.0, graph=True, warmup=1)                       # lr = 0: the weights stay put
    try:
        vals = [float(ts.step(x, y)) for _ in range(6)]
        assert len(set(vals[2:])) == len(vals[2:])                       # every replay has its own masks
        ts.counters.t[0] -= 1                                            # rewind the dropout counter by one step
        assert float(ts.step(x, y)) == vals[-1]
    finally:
        ts.close()
