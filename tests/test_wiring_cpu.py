"""Host-side wiring of the whole block stack, driven on CPU through the numpy emulation of the
C ABI (tests/cpu_emulation.py) and checked against the reference's golden fixtures: einsum
specs, ctypes argument order, hand-written backward formulas, the three gradient paths into
x and the res_att chain.  The real kernels are checked by the -m gpu tests."""
import pytest
import torch

import cpu_emulation
from dstagnn_drought_b200 import make_model
from helpers import GOLDEN, load_golden, rel_err


@pytest.mark.parametrize("path", ["dense", "sparse"])
@pytest.mark.parametrize("name", GOLDEN)
def test_block_stack_wiring_matches_reference(name, path, monkeypatch):
    """Both Chebyshev paths (dense N x N contraction / edge lists) against the reference fixtures."""
    monkeypatch.setenv("DSTAG_CHEB", path)
    fake = cpu_emulation.install(monkeypatch)
    d, cfg, w, g, cheb = load_golden(name)
    net = make_model("cpu", 1, cfg["nb"], 1, cfg["K"], 32, 32, 1, torch.from_numpy(d["adj_tmd"]),
                     torch.from_numpy(d["adj_pa"]), torch.from_numpy(d["adj_tmd"]), cfg["P"], cfg["T"],
                     cfg["N"], cfg["D"], 32, 32, cfg["H"])
    net.load_state_dict(w)
    net.eval()
    assert all(b.cheb_path == path for b in net.BlockList)
    blocks, atts = [], []

    def grab(_m, _i, o):
        blocks.append(o[0].detach())
        atts.append(o[1].detach())

    for b in net.BlockList:
        b.register_forward_hook(grab)
    y = net(torch.from_numpy(d["x"]))
    for i in range(cfg["nb"]):
        assert rel_err(blocks[i], d[f"blk64.{i}"]) < 2e-5, f"block {i}"
        assert rel_err(atts[i], d[f"att64.{i}"]) < 2e-5, f"logits {i}"
    assert rel_err(y.detach(), d["y64"]) < 2e-5
    torch.nn.SmoothL1Loss()(y, torch.from_numpy(d["target"])).backward()
    dead = set(str(k) for k in d["dead"])
    for k, p in net.named_parameters():
        if k in dead:
            assert p.grad is None, k
        else:
            assert p.grad is not None, k
            assert rel_err(p.grad, g[k]) < 5e-5, k
    assert fake.launches > 50


@pytest.mark.parametrize("T", [12, 140])
def test_gtu_stack_modes_against_conv2d(T, monkeypatch):
    """Host logic of the three GTU modes through the emulated ABI: fused forward + fused backward (T = 12), and the long
    series (T = 140: more than 128 output steps per node) whose forward runs fused and whose backward goes through the
    contraction engine from the saved gate factors (dstag_gtu_gate_bwd_sv)."""
    from dstagnn_drought_b200.ops import GtuStack
    fake = cpu_emulation.install(monkeypatch)
    calls = []
    for name in ("dstag_gtu_conv_gate_fwd", "dstag_gtu_gate_conv_bwd", "dstag_gtu_gate_bwd_sv", "dstag_gtu_gate_bwd"):
        orig = getattr(type(fake), name)
        monkeypatch.setattr(type(fake), name, (lambda o, n: lambda self, *a, **k: (calls.append(n), o(self, *a, **k))[1])(orig, name))
    torch.manual_seed(3)
    X, Cc = 2, 32
    g = torch.randn(X, T, Cc)
    ws = [torch.randn(2 * Cc, Cc, 1, ks) * 0.2 for ks in (3, 5, 7)]
    bs = [torch.randn(2 * Cc) for _ in range(3)]
    leaves = [t.clone().requires_grad_(True) for t in (g, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2])]
    out = GtuStack.apply(*leaves, True)
    dout = torch.randn_like(out)
    out.backward(dout)
    ref_leaves = [t.double().clone().requires_grad_(True) for t in (g, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2])]
    xc = ref_leaves[0].permute(2, 0, 1).unsqueeze(0)
    outs = []
    for i in range(3):
        u = torch.nn.functional.conv2d(xc, ref_leaves[1 + 2 * i], ref_leaves[2 + 2 * i])
        outs.append(torch.tanh(u[:, :Cc]) * torch.sigmoid(u[:, -Cc:]))
    ref = torch.cat(outs, dim=-1)[0].permute(1, 2, 0)
    ref.backward(dout.double())
    assert rel_err(out.detach(), ref.detach()) < 2e-5
    for a, b in zip(leaves, ref_leaves):
        assert rel_err(a.grad, b.grad) < 5e-5
    assert calls.count("dstag_gtu_conv_gate_fwd") == 3
    if T == 12:
        assert calls.count("dstag_gtu_gate_conv_bwd") == 3 and "dstag_gtu_gate_bwd_sv" not in calls
        assert "dstag_gtu_gate_bwd" not in calls
    else:       # (the emulation's _sv entry point forwards to its gate_bwd with the factors given)
        assert calls.count("dstag_gtu_gate_bwd_sv") == 3 and "dstag_gtu_gate_conv_bwd" not in calls


@pytest.mark.parametrize("T", [12, 140])
def test_gtu_stack_alias_output_accumulates_gradient(T, monkeypatch):
    """The alias output hands g on to a second consumer; the stack's backward accumulates into that consumer's
    gradient (in the kernels' dg buffer) and the sum equals what autograd would have added."""
    from dstagnn_drought_b200.ops import GtuStack
    cpu_emulation.install(monkeypatch)
    torch.manual_seed(8)
    X, Cc = 3, 32
    g = torch.randn(X, T, Cc)
    ws = [torch.randn(2 * Cc, Cc, 1, ks) * 0.2 for ks in (3, 5, 7)]
    bs = [torch.randn(2 * Cc) for _ in range(3)]
    mix = torch.randn(X, T, Cc)

    def run(alias):
        leaves = [t.clone().requires_grad_(True) for t in (g, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2])]
        gin = leaves[0] * 1.0                     # (a non-leaf, as in the block)
        if alias:
            cat, g2 = GtuStack.apply(gin, *leaves[1:], True, True)
        else:
            cat, g2 = GtuStack.apply(gin, *leaves[1:], True), gin
        loss = (cat ** 2).sum() + (g2 * mix).sum()
        loss.backward()
        return cat.detach(), [t.grad for t in leaves]

    cat_a, grads_a = run(True)
    cat_p, grads_p = run(False)
    assert torch.equal(cat_a, cat_p)
    for a, b in zip(grads_a, grads_p):
        assert rel_err(a, b) < 1e-6
    # only the alias used downstream: its gradient passes straight through
    leaf = g.clone().requires_grad_(True)
    _, g2 = GtuStack.apply(leaf * 1.0, *ws[:1], bs[0], ws[1], bs[1], ws[2], bs[2], True, True)
    (g2 * mix).sum().backward()
    assert torch.equal(leaf.grad, mix)
