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
