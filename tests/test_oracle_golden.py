"""Pin the CPU oracle against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import dstagnn_oracle as O
from helpers import GOLDEN, load_golden, rel_err

# fp64 oracle vs fp64 reference (fixture stored in fp32): only storage rounding remains
TOL64 = 5e-6


@pytest.mark.parametrize("name", GOLDEN)
@pytest.mark.parametrize("faithful", [True, False])
def test_oracle_matches_reference_fp64(name, faithful):
    d, cfg, w, g, cheb = load_golden(name)
    p = {k: v.double().requires_grad_(True) for k, v in w.items()}
    x = torch.from_numpy(d["x"]).double()
    y, blocks, atts = O.net_forward(p, x, torch.from_numpy(d["adj_pa"]).double(),
                                    [c.double() for c in cheb], cfg["nb"], cfg["H"], 32,
                                    faithful=faithful, return_blocks=True)
    assert rel_err(y, d["y64"]) < TOL64
    for i in range(cfg["nb"]):
        assert rel_err(blocks[i], d[f"blk64.{i}"]) < TOL64
        assert rel_err(atts[i], d[f"att64.{i}"]) < TOL64
    loss = torch.nn.SmoothL1Loss()(y, torch.from_numpy(d["target"]).double())
    assert abs(loss.item() - float(d["loss64"])) < 1e-9
    loss.backward()
    dead = set(str(k) for k in d["dead"])
    for k, v in p.items():
        if k in dead:
            assert v.grad is None, k
        else:
            assert rel_err(v.grad, g[k]) < TOL64, k


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_param_shapes_match_state_dict(name):
    d, cfg, w, g, cheb = load_golden(name)
    shapes = O.param_shapes(cfg["nb"], 1, cfg["K"], 32, cfg["N"], cfg["T"], cfg["D"], 32, cfg["H"], cfg["P"])
    assert [k for k, _ in shapes] == list(w.keys())
    for k, s in shapes:
        assert tuple(w[k].shape) == s, k


def test_cheb_polynomials_match_reference():
    for name in GOLDEN:
        d, cfg, w, g, cheb = load_golden(name)
        mine = O.cheb_polynomials(O.scaled_laplacian(d["adj_tmd"]), cfg["K"])
        for a, b in zip(mine, cheb):
            assert rel_err(torch.from_numpy(a), b) < 1e-4
