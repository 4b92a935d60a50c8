"""GPU parity tests proper: the CUDA path (through the C ABI) against the committed golden
fixtures of the reference and against the CPU oracle, whole-net forward + backward.

Tolerance (north star: fp32, max rel err <= 1e-4): per tensor, max|err| / max|ref| against the
fp64 result; predictions and block outputs must be < 1e-4; gradients < 1e-4 or 3x the
reference's own fp32 error on that tensor, whichever is larger (the tiny
spatial-path gradients of the reference itself are only good to 3-5e-4)."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL_PRED = 1e-4
TOL_GRAD = 1e-4


def _build(cfg, d, dev):
    from dstagnn_drought_b200 import make_model
    return make_model(dev, 1, cfg["nb"], 1, cfg["K"], 32, 32, 1, torch.from_numpy(d["adj_tmd"]),
                      torch.from_numpy(d["adj_pa"]), torch.from_numpy(d["adj_tmd"]), cfg["P"], cfg["T"],
                      cfg["N"], cfg["D"], 32, 32, cfg["H"])


@pytest.mark.parametrize("path", ["dense", "sparse"])
@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("name", GOLDEN)
def test_net_matches_reference_golden(name, precision, path, monkeypatch):
    """Both Chebyshev paths -- the dense tensor-core contraction of the N x N maps and the edge-list path that
    sparse graphs take by default -- against the reference's own outputs and gradients."""
    import dstagnn_drought_b200 as D
    monkeypatch.setenv("DSTAG_CHEB", path)
    D.set_precision(precision)
    dev = torch.device("cuda:0")
    d, cfg, w, g, cheb = load_golden(name)
    net = _build(cfg, d, dev)
    net.load_state_dict(w)
    net.eval()
    blocks, atts = [], []
    for b in net.BlockList:
        b.register_forward_hook(lambda m, i, o: (blocks.append(o[0].detach()), atts.append(o[1].detach()), None)[2])
    x = torch.from_numpy(d["x"]).to(dev)
    y = net(x)
    assert y.shape == d["y64"].shape
    assert rel_err(y.detach(), d["y64"]) < TOL_PRED
    for i in range(cfg["nb"]):
        assert rel_err(blocks[i], d[f"blk64.{i}"]) < TOL_PRED, f"block {i} output"
        assert rel_err(atts[i], d[f"att64.{i}"]) < TOL_PRED, f"block {i} logits"
    loss = torch.nn.SmoothL1Loss()(y, torch.from_numpy(d["target"]).to(dev))
    assert abs(loss.item() - float(d["loss64"])) < 1e-5
    loss.backward()
    ref_err = dict(zip((str(k) for k in d["ref_fp32_grad_err_keys"]), d["ref_fp32_grad_err"]))
    dead = set(str(k) for k in d["dead"])
    worst = {}
    for k, p in net.named_parameters():
        if k in dead:
            assert p.grad is None, f"{k} must stay a dead parameter (grad None)"
            continue
        assert p.grad is not None, k
        e = rel_err(p.grad, g[k])
        worst[k] = e
        assert e < max(TOL_GRAD, 3 * ref_err[k]), f"{k}: {e:.3e} (reference fp32 itself: {ref_err[k]:.3e})"
    print(name, precision, path, "pred", rel_err(y.detach(), d["y64"]), "worst grad", max(worst.values()))
    D.set_precision("tf32x3")


# Oracle-only shapes: (shape, data seed, conditioned).
# conditioned = True: the seed was found offline by tests/tools/find_conditioned_seed.py using ONLY the fp64 oracle
#   (tests/helpers.conditioned_case: no ReLU pre-activation within ~1e-6 of its kink); the test re-verifies
#   that before it trusts the seed, and holds every gradient to max(1e-4, 3 x fp32-oracle error).
# conditioned = False: shapes with so many ReLU units (8 M at PEMS07, 59 M at GAMBIA) that no input avoids the
#   kinks; the tolerance of a gradient tensor additionally admits twice what flipping the units within
#   1e-6 * max|pre-activation| of a kink does to it in the fp64 oracle (tests/helpers.kink_slack).  Forward
#   quantities (predictions, block outputs, attention logits) are continuous and held to 1e-4 everywhere.
# Nothing is redrawn: the CUDA forward is run-to-run reproducible (no atomics), one run decides.
ORACLE_CASES = {
    # ragged PEMS08 / PEMS04 node counts at reduced width, a longer horizon, the GAMBIA layout at small N
    "n170_d64": (dict(N=170, T=12, K=3, H=3, D=64, nb=2, B=3), 4, True),
    "n307_d128": (dict(N=307, T=12, K=3, H=3, D=128, nb=2, B=2), 3, True),
    "n65_t24": (dict(N=65, T=24, K=2, H=2, D=64, nb=2, B=2), 3, True),
    "n37_t144": (dict(N=37, T=144, K=2, H=2, D=64, nb=2, B=2), 5, True),
    # the BASELINE configurations at their REAL shapes, batch reduced to what the fp64 CPU
    # oracle finishes in seconds
    "pems08_full": (dict(N=170, T=12, K=3, H=3, D=512, nb=4, B=2), 12, True),
    "pems04_full": (dict(N=307, T=12, K=3, H=3, D=512, nb=4, B=2), 21, True),
    "pems03_full": (dict(N=358, T=12, K=3, H=3, D=512, nb=4, B=2), 4, False),
    "pems07_full": (dict(N=883, T=12, K=3, H=4, D=512, nb=4, B=2), 4, False),
    # GAMBIA drought conf (in_channels = 1): N = 2139 > 1024 takes the shared-memory strip map
    # kernel, the block-per-row LayerNorm / softmax-backward kernels and the unfused tail (T = 144)
    "gambia_full": (dict(N=2139, T=144, K=2, H=2, D=64, nb=2, B=1), 4, False),
    # N = 4096: a 16-row logit strip does not fit in shared memory -> logits GEMM + row softmax + Hadamard
    # operand (the DSTAG_ERR_SMEM branch of ops.SpatialChebConv)
    "syn4096_nb2": (dict(N=4096, T=12, K=3, H=4, D=512, nb=2, B=1), 4, False),
}
COND_THR = 5e-5
KINK_TAU = 1e-6


# every case on the path its graph selects (the 1 %-dense synthetic graphs take the edge lists), plus the dense
# tensor-core contraction forced at the shapes that exercise its variants (TMA-fed maps at N = 883 / 307, the
# shared-memory strip map kernel at N = 2139, the GEMM + row softmax + Hadamard-operand branch at N = 4096)
CASE_PATHS = [(c, "auto") for c in sorted(ORACLE_CASES)] + [(c, "dense") for c in
                                                             ("n307_d128", "pems07_full", "gambia_full", "syn4096_nb2")]


@pytest.mark.parametrize("case,path", CASE_PATHS)
def test_net_matches_oracle_at_config_shapes(case, path, monkeypatch):
    """Whole-net forward + backward against the fp64 CPU oracle: predictions, every block output, every
    attention-logit tensor < 1e-4; every gradient < max(1e-4, 3 x the error the same algorithm makes in plain
    fp32 on the CPU) [+ the kink slack at unconditioned shapes].  No retries."""
    monkeypatch.setenv("DSTAG_CHEB", path)
    from oracle import dstagnn_oracle as O
    from helpers import _oracle_run, conditioned_case, kink_slack, oracle_fp32_grad_err
    shape, seed, conditioned = ORACLE_CASES[case]
    dev = torch.device("cuda:0")
    N, T, K, H, Dm, nb, B = (shape[k] for k in ("N", "T", "K", "H", "D", "nb", "B"))
    onet = O.build_oracle(N, T, K, H, Dm, nb, dtype=torch.float64, faithful=False)
    if conditioned:
        x, tgt, y_ref, ograds, _ = conditioned_case(onet, B, N, T, seeds=(seed,), thr=COND_THR)
        slack = {k: 0.0 for k in ograds}
    else:
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(B, N, 1, T, generator=g)
        tgt = torch.randn(B, N, 12, generator=g)
        y_ref, ograds = _oracle_run(onet, x.double(), tgt.double())
        slack = kink_slack(onet, x, tgt, ograds, KINK_TAU)
    fp32_err = oracle_fp32_grad_err(onet, x, tgt, ograds)
    with torch.no_grad():
        _, oblocks, oatts = O.net_forward(onet.named(), x.double(), onet.adj_pa, onet.cheb, nb, H, 32,
                                          faithful=False, return_blocks=True)
    tmd, pa = O.synthetic_graph(N, 0)
    from dstagnn_drought_b200 import make_model
    net = make_model(dev, 1, nb, 1, K, 32, 32, 1, torch.from_numpy(tmd), torch.from_numpy(pa),
                     torch.from_numpy(tmd), 12, T, N, Dm, 32, 32, H)
    net.load_state_dict({k: v.detach().float() for k, v in onet.named().items()})
    net.eval()
    assert all(b.cheb_path == ("sparse" if path == "auto" else path) for b in net.BlockList)
    blocks, atts = [], []
    for b in net.BlockList:
        b.register_forward_hook(lambda m, i, o: (blocks.append(o[0].detach()), atts.append(o[1].detach()), None)[2])
    y = net(x.to(dev))
    torch.nn.SmoothL1Loss()(y, tgt.to(dev)).backward()
    assert rel_err(y.detach(), y_ref) < TOL_PRED
    for i in range(nb):
        assert rel_err(blocks[i], oblocks[i]) < TOL_PRED, f"block {i} output"
        assert rel_err(atts[i], oatts[i]) < TOL_PRED, f"block {i} logits"
    bad, worst, strict = [], 0.0, 0
    for k, p in net.named_parameters():
        if ograds[k] is None:
            assert p.grad is None, k
            continue
        e = rel_err(p.grad, ograds[k])
        worst = max(worst, e)
        tol = max(TOL_GRAD, 3 * fp32_err[k], 2 * slack[k])
        strict += tol == TOL_GRAD
        if not e < tol:
            bad.append(f"{k}: {e:.3e} (fp32 oracle itself: {fp32_err[k]:.3e}, kink slack {slack[k]:.3e})")
    print(case, path, "data seed", seed, "pred", rel_err(y.detach(), y_ref), "worst grad", worst,
          f"; {strict} of {sum(v is not None for v in ograds.values())} gradient tensors held to {TOL_GRAD}")
    assert not bad, "; ".join(bad)


def test_forward_is_run_to_run_reproducible():
    """Every forward reduction has a fixed order (split-K partials go through a workspace and
    dstag_sum_slices, never atomics): repeated forwards are bit-identical, also at the T = 144 shape whose
    pre_conv / final_conv products split their reduction."""
    from oracle import dstagnn_oracle as O
    from dstagnn_drought_b200 import make_model
    dev = torch.device("cuda:0")
    for (N, T, K, H, Dm, nb, B) in ((37, 144, 2, 2, 64, 2, 2), (307, 12, 3, 3, 128, 2, 2)):
        tmd, pa = O.synthetic_graph(N, 0)
        torch.manual_seed(1)
        net = make_model(dev, 1, nb, 1, K, 32, 32, 1, torch.from_numpy(tmd), torch.from_numpy(pa),
                         torch.from_numpy(tmd), 12, T, N, Dm, 32, 32, H)
        net.eval()
        x = torch.randn(B, N, 1, T, generator=torch.Generator().manual_seed(5)).to(dev)
        with torch.no_grad():
            ys = [net(x).clone() for _ in range(5)]
        assert all(torch.equal(ys[0], yy) for yy in ys[1:])


def test_train_mode_runs_and_dropout_is_active():
    dev = torch.device("cuda:0")
    d, cfg, w, g, cheb = load_golden("g_n20")
    net = _build(cfg, d, dev)
    net.load_state_dict(w)
    x = torch.from_numpy(d["x"]).to(dev)
    net.train()
    torch.manual_seed(0)
    y1 = net(x)
    y2 = net(x)
    net.eval()
    y0 = net(x)
    assert not torch.equal(y1, y2)                       # fresh masks per call
    assert rel_err(y1.detach(), y0.detach()) < 0.5       # but the same function in expectation
    y1.sum().backward()
    assert all(torch.isfinite(p.grad).all() for p in net.parameters() if p.grad is not None)


def test_inference_no_grad_and_block_level_api():
    """eval / no_grad inference gives the same predictions as the autograd forward; a block
    called on its own (the reference's `block(x, res_att)` contract) accepts a plain contiguous
    (B, N, F, T) input as well as the channels-last view a preceding block returns, and returns a
    (B, N, C, T) tensor plus the (B, F, H, T, T) logits."""
    dev = torch.device("cuda:0")
    d, cfg, w, g, cheb = load_golden("g_n20")
    net = _build(cfg, d, dev)
    net.load_state_dict(w)
    net.eval()
    x = torch.from_numpy(d["x"]).to(dev)
    y_grad = net(x)
    with torch.no_grad():
        y_inf = net(x)
    assert torch.equal(y_grad.detach(), y_inf)
    assert rel_err(y_inf, d["y64"]) < TOL_PRED
    with torch.no_grad():
        out0, att0 = net.BlockList[0](x, 0)
        B, N, C, T = out0.shape
        assert (B, N, T) == (x.shape[0], cfg["N"], cfg["T"]) and C == 32
        assert att0.shape == (B, 1, cfg["H"], T, T)
        out1_view, _ = net.BlockList[1](out0, att0)                       # channels-last view from block 0
        out1_copy, _ = net.BlockList[1](out0.contiguous(), att0)          # the same values, standard layout
    assert torch.equal(out1_view, out1_copy)
    assert rel_err(out1_view, d["blk64.1"]) < TOL_PRED
