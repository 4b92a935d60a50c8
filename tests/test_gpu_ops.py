"""GPU unit tests: every C-ABI kernel against a plain PyTorch (fp64) evaluation of the same
formula, forward and backward, incl. ragged sizes.  Tolerances are relative to max|ref|."""
import math

import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu

TOL = 2e-5   # fp32 kernels vs fp64 evaluation


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(params=["fp32", "tf32x3"])
def precision(request):
    import dstagnn_drought_b200 as D
    old = D.get_precision()
    D.set_precision(request.param)
    yield request.param
    D.set_precision(old)


def _grads(fn, inputs, dout):
    xs = [t.detach().clone().requires_grad_(True) for t in inputs]
    out = fn(*xs)
    out.backward(dout.to(out.dtype))
    return out.detach(), [x.grad for x in xs]


@pytest.mark.parametrize("shape", [(1, 5, 7, 3), (3, 70, 65, 33), (2, 129, 64, 130), (1, 300, 33, 1000)])
def test_contract_matmul_fwd_bwd(dev, precision, shape):
    from dstagnn_drought_b200 import contract
    b, m, n, k = shape
    torch.manual_seed(0)
    A = torch.randn(b, m, k, device=dev)
    B = torch.randn(b, n, k, device=dev)
    bias = torch.randn(n, device=dev)
    dout = torch.randn(b, m, n, device=dev)
    out, (dA, dB, db) = _grads(lambda a, bb, c: contract("bmk,bnk->bmn", a, bb, c_add=("n", c), alpha=0.5),
                               [A, B, bias], dout)
    ref, (rA, rB, rb) = _grads(lambda a, bb, c: 0.5 * torch.einsum("bmk,bnk->bmn", a, bb) + c,
                               [A.double(), B.double(), bias.double()], dout.double())
    for got, want in ((out, ref), (dA, rA), (dB, rB), (db, rb)):
        assert rel_err(got, want) < TOL


def test_contract_weight_grad_split_k_and_relu(dev, precision):
    from dstagnn_drought_b200 import contract
    torch.manual_seed(1)
    X = torch.randn(9000, 48, device=dev)
    W = torch.randn(40, 48, device=dev)
    dout = torch.randn(9000, 40, device=dev)
    out, (dX, dW) = _grads(lambda x, w: contract("xk,ok->xo", x, w, relu=True), [X, W], dout)
    ref, (rX, rW) = _grads(lambda x, w: torch.relu(x @ w.t()), [X.double(), W.double()], dout.double())
    assert rel_err(out, ref) < TOL and rel_err(dX, rX) < TOL and rel_err(dW, rW) < TOL


def test_contract_permuted_and_two_level_axes(dev, precision):
    from dstagnn_drought_b200 import contract
    torch.manual_seed(2)
    Bz, F, T, N, Dm = 3, 4, 12, 37, 24
    O = torch.randn(Bz, F, T, N, device=dev)
    W = torch.randn(Dm, T, 1, F, device=dev)
    bias = torch.randn(Dm, device=dev)
    dout = torch.randn(Bz, N, Dm, device=dev)
    out, (dO, dW, db) = _grads(lambda o, w, c: contract("bftn,dtf->bnd", o, w[:, :, 0, :], c_add=("d", c)),
                               [O, W, bias], dout)
    ref, (rO, rW, rb) = _grads(lambda o, w, c: torch.einsum("bftn,dtf->bnd", o, w[:, :, 0, :]) + c,
                               [O.double(), W.double(), bias.double()], dout.double())
    for got, want in ((out, ref), (dO, rO), (dW, rW), (db, rb)):
        assert rel_err(got, want) < TOL


@pytest.mark.parametrize("rows,L,tab_mod", [(7, 5, 7), (24, 33, 12), (300, 883, 300), (1766, 512, 883)])
def test_layernorm_add(dev, rows, L, tab_mod):
    from dstagnn_drought_b200.ops import LayerNormAdd
    torch.manual_seed(3)
    a = torch.randn(rows, L, device=dev)
    tab = torch.randn(tab_mod, L, device=dev)
    gam, bet = torch.rand(L, device=dev), torch.rand(L, device=dev)
    dout = torch.randn(rows, L, device=dev)

    def ref_fn(a_, t_, g_, b_):
        pre = a_ + t_.repeat(rows // tab_mod, 1)
        return torch.nn.functional.layer_norm(pre, (L,), g_, b_, 1e-5)

    out, gr = _grads(lambda *xs: LayerNormAdd.apply(*xs, False, 0.0, 0), [a, tab, gam, bet], dout)
    ref, rr = _grads(ref_fn, [t.double() for t in (a, tab, gam, bet)], dout.double())
    assert rel_err(out, ref) < TOL
    for got, want in zip(gr, rr):
        assert rel_err(got, want) < 5e-5


def test_layernorm_gather_embedT(dev):
    from dstagnn_drought_b200.ops import LayerNormAdd
    torch.manual_seed(4)
    Bz, N, T = 3, 41, 12
    x3 = torch.randn(Bz, N, T, device=dev)
    tab = torch.randn(T, N, device=dev)
    gam, bet = torch.rand(N, device=dev), torch.rand(N, device=dev)
    dout = torch.randn(Bz * T, N, device=dev)

    def ref_fn(x_, t_, g_, b_):
        pre = x_.permute(0, 2, 1) + t_
        return torch.nn.functional.layer_norm(pre, (N,), g_, b_, 1e-5).reshape(Bz * T, N)

    out, gr = _grads(lambda *xs: LayerNormAdd.apply(*xs, True, 0.0, 0), [x3, tab, gam, bet], dout)
    ref, rr = _grads(ref_fn, [t.double() for t in (x3, tab, gam, bet)], dout.double())
    assert rel_err(out, ref) < TOL
    for got, want in zip(gr, rr):
        assert rel_err(got, want) < 5e-5


def test_layernorm_dropout_statistics_and_backward_mask(dev):
    from dstagnn_drought_b200.ops import LayerNormAdd
    torch.manual_seed(5)
    rows, L, p = 2000, 512, 0.05
    a = torch.randn(rows, L, device=dev, requires_grad=True)
    tab = torch.zeros(rows, L, device=dev)
    gam, bet = torch.ones(L, device=dev), torch.full((L,), 3.0, device=dev)
    y = LayerNormAdd.apply(a, tab, gam, bet, False, p, 1234)
    y0 = LayerNormAdd.apply(a, tab, gam, bet, False, 0.0, 0)
    dropped = (y == 0)
    frac = dropped.float().mean().item()
    assert abs(frac - p) < 0.003
    assert torch.allclose(y[~dropped], y0[~dropped] / (1 - p), rtol=1e-5, atol=1e-5)
    y2 = LayerNormAdd.apply(a, tab, gam, bet, False, p, 1234)
    assert torch.equal(y, y2)                                    # same seed -> same mask
    # backward uses the same mask: d/d(beta) = sum over kept rows of 1/(1-p)
    gb = torch.zeros(L, device=dev, requires_grad=True)
    yb = LayerNormAdd.apply(a.detach(), tab, gam, bet + gb, False, p, 1234)
    yb.sum().backward()
    want = (~dropped).float().sum(0) / (1 - p)
    assert torch.allclose(gb.grad, want, rtol=1e-4)


@pytest.mark.parametrize("cfg", [(2, 1, 2, 12, None), (2, 5, 3, 12, 1), (2, 5, 3, 12, 5), (1, 3, 2, 40, 3)])
def test_temporal_attention(dev, cfg):
    from dstagnn_drought_b200.ops import TemporalAttention
    Bz, F, H, T, res_F = cfg
    dk = 32
    torch.manual_seed(6)
    qkv = torch.randn(Bz, F, T, 3, H, dk, device=dev)
    res = torch.randn(Bz, res_F, H, T, T, device=dev) if res_F else None
    dctx = torch.randn(Bz, F, T, H * dk, device=dev)
    dlog = torch.randn(Bz, F, H, T, T, device=dev)

    def ref_fn(qkv_, res_):
        q, k, v = (qkv_[:, :, :, i].transpose(2, 3) for i in range(3))       # (B,F,H,T,dk)
        s = q @ k.transpose(-1, -2) / math.sqrt(dk)
        if res_ is not None:
            s = s + res_
        ctx = (torch.softmax(s, dim=3) @ v).transpose(2, 3).reshape(Bz, F, T, H * dk)
        return ctx, s

    xs = [qkv.clone().requires_grad_(True)] + ([res.clone().requires_grad_(True)] if res_F else [])
    ctx, log = TemporalAttention.apply(xs[0], xs[1] if res_F else None, Bz, F, H, T, dk)
    torch.autograd.backward([ctx, log], [dctx, dlog])
    rs = [qkv.double().requires_grad_(True)] + ([res.double().requires_grad_(True)] if res_F else [])
    rctx, rlog = ref_fn(rs[0], rs[1] if res_F else None)
    torch.autograd.backward([rctx, rlog], [dctx.double(), dlog.double()])
    assert rel_err(ctx, rctx) < TOL and rel_err(log, rlog) < TOL
    for a, b in zip(xs, rs):
        assert rel_err(a.grad, b.grad) < 5e-5


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "unfused"])
@pytest.mark.parametrize("cfg", [(2, 9, 3, 1), (2, 70, 3, 12), (1, 200, 2, 40), (1, 883, 2, 8), (1, 1100, 1, 4)])
def test_spatial_cheb_conv(dev, precision, cfg, monkeypatch, fused):
    """fused = the map kernel (logits + bias + softmax + T_k (.) P in one pass); unfused = the branch that large
    graphs take (N > ~2900: logits GEMM -> row softmax -> Hadamard operand fused into the contraction's
    load), forced here with DSTAG_SAT_FUSED=0 so that it is covered at sizes an fp64 check affords."""
    from dstagnn_drought_b200.ops import SpatialChebConv, _pad_ld
    if not fused:
        monkeypatch.setenv("DSTAG_SAT_FUSED", "0")
    Bz, N, K, FT = cfg
    dk = 32
    ld = _pad_ld(N)
    torch.manual_seed(7)
    qk = torch.randn(Bz, N, 2, K, dk, device=dev)
    bias = torch.zeros(K, N, ld, device=dev)
    bias[..., :N] = torch.randn(K, N, N, device=dev)
    tk = torch.zeros(K, N, ld, device=dev)
    tk[..., :N] = torch.randn(K, N, N, device=dev) * (torch.rand(K, N, N, device=dev) < 0.3)
    x3 = torch.randn(Bz, N, FT, device=dev)
    dZ = torch.randn(K, Bz, N, FT, device=dev)

    def ref_fn(qk_, bias_, x_):
        s = torch.einsum("bjkd,bikd->bkji", qk_[:, :, 1], qk_[:, :, 0]) / math.sqrt(dk) + bias_[..., :N]
        a = torch.softmax(s, dim=-1) * tk[..., :N].to(s.dtype)
        return torch.einsum("bkji,bie->kbje", a, x_)

    out, gr = _grads(lambda a, b, c: SpatialChebConv.apply(a, b, tk, c), [qk, bias, x3], dZ)
    ref, rr = _grads(ref_fn, [qk.double(), bias.double(), x3.double()], dZ.double())
    assert rel_err(out, ref) < TOL
    for got, want in ((gr[0], rr[0]), (gr[2], rr[2]), (gr[1][..., :N], rr[1][..., :N])):      # gradients: 5e-5 as everywhere
        assert rel_err(got, want) < 5e-5
    assert float(gr[1][..., N:].abs().max()) == 0.0 if ld > N else True


@pytest.mark.parametrize("cfg", [(1, 1, 1, 32), (2, 15, 3, 32), (2, 16, 1, 32), (3, 17, 2, 32), (2, 64, 3, 32),
                                 (1, 65, 2, 32), (2, 307, 3, 32), (1, 1021, 2, 32), (1, 1024, 1, 32),
                                 (2, 50, 2, 16), (1, 1030, 1, 32)])
@pytest.mark.parametrize("with_tp", [True, False])
def test_sat_softmax_fwd_kernel(dev, cfg, with_tp):
    """dstag_sat_softmax_fwd alone through the C ABI: the tensor-core variant (d_k = 32, ld <= 1024) at ragged
    strip / tile sizes, the SIMT rows kernel (d_k = 16) and the shared-memory strips (ld > 1024); P and
    T_k (.) P against an fp64 evaluation, zero padding columns, rows of P sum to 1.  Tolerance 2e-6 absolute
    (probabilities <= 1; 3xTF32 logits with K = 32)."""
    from dstagnn_drought_b200 import _lib
    Bz, N, K, dk = cfg
    ld = (N + 3) // 4 * 4
    lib = _lib.load()
    torch.manual_seed(11)
    qk = torch.randn(Bz, N, 2, K, dk, device=dev)
    bias = torch.zeros(K, N, ld, device=dev)
    bias[..., :N] = torch.randn(K, N, N, device=dev)
    tk = torch.zeros(K, N, ld, device=dev)
    tk[..., :N] = torch.randn(K, N, N, device=dev)
    P = torch.full((Bz, K, N, ld), float("nan"), device=dev)
    TP = torch.full((Bz, K, N, ld), float("nan"), device=dev) if with_tp else None
    _lib.check(lib.dstag_sat_softmax_fwd(qk.data_ptr(), bias.data_ptr(), tk.data_ptr(), P.data_ptr(),
                                         TP.data_ptr() if with_tp else None, Bz, N, K, dk, ld, _lib.stream_ptr(P)),
               "dstag_sat_softmax_fwd")
    s = torch.einsum("bjkd,bikd->bkji", qk[:, :, 1].double(), qk[:, :, 0].double()) / math.sqrt(dk) + bias[..., :N].double()
    ref = torch.softmax(s, dim=-1)
    assert float((P[..., :N].double() - ref).abs().max()) < 2e-6
    assert float((P[..., :N].double().sum(-1) - 1).abs().max()) < 1e-5
    assert bool((P[..., N:] == 0).all())
    if with_tp:
        want = ref * tk[..., :N].double()
        assert float((TP[..., :N].double() - want).abs().max()) < 2e-6 * max(1.0, float(tk.abs().max()))
        assert bool((TP[..., N:] == 0).all())


@pytest.mark.parametrize("cfg", [(2, 9, 3, 4), (2, 70, 3, 12), (3, 200, 2, 384), (1, 883, 3, 8), (1, 1100, 1, 640)])
def test_sparse_cheb_conv_matches_dense_formula(dev, precision, cfg):
    """sparse.SparseChebConv (edge lists: csrc/sparse.cu) against an fp64 evaluation of the dense formula of
    model/DSTAGNN_my.py:126-130, forward and every gradient, on sparse T_k (identity + ~3 % random edges) and a
    sparse adj_pa; ragged N, row lengths 4 ... 640 (registers-cached and streamed edge dots)."""
    from dstagnn_drought_b200.ops import _pad_ld
    from dstagnn_drought_b200.sparse import EdgeLists, SparseChebConv
    Bz, N, K, FT = cfg
    dk = 32
    ld = _pad_ld(N)
    g = torch.Generator().manual_seed(5)
    polys = [torch.eye(N)] + [torch.randn(N, N, generator=g) * (torch.rand(N, N, generator=g) < 0.03) + torch.eye(N)
                              for _ in range(K - 1)]
    pa = (torch.rand(N, N, generator=g) < 0.04).float()
    edges = {n: t.to(dev) for n, t in EdgeLists([p.numpy() for p in polys], pa.numpy()).tensors.items()}
    tk = torch.zeros(K, N, ld)
    for k in range(K):
        tk[k, :, :N] = polys[k].t()
    tk = tk.to(dev)
    qk = torch.randn(Bz, N, 2, K, dk, generator=g).to(dev)
    bias = torch.zeros(K, N, ld)
    bias[..., :N] = torch.randn(K, N, N, generator=g) * pa.t()            # (adj_pa . mask)^T: non-zero on adj_pa only
    bias = bias.to(dev)
    x3 = torch.randn(Bz, N, FT, generator=g).to(dev)
    dZ = torch.randn(K, Bz, N, FT, generator=g).to(dev)

    def ref_fn(qk_, bias_, x_):
        s = torch.einsum("bjkd,bikd->bkji", qk_[:, :, 1], qk_[:, :, 0]) / math.sqrt(dk) + bias_[..., :N]
        a = torch.softmax(s, dim=-1) * tk[..., :N].to(s.dtype)
        return torch.einsum("bkji,bie->kbje", a, x_)

    out, gr = _grads(lambda a, b, c: SparseChebConv.apply(a, b, tk, c, edges), [qk, bias, x3], dZ)
    ref, rr = _grads(ref_fn, [qk.double(), bias.double(), x3.double()], dZ.double())
    assert rel_err(out, ref) < TOL
    assert rel_err(gr[0], rr[0]) < 5e-5 and rel_err(gr[2], rr[2]) < 5e-5
    # d bias: the mask gradient keeps the adj_pa entries only (dmask = adj_pa . dbias^T) -- compare there
    keep = pa.t().to(dev)
    assert rel_err(gr[1][..., :N] * keep, rr[1][..., :N] * keep.double()) < 5e-5
    assert float((gr[1][..., :N] * (1 - keep)).abs().max()) == 0.0


@pytest.mark.parametrize("cfg", [(5, 32, 12), (300, 32, 12), (7, 32, 20), (6, 8, 144), (1531, 32, 12), (40, 32, 7),
                                 (9, 32, 130), (5, 32, 144)])   # (last: fused forward, engine backward from the gate factors)
def test_gtu_stack(dev, precision, cfg):
    from dstagnn_drought_b200.ops import GtuStack
    X, Cc, T = cfg
    torch.manual_seed(8)
    g = torch.randn(X, T, Cc, device=dev)                                     # channels-last
    ws = [torch.randn(2 * Cc, Cc, 1, ks, device=dev) * 0.2 for ks in (3, 5, 7)]
    bs = [torch.randn(2 * Cc, device=dev) for _ in range(3)]
    dout = torch.randn(X, 3 * T - 12, Cc, device=dev)                         # cat is channels-last too

    def ref_fn(g_, w3, b3, w5, b5, w7, b7):
        xc = g_.permute(2, 0, 1).unsqueeze(0)                                 # (1, C, X, T)
        outs = []
        for w, b in ((w3, b3), (w5, b5), (w7, b7)):
            u = torch.nn.functional.conv2d(xc, w, b)
            outs.append(torch.tanh(u[:, :Cc]) * torch.sigmoid(u[:, -Cc:]))
        return torch.cat(outs, dim=-1)[0].permute(1, 2, 0)                    # (C, X, S) -> (X, S, C)

    args = [g, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2]]
    out, gr = _grads(lambda *xs: GtuStack.apply(*xs), args, dout)
    ref, rr = _grads(ref_fn, [a.double() for a in args], dout.double())
    assert rel_err(out, ref) < TOL
    for got, want in zip(gr, rr):
        assert rel_err(got, want) < 5e-5


@pytest.mark.parametrize("cfg", [
    # (X, T, Tp, taps, planes, NB): GTU conv weight gradients (taps = consecutive time steps) ...
    (5, 12, 10, 3, False, 64), (301, 12, 8, 5, False, 64), (1531, 12, 6, 7, False, 64), (77, 20, 14, 7, False, 64),
    (40, 7, 1, 7, False, 64), (3, 12, 10, 3, False, 32),
    # ... and the Theta gradient (every tap a plane of its own, identity rows)
    (5000, 1, 1, 3, True, 32), (37, 1, 1, 2, True, 32), (129, 1, 1, 8, True, 64)])
def test_gram_rows_kernel(dev, cfg):
    """csrc/gram_tc.cu through the C ABI: D[(j, c), n] = sum_r A_j[r, c] B[r, n] against fp64, ragged row counts,
    both tap layouts, one and two M tiles; the per-CTA partials are added by dstag_sum_slices."""
    from dstagnn_drought_b200 import _lib
    lib = _lib.load()
    X, T, Tp, taps, planes, NB = cfg
    torch.manual_seed(21)
    R = X * Tp
    if planes:
        a = torch.randn(taps, R, 32, device=dev)
        tap_stride = R * 32
        Aj = [a[j].double() for j in range(taps)]
    else:
        a = torch.randn(X, T, 32, device=dev)
        tap_stride = 32
        Aj = [a[:, j:j + Tp, :].reshape(R, 32).double() for j in range(taps)]
    b = torch.randn(R, NB, device=dev)
    slices = lib.dstag_gram_slices()
    mrows = 128 * ((taps + 3) // 4)
    part = torch.full((slices, mrows, NB), float("nan"), device=dev)
    st = _lib.stream_ptr(a)
    _lib.check(lib.dstag_gram_rows(a.data_ptr(), b.data_ptr(), part.data_ptr(), X, T, Tp, taps, tap_stride, NB, st),
               "dstag_gram_rows")
    out = torch.empty((mrows, NB), device=dev)
    _lib.check(lib.dstag_sum_slices(part.data_ptr(), part.stride(0), slices, out.data_ptr(), mrows, NB, NB, st),
               "dstag_sum_slices")
    ref = torch.cat([A.t() @ b.double() for A in Aj], 0)                          # ((j, c), n)
    assert rel_err(out[:taps * 32], ref) < 2e-5
    assert float(out[taps * 32:].abs().max() if taps * 32 < mrows else 0.0) == 0.0   # rows of absent taps are zero
    # run-to-run reproducible
    part2 = torch.empty_like(part)
    _lib.check(lib.dstag_gram_rows(a.data_ptr(), b.data_ptr(), part2.data_ptr(), X, T, Tp, taps, tap_stride, NB, st),
               "dstag_gram_rows")
    assert torch.equal(part, part2)


@pytest.mark.parametrize("cfg", [(3, 5, 12, 1, 32), (3, 1531, 12, 1, 32), (1, 77, 3, 1, 8), (2, 40, 12, 2, 128), (4, 9, 7, 1, 64),
                                 (2, 301, 12, 1, 16)])
def test_thin_theta_mix(dev, monkeypatch, cfg):
    """Theta mix of the first block (K * F_in <= 4 values per row): streaming kernels vs fp64, and the same result as
    the contraction engine they replace (DSTAG_THIN=0)."""
    from dstagnn_drought_b200 import contract
    K, X, T, Fin, Cc = cfg
    torch.manual_seed(4)
    Z, th = torch.randn(K, X, T, Fin, device=dev), torch.randn(K, Fin, Cc, device=dev)
    dout = torch.randn(X, T, Cc, device=dev)
    fn = lambda z, t: contract("kxtf,kfc->xtc", z, t, relu=True)
    from dstagnn_drought_b200 import _lib
    before = _lib.launch_count()
    out, (dZ, dT) = _grads(fn, [Z, th], dout)
    assert _lib.launch_count() - before == 3            # expand; one backward pass; fixed-order fold of its partials
    ref, (rZ, rT) = _grads(lambda z, t: torch.relu(torch.einsum("kxtf,kfc->xtc", z, t)), [Z.double(), th.double()],
                           dout.double())
    for got, want in ((out, ref), (dZ, rZ), (dT, rT)):
        assert rel_err(got, want) < TOL
    monkeypatch.setenv("DSTAG_THIN", "0")
    out2, (dZ2, dT2) = _grads(fn, [Z, th], dout)
    for got, want in ((out2, ref), (dZ2, rZ), (dT2, rT)):
        assert rel_err(got, want) < 1e-4


@pytest.mark.parametrize("cfg", [(5, 12, 1, 32), (1531, 12, 1, 32), (64, 7, 3, 64), (33, 12, 4, 8)])
def test_thin_residual_conv(dev, cfg):
    """1x1 residual convolution of the first block: out = x w^T + b with w (C, F_in); x needs no gradient."""
    from dstagnn_drought_b200 import contract
    X, T, Fin, Cc = cfg
    torch.manual_seed(5)
    x, w, b = torch.randn(X, T, Fin, device=dev), torch.randn(Cc, Fin, device=dev), torch.randn(Cc, device=dev)
    dout = torch.randn(X, T, Cc, device=dev)
    out, (dx, dw, db) = _grads(lambda x_, w_, b_: contract("xtf,cf->xtc", x_, w_, c_add=("c", b_)), [x, w, b], dout)
    ref, (rx, rw, rb) = _grads(lambda x_, w_, b_: torch.einsum("xtf,cf->xtc", x_, w_) + b_,
                               [x.double(), w.double(), b.double()], dout.double())
    for got, want in ((out, ref), (dx, rx), (dw, rw), (db, rb)):
        assert got.shape == want.shape and rel_err(got, want) < TOL
    wd = w.detach().clone().requires_grad_(True)
    contract("xtf,cf->xtc", x, wd).backward(dout)          # weight gradient alone (no dz, no bias)
    assert rel_err(wd.grad, rw) < TOL


@pytest.mark.parametrize("cfg", [(3, 12, 36, 32), (1531, 12, 36, 32), (37, 20, 60, 32), (64, 3, 4, 8), (501, 6, 12, 16),
                                 (9, 7, 24, 64)])
def test_fcmy_param_grads(dev, cfg):
    """Streaming T x S reduction (one and three rows of `a` per thread, ragged node slabs) against fp64 einsum."""
    from dstagnn_drought_b200.ops import fcmy_param_grads
    X, T, S, Cc = cfg
    torch.manual_seed(21)
    dtc, cat = torch.randn(X, T, Cc, device=dev), torch.randn(X, S, Cc, device=dev)
    dw, db = fcmy_param_grads(dtc, cat)
    assert rel_err(dw, torch.einsum("xtc,xsc->ts", dtc.double(), cat.double())) < TOL
    assert rel_err(db, dtc.double().sum((0, 2))) < TOL


@pytest.mark.parametrize("cfg", [(148, 128, 32), (64, 512, 384), (17, 5, 12), (3, 7, 6), (295, 96, 32), (20, 1000, 64)])
def test_sum_slices_fixed_order(dev, cfg):
    """dstag_sum_slices (serial and slice-parallel kernels): the sum of the slices, identical from run to run."""
    from dstagnn_drought_b200 import _lib
    lib = _lib.load()
    ns, rows, cols = cfg
    torch.manual_seed(5)
    ws = torch.randn(ns, rows, cols, device=dev)
    out = torch.full((rows, cols), float("nan"), device=dev)
    st = _lib.stream_ptr(ws)
    _lib.check(lib.dstag_sum_slices(ws.data_ptr(), ws.stride(0), ns, out.data_ptr(), rows, cols, cols, st), "dstag_sum_slices")
    assert rel_err(out, ws.double().sum(0)) < 2e-6
    out2 = torch.empty_like(out)
    _lib.check(lib.dstag_sum_slices(ws.data_ptr(), ws.stride(0), ns, out2.data_ptr(), rows, cols, cols, st), "dstag_sum_slices")
    assert torch.equal(out, out2)


@pytest.mark.parametrize("cfg", [(257, 32, 12), (11, 32, 144)])
def test_gtu_fused_forward_without_grad_matches_gemm_path(dev, cfg, monkeypatch):
    """Inference (no saved gate factors; any T, also the T' > 128 rows per node of the drought configuration)
    through the fused tcgen05 kernel against the contraction-engine + gate-kernel path."""
    from dstagnn_drought_b200.ops import GtuStack
    X, Cc, T = cfg
    torch.manual_seed(18)
    g = torch.randn(X, T, Cc, device=dev)
    ws = [torch.randn(2 * Cc, Cc, 1, ks, device=dev) * 0.2 for ks in (3, 5, 7)]
    bs = [torch.randn(2 * Cc, device=dev) for _ in range(3)]
    args = [g, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2]]
    with torch.no_grad():
        fused = GtuStack.apply(*args, False)
        monkeypatch.setenv("DSTAG_GTU", "gemm")
        plain = GtuStack.apply(*args, False)
    xc = g.double().permute(2, 0, 1).unsqueeze(0)
    outs = []
    for w, b in zip(ws, bs):
        u = torch.nn.functional.conv2d(xc, w.double(), b.double())
        outs.append(torch.tanh(u[:, :Cc]) * torch.sigmoid(u[:, -Cc:]))
    ref = torch.cat(outs, dim=-1)[0].permute(1, 2, 0)
    assert rel_err(fused, ref) < TOL and rel_err(plain, ref) < TOL


@pytest.mark.parametrize("cfg", [(3, 32, 12, True), (500, 32, 12, False), (37, 32, 12, True), (9, 32, 144, True),
                                 (5, 16, 20, False)])
def test_block_tail(dev, cfg):
    from dstagnn_drought_b200.ops import BlockTail
    X, Cc, T, with_g = cfg
    torch.manual_seed(9)
    tc, r, g = (torch.randn(X, T, Cc, device=dev) for _ in range(3))          # all channels-last
    gam, bet = torch.rand(Cc, device=dev), torch.rand(Cc, device=dev)
    dout = torch.randn(X, T, Cc, device=dev)

    def ref_fn(tc_, g_, r_, ga_, be_):
        tco = torch.relu(tc_ + g_) if with_g else torch.relu(tc_)
        return torch.nn.functional.layer_norm(torch.relu(r_ + tco), (Cc,), ga_, be_, 1e-5)

    def fn(tc_, g_, r_, ga_, be_):
        return BlockTail.apply(tc_, g_ if with_g else None, r_, ga_, be_, 0.0, 0)

    out, gr = _grads(fn, [tc, g, r, gam, bet], dout)
    ref, rr = _grads(ref_fn, [t.double() for t in (tc, g, r, gam, bet)], dout.double())
    assert rel_err(out, ref) < TOL
    for i, (got, want) in enumerate(zip(gr, rr)):
        if i == 1 and not with_g:
            continue
        assert rel_err(got, want) < 5e-5, i


def test_block_tail_dropout_is_consistent(dev):
    from dstagnn_drought_b200.ops import BlockTail
    torch.manual_seed(10)
    X, Cc, T, p = 4000, 32, 12, 0.05
    tc = (torch.rand(X, T, Cc, device=dev) + 0.5).requires_grad_(True)      # strictly positive
    r = torch.full((X, T, Cc), -100.0, device=dev)                          # second relu sees tco only
    gam, bet = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
    out = BlockTail.apply(tc, None, r + 100.0, gam, bet, p, 77)
    out.sum().backward()
    # LN output sums to zero per column, so instead check the mask through the gradient pattern:
    out2 = BlockTail.apply(tc, None, r + 100.0, gam, bet, p, 77)
    assert torch.equal(out, out2)
    out3 = BlockTail.apply(tc, None, r + 100.0, gam, bet, p, 78)
    assert not torch.equal(out, out3)


def test_transpose_softmax_colsum_relu(dev):
    import ctypes as C
    from dstagnn_drought_b200 import _lib
    from dstagnn_drought_b200.contraction import colsum, relu_bwd
    from dstagnn_drought_b200.ops import transpose_raw
    lib = _lib.load()
    torch.manual_seed(11)
    x = torch.randn(3, 45, 70, device=dev)
    assert torch.equal(transpose_raw(x), x.transpose(1, 2).contiguous())
    # index-pair swaps folded into the transpose: (b, n, (t,f)) -> (b, (f,t), n) and back
    Bz, N_, T_, F_ = 2, 45, 6, 5
    xc = torch.randn(Bz, N_, T_ * F_, device=dev)
    fwd = transpose_raw(xc, pad_rows=True, c_d=F_)
    want = xc.view(Bz, N_, T_, F_).permute(0, 3, 2, 1).reshape(Bz, F_ * T_, N_)
    assert torch.equal(fwd, want)
    assert torch.equal(transpose_raw(fwd, r_d=T_), xc)
    # float4 tile kernel (leading dimensions multiples of 4), both directions, ragged tiles, PEMS07's own shape
    for Bz, N_, T_, F_ in ((2, 45, 4, 8), (3, 883, 12, 32), (1, 130, 17, 4), (2, 64, 1, 64)):
        xc = torch.randn(Bz, N_, T_ * F_, device=dev)
        fwd = transpose_raw(xc, pad_rows=True, c_d=F_)
        assert fwd.stride(1) % 4 == 0
        want = xc.view(Bz, N_, T_, F_).permute(0, 3, 2, 1).reshape(Bz, F_ * T_, N_)
        assert torch.equal(fwd, want)
        assert torch.equal(transpose_raw(fwd, r_d=T_), xc)
        assert torch.equal(transpose_raw(xc), xc.transpose(1, 2).contiguous())
        assert torch.equal(transpose_raw(xc, pad_rows=True), xc.transpose(1, 2))
    for rows, cols in ((7, 1), (1000, 64), (64, 5000), (33, 130)):
        m = torch.randn(rows, cols, device=dev)
        assert rel_err(colsum(m, rows, cols), m.double().sum(0)) < TOL
    S = torch.randn(50, 40, device=dev)
    P = S.clone()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.dstag_softmax_fwd(P.data_ptr(), 50, 33, 40, st), "softmax")
    assert rel_err(P[:, :33], torch.softmax(S[:, :33].double(), -1)) < TOL
    assert float(P[:, 33:].abs().max()) == 0.0
    y, dy = torch.randn(1001, device=dev), torch.randn(1001, device=dev)
    assert torch.equal(relu_bwd(y, dy), dy * (y > 0))


@pytest.mark.parametrize("cfg", [(3, True), (500, False), (37, True)])
def test_fused_tail(dev, precision, cfg):
    """fcmy + tail in one kernel (C = 32, T = 12) against the unfused formula."""
    from dstagnn_drought_b200.ops import FusedTail
    X, with_g = cfg
    Cc, T, S = 32, 12, 24
    assert FusedTail.supported(Cc, T) and not FusedTail.supported(32, 18)
    torch.manual_seed(12)
    cat = torch.randn(X, S, Cc, device=dev)                                   # everything channels-last
    wf, bf = torch.randn(T, S, device=dev) * 0.3, torch.rand(T, device=dev)
    g, r = torch.randn(X, T, Cc, device=dev), torch.randn(X, T, Cc, device=dev)
    gam, bet = torch.rand(Cc, device=dev), torch.rand(Cc, device=dev)
    dout = torch.randn(X, T, Cc, device=dev)

    def ref_fn(cat_, wf_, bf_, g_, r_, ga_, be_):
        tc = torch.einsum("xsc,ts->xtc", cat_, wf_) + bf_[None, :, None]
        tco = torch.relu(tc + g_) if with_g else torch.relu(tc)
        return torch.nn.functional.layer_norm(torch.relu(r_ + tco), (Cc,), ga_, be_, 1e-5)

    def fn(cat_, wf_, bf_, g_, r_, ga_, be_):
        return FusedTail.apply(cat_, wf_, bf_, g_ if with_g else None, r_, ga_, be_, 0.0, 0)

    args = [cat, wf, bf, g, r, gam, bet]
    out, gr = _grads(fn, args, dout)
    ref, rr = _grads(ref_fn, [a.double() for a in args], dout.double())
    assert rel_err(out, ref) < TOL
    for i, (got, want) in enumerate(zip(gr, rr)):
        if i == 3 and not with_g:
            continue
        assert rel_err(got, want) < 5e-5, i


def test_fused_tail_dropout_matches_unfused_mask(dev):
    """Same seed => the fused kernel drops exactly the elements the unfused tail kernel drops."""
    from dstagnn_drought_b200.ops import BlockTail, FusedTail
    torch.manual_seed(13)
    X, Cc, T, S, p = 300, 32, 12, 24, 0.05
    cat = torch.randn(X, S, Cc, device=dev)
    wf, bf = torch.randn(T, S, device=dev) * 0.3, torch.rand(T, device=dev)
    r = torch.randn(X, T, Cc, device=dev)
    gam, bet = torch.rand(Cc, device=dev), torch.rand(Cc, device=dev)
    tc = (torch.einsum("xsc,ts->xtc", cat, wf) + bf[None, :, None]).contiguous()
    a = FusedTail.apply(cat, wf, bf, None, r, gam, bet, p, 4242)
    b = BlockTail.apply(tc, None, r, gam, bet, p, 4242)
    assert rel_err(a, b) < 1e-5
