"""GPU tests of the tcgen05 contraction engine (csrc/gemm_tc.cu) against fp64 einsum:
both operand-load modes (K-vector / row-gather), fused Hadamard operand, epilogue variants,
split-K atomics, ragged tile edges, all three N-tile widths, both precisions."""
import ctypes as C

import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu

TOL = {"tf32x3": 2e-5, "tf32": 3e-3}


def _uses_tc(spec, A, B, out, precision, **kw):
    from dstagnn_drought_b200 import _lib, contraction as ct
    g = ct.build_descriptor(spec, A, B, out, precision=ct._PRECISION[precision], **kw)
    return bool(_lib.load().dstag_gemm_uses_tensor_cores(C.byref(g)))


def _run(spec, A, B, precision, **kw):
    from dstagnn_drought_b200 import contraction as ct
    shape = ct.output_shape(spec, A, B)
    out = kw.pop("out", None)
    if out is None:
        out = torch.full(shape, float("nan"), device=A.device)
    assert _uses_tc(spec, A, B, out, precision, **{k: v for k, v in kw.items() if k in ("a_mul", "c_add", "c_mul")})
    ct.contract_raw(spec, A, B, out=out, precision=ct._PRECISION[precision], **kw)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
@pytest.mark.parametrize("shape", [(1, 128, 128, 32), (1, 128, 128, 256), (2, 300, 200, 100), (1, 883, 384, 883),
                                   (3, 129, 40, 70), (1, 1000, 24, 37), (1, 64, 64, 2048), (5, 257, 130, 33)])
def test_tc_k_contiguous_operands(shape, precision):
    """A (b,m,k) and B (b,n,k): both K-contiguous -> float4 producer path (when aligned)."""
    b, m, n, k = shape
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    A = torch.randn(b, m, k, device=dev)
    B = torch.randn(b, n, k, device=dev)
    out = _run("bmk,bnk->bmn", A, B, precision)
    ref = torch.einsum("bmk,bnk->bmn", A.double(), B.double())
    assert rel_err(out, ref) < TOL[precision]


@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
@pytest.mark.parametrize("shape", [(2, 256, 128, 64), (1, 883, 384, 883), (2, 150, 72, 333)])
def test_tc_row_gather_operands(shape, precision):
    """A (b,k,m) and B (b,k,n): M/N-contiguous, K strided -> row-gather producer path."""
    b, m, n, k = shape
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    A = torch.randn(b, k, m, device=dev)
    B = torch.randn(b, k, n, device=dev)
    out = _run("bkm,bkn->bmn", A, B, precision)
    ref = torch.einsum("bkm,bkn->bmn", A.double(), B.double())
    assert rel_err(out, ref) < TOL[precision]


@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
def test_tc_cheb_shapes_with_hadamard_operand(precision):
    """The three Chebyshev contractions at a ragged N (padded maps, fused T_k, padded output)."""
    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    Bz, K, N, E, ld = 2, 3, 307, 96, 320
    P = torch.zeros(Bz, K, N, ld, device=dev)
    P[..., :N] = torch.softmax(torch.randn(Bz, K, N, N, device=dev), -1)
    T = torch.zeros(K, N, ld, device=dev)
    T[..., :N] = torch.randn(K, N, N, device=dev)
    X = torch.randn(Bz, N, E, device=dev)
    Pv, Tv = P[..., :N], T[..., :N]
    Z = _run("bkji,bie->kbje", Pv, X, precision, a_mul=("kji", Tv))
    refZ = torch.einsum("bkji,kji,bie->kbje", Pv.double(), Tv.double(), X.double())
    assert rel_err(Z, refZ) < TOL[precision]
    dZ = torch.randn_like(Z)
    dX = _run("bkji,kbje->bie", Pv, dZ, precision, a_mul=("kji", Tv))
    assert rel_err(dX, torch.einsum("bkji,kji,kbje->bie", Pv.double(), Tv.double(), dZ.double())) < TOL[precision]
    dS = torch.zeros(Bz, K, N, ld, device=dev)
    _run("kbje,bie->bkji", dZ, X, precision, out=dS[..., :N], c_mul=("kji", Tv))
    assert rel_err(dS[..., :N], torch.einsum("kbje,bie,kji->bkji", dZ.double(), X.double(), Tv.double())) < TOL[precision]
    assert float(dS[..., N:].abs().max()) == 0.0


@pytest.mark.parametrize("precision", ["tf32x3"])
def test_tc_epilogue_variants_and_split_k(precision):
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    X = torch.randn(20000, 96, device=dev)
    W = torch.randn(64, 96, device=dev)
    bias = torch.randn(64, device=dev)
    out = _run("xk,ok->xo", X, W, precision, c_add=("o", bias), relu=True, alpha=0.7)
    ref = torch.relu(0.7 * X.double() @ W.double().t() + bias.double())
    assert rel_err(out, ref) < TOL[precision]
    # weight gradient: tiny output, K = 20000 -> split-K with atomic accumulation
    dO = torch.randn(20000, 64, device=dev)
    dW = _run("xo,xk->ok", dO, X, precision, out=torch.empty(64, 96, device=dev))
    assert rel_err(dW, dO.double().t() @ X.double()) < TOL[precision]
    # accumulate into existing contents
    acc = torch.ones(20000, 64, device=dev)
    from dstagnn_drought_b200 import _lib
    _run("xk,ok->xo", X, W, precision, out=acc, accumulate=_lib.ACC_ADD, split_k=1)
    assert rel_err(acc, 1.0 + X.double() @ W.double().t()) < TOL[precision]


@pytest.mark.parametrize("precision", ["tf32x3"])
def test_tc_unfolded_conv_and_m_contiguous_output(precision):
    dev = torch.device("cuda:0")
    torch.manual_seed(4)
    Xn, Cc, T, ks = 3000, 32, 12, 5
    g3 = torch.randn(Xn, Cc, T, device=dev)
    w = torch.randn(2 * Cc, Cc, ks, device=dev) * 0.1
    b = torch.randn(2 * Cc, device=dev)
    u = _run("xctd,ocd->xot", g3.unfold(2, ks, 1), w, precision, c_add=("o", b))
    ref = torch.nn.functional.conv2d(g3.double().permute(1, 0, 2).unsqueeze(0), w.double().unsqueeze(2),
                                     b.double())[0].permute(1, 0, 2)
    assert rel_err(u, ref) < TOL[precision]
    du = torch.randn_like(u)
    from dstagnn_drought_b200 import _lib
    dg = torch.zeros_like(g3)
    _run("xot,ocd->xctd", du, w, precision, out=dg.unfold(2, ks, 1), accumulate=_lib.ACC_ATOMIC, split_k=1)
    g3r = g3.double().requires_grad_(True)
    torch.nn.functional.conv2d(g3r.permute(1, 0, 2).unsqueeze(0), w.double().unsqueeze(2),
                               b.double())[0].permute(1, 0, 2).backward(du.double())
    assert rel_err(dg, g3r.grad) < TOL[precision]
    dw = _run("xot,xctd->ocd", du, g3.unfold(2, ks, 1), precision, out=torch.empty(2 * Cc, Cc, ks, device=dev))
    ref_dw = torch.einsum("xot,xctd->ocd", du.double(), g3.double().unfold(2, ks, 1))
    assert rel_err(dw, ref_dw) < TOL[precision]


def test_tc_many_tiles_persistent_loop():
    """More tiles than SMs: every CTA walks several items through both TMEM accumulator stages."""
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    A = torch.randn(6, 1100, 200, device=dev)
    B = torch.randn(6, 520, 200, device=dev)
    out = _run("bmk,bnk->bmn", A, B, "tf32x3")
    assert rel_err(out, torch.einsum("bmk,bnk->bmn", A.double(), B.double())) < TOL["tf32x3"]


@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
@pytest.mark.parametrize("N", [130, 256, 883])
def test_tc_two_level_rows_split_to_batch(N, precision):
    """pre_conv's shape: A (b,f,t,n) has rows (b,n) with n contiguous (row padded to 16 bytes) -- the
    engine reruns it as a batched problem over b so the row-contiguous vector path applies; with a
    bias epilogue and an odd N."""
    from dstagnn_drought_b200 import contraction as ct
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    Bz, F, T, D = 3, 4, 12, 72
    A = ct.empty_rows((Bz, F, T, N), dev).normal_()
    W = torch.randn(D, T, F, device=dev)
    bias = torch.randn(D, device=dev)
    out = _run("bftn,dtf->bnd", A, W, precision, c_add=("d", bias))
    ref = torch.einsum("bftn,dtf->bnd", A.double(), W.double()) + bias.double()
    assert rel_err(out, ref) < TOL[precision]
    # and its weight gradient (K = (b,n) two-level) / input gradient (output rows padded)
    dZ = torch.randn(Bz, N, D, device=dev)
    dW = _run("bnd,bftn->dtf", dZ, A, precision)
    assert rel_err(dW, torch.einsum("bnd,bftn->dtf", dZ.double(), A.double())) < TOL[precision]
    dA = ct.empty_rows((Bz, F, T, N), dev)
    _run("dtf,bnd->bftn", W, dZ, precision, out=dA)
    assert rel_err(dA, torch.einsum("dtf,bnd->bftn", W.double(), dZ.double())) < TOL[precision]


@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
@pytest.mark.parametrize("shape", [(6, 70, 32, 70), (4, 200, 24, 130), (2, 300, 64, 96), (2, 129, 16, 257)])
def test_tc_row_contiguous_operands_narrow_n(shape, precision):
    """MN-major operands with the 32- and 64-wide N tiles (the spatial-attention dq/dk shapes):
    the row-quad producer has 2 (N tile 32) or 4 (N tile 64) k's per thread."""
    b, m, n, k = shape
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    A = torch.randn(b, k, m + 2, device=dev)[..., :m] if m % 4 == 2 else torch.randn(b, k, m, device=dev)
    B = torch.randn(b, k, n, device=dev)
    out = _run("bkm,bkn->bmn", A, B, precision)
    ref = torch.einsum("bkm,bkn->bmn", A.double(), B.double())
    assert rel_err(out, ref) < TOL[precision]


@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
@pytest.mark.parametrize("T", [12, 8, 6])
def test_tc_epilogue_aligned_runs_of_columns(T, precision):
    """Output columns n = (x, t) with t innermost and the m index in between (the theta-mix input
    gradient): runs of T contiguous floats.  T % 4 == 0 takes the float4 transposed epilogue, T = 6
    the scalar one; with an epilogue multiplier / accumulate in the same layout."""
    from dstagnn_drought_b200._lib import ACC_ADD
    dev = torch.device("cuda:0")
    torch.manual_seed(9)
    X, M, K = 301, 96, 32
    th = torch.randn(M, K, device=dev)
    gx = torch.randn(X, T, K, device=dev)
    ref = torch.einsum("mk,xtk->xmt", th.double(), gx.double())
    out = _run("mk,xtk->xmt", th, gx, precision)
    assert rel_err(out, ref) < TOL[precision]
    mul = torch.randn(X, M, T, device=dev)
    out2 = torch.ones(X, M, T, device=dev)
    _run("mk,xtk->xmt", th, gx, precision, out=out2, c_mul=("xmt", mul), accumulate=ACC_ADD)
    assert rel_err(out2, ref * mul.double() + 1.0) < TOL[precision]
    out3 = _run("mk,xtk->xmt", th, gx, precision, relu=True, alpha=0.5)
    assert rel_err(out3, torch.relu(0.5 * ref)) < TOL[precision]


@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
def test_tc_vector_loads_on_aligned_two_level_runs(precision):
    """Channels-last shapes: a reduction index (k, f) and a row index (x, c) that are contiguous only in
    runs (f / c innermost, run length % 4 == 0) still take the 16-byte copy paths."""
    dev = torch.device("cuda:0")
    torch.manual_seed(21)
    K_, X, T, F, C = 3, 333, 12, 32, 32
    Z = torch.randn(K_, X, T, F, device=dev)
    th = torch.randn(K_, F, C, device=dev)
    g = _run("kxtf,kfc->xtc", Z, th, precision, relu=True)                      # K = (k, f): runs of F
    assert rel_err(g, torch.relu(torch.einsum("kxtf,kfc->xtc", Z.double(), th.double()))) < TOL[precision]
    dZ = _run("xtc,kfc->kxtf", g, th, precision)                                # columns (k, f): runs of F
    assert rel_err(dZ, torch.einsum("xtc,kfc->kxtf", g.double(), th.double())) < TOL[precision]
    dth = _run("kxtf,xtc->kfc", Z, g, precision)                                # rows (k, f), K = (x, t)
    assert rel_err(dth, torch.einsum("kxtf,xtc->kfc", Z.double(), g.double())) < TOL[precision]
    S = 24
    cat = torch.randn(X, S, C, device=dev)
    W = torch.randn(T, S, device=dev)
    tc = _run("xsc,ts->xtc", cat, W, precision)                                 # rows (x, c): runs of C
    assert rel_err(tc, torch.einsum("xsc,ts->xtc", cat.double(), W.double())) < TOL[precision]


@pytest.mark.parametrize("case", [
    ("bmk,bnk->bmn", (2, 130, 70), (2, 45, 70), None),            # path (a)/(b): ragged M, N, K
    ("bkm,bkn->bmn", (3, 50, 257), (3, 50, 36), None),            # MN-major operands, ragged rows
    ("mk,xtk->xmt", (96, 32), (77, 12, 32), None),                # (b4) float4 transposed stores
    ("mk,xtk->xmt", (40, 24), (31, 6, 24), None),                 # scalar transposed stores
    ("xtk,ok->xto", (61, 10, 96), (64, 96), None),                # two-level rows
    ("bnd,bftn->dtf", (3, 131, 40), (3, 4, 6, 131), "rows"),      # K outer level folded into an atomic batch
])
def test_tc_writes_stay_inside_the_output(case):
    """Guard bands around the output (and a padded leading dimension inside it) must survive every
    epilogue path: the engine writes nothing outside the logical output elements."""
    from dstagnn_drought_b200 import contraction as ct
    spec, sa, sb, pad = case
    dev = torch.device("cuda:0")
    torch.manual_seed(33)
    A = torch.randn(*sa, device=dev)
    B = ct.empty_rows(sb, dev).normal_() if pad == "rows" else torch.randn(*sb, device=dev)
    shape = ct.output_shape(spec, A, B)
    n = 1
    for v in shape:
        n *= v
    ld = shape[-1] + 3                                            # padded rows: the pad must stay untouched too
    rows = n // shape[-1]
    G = 1024
    buf = torch.full((2 * G + rows * ld,), 12345.0, device=dev)
    out = buf[G:G + rows * ld].view(*shape[:-1], ld)[..., :shape[-1]]
    ct.contract_raw(spec, A, B, out=out, precision=ct._PRECISION["tf32x3"])
    torch.cuda.synchronize()
    ref = torch.einsum(spec, A.double(), B.double())
    assert rel_err(out, ref) < TOL["tf32x3"]
    assert bool((buf[:G] == 12345.0).all()) and bool((buf[G + rows * ld:] == 12345.0).all()), "guard band overwritten"
    padcols = buf[G:G + rows * ld].view(rows, ld)[:, shape[-1]:]
    assert bool((padcols == 12345.0).all()), "row padding overwritten"
