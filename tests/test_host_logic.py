"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol,
the einsum front end builds descriptors that reproduce torch.einsum under an emulator of the
kernel's addressing rules, and the module API matches the reference's state_dict / init."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from dstagnn_drought_b200 import _lib, contraction as ct
from dstagnn_drought_b200 import make_model
from helpers import GOLDEN, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "dstagnn_b200.h")).read()
    names = set(re.findall(r"\b(dstag_[a-z0-9_]+)\s*\(", hdr))
    assert names, "no prototypes found"
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.dstag_abi_version() == _lib.ABI_VERSION


def test_gemm_rejects_bad_descriptor_without_gpu():
    lib = _lib.load()
    g = _lib.Gemm()
    assert lib.dstag_gemm(C.byref(g), None) == -1          # DSTAG_ERR_BAD_ARG, nothing launched
    assert b"empty problem" in lib.dstag_last_error()


def test_cpu_tensors_are_refused():
    a = torch.randn(4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        ct.contract_raw("ij,jk->ik", a, a)


# ---- emulator of the kernel's two-level addressing (csrc/gemm_simt.cu: axoff) ----------
def _axoff(ax, idx):
    idx = np.asarray(idx, dtype=np.int64)
    if ax.d1 == 1:
        return idx * ax.s0
    return (idx // ax.d1) * ax.s0 + (idx % ax.d1) * ax.s1


def _gather(op, base_ptr, flat, nb, nr, nc):
    off = (_axoff(op.batch, np.arange(nb))[:, None, None] + _axoff(op.row, np.arange(nr))[None, :, None]
           + _axoff(op.col, np.arange(nc))[None, None, :])
    return flat[off + (op.ptr - base_ptr) // 4], off + (op.ptr - base_ptr) // 4


def emulate(g, tensors):
    """tensors: dict name -> torch CPU tensor whose storage backs the operand."""
    def flat(t):
        st = t.untyped_storage()
        return np.frombuffer((C.c_char * st.nbytes()).from_address(st.data_ptr()), dtype=np.float32), st.data_ptr()

    fa, pa = flat(tensors["A"]); fb, pb = flat(tensors["B"]); fc, pc = flat(tensors["C"])
    A, _ = _gather(g.A, pa, fa, g.batch, g.M, g.K)
    if g.Amul.ptr:
        fm, pm = flat(tensors["Amul"])
        A = A * _gather(g.Amul, pm, fm, g.batch, g.M, g.K)[0]
    B, _ = _gather(g.B, pb, fb, g.batch, g.K, g.N)
    R = g.alpha * np.einsum("bmk,bkn->bmn", A.astype(np.float64), B.astype(np.float64))
    if g.Cadd.ptr:
        fm, pm = flat(tensors["Cadd"])
        R = R + _gather(g.Cadd, pm, fm, g.batch, g.M, g.N)[0]
    if g.Cmul.ptr:
        fm, pm = flat(tensors["Cmul"])
        R = R * _gather(g.Cmul, pm, fm, g.batch, g.M, g.N)[0]
    if g.relu:
        R = np.maximum(R, 0)
    _, coff = _gather(g.C, pc, fc, g.batch, g.M, g.N)
    if g.accumulate == _lib.ACC_STORE:
        assert len(np.unique(coff)) == coff.size, "store-mode output must not alias"
        fc[coff] = R
    else:
        np.add.at(fc, coff.ravel(), R.ravel().astype(np.float32))


def _run(spec, A, B, out, **kw):
    g = ct.build_descriptor(spec, A, B, out, **kw)
    ts = {"A": A, "B": B, "C": out}
    for name, key in (("Amul", "a_mul"), ("Cadd", "c_add"), ("Cmul", "c_mul")):
        if kw.get(key) is not None:
            ts[name] = kw[key][1]
    emulate(g, ts)
    return g


def test_descriptor_plain_and_permuted_views():
    torch.manual_seed(0)
    B_, F_, T_, N_, D_ = 2, 3, 4, 5, 6
    O = torch.randn(B_, F_, T_, N_)
    W = torch.randn(D_, T_, 1, F_)
    out = torch.zeros(B_, N_, D_)
    bias = torch.randn(D_)
    g = _run("bftn,dtf->bnd", O, W[:, :, 0, :], out, c_add=("d", bias))
    ref = torch.einsum("bftn,dtf->bnd", O, W[:, :, 0, :]) + bias
    assert torch.allclose(out, ref, atol=1e-5)
    assert (g.M, g.N, g.K, g.batch) == (B_ * N_, D_, F_ * T_, 1)
    # the reduction index is enumerated (t,f): contiguous in the conv weight, two-level in O
    assert (g.B.row.s0, g.B.row.d1) == (1, 1) and (g.A.col.d1, g.A.col.s0, g.A.col.s1) == (F_, N_, T_ * N_)


def test_descriptor_batch_pair_amul_and_padded_output():
    torch.manual_seed(1)
    Bz, K, N, E, ld = 2, 3, 7, 5, 8
    P = torch.randn(Bz, K, N, ld)
    X = torch.randn(Bz, N, E)
    T = torch.randn(K, N, ld)
    Z = torch.zeros(K, Bz, N, E)
    _run("bkji,bie->kbje", P[..., :N], X, Z, a_mul=("kji", T[..., :N]))
    ref = torch.einsum("bkji,kji,bie->kbje", P[..., :N], T[..., :N], X)
    assert torch.allclose(Z, ref, atol=1e-5)
    # backward-style role swap with an epilogue multiplier into a padded buffer
    dS = torch.zeros(Bz, K, N, ld)
    _run("kbje,bie->bkji", Z, X, dS[..., :N], c_mul=("kji", T[..., :N]))
    ref = torch.einsum("kbje,bie,kji->bkji", Z, X, T[..., :N])
    assert torch.allclose(dS[..., :N], ref, atol=1e-4) and float(dS[..., N:].abs().max()) == 0.0


def test_descriptor_unfolded_conv_view_and_atomic_scatter():
    torch.manual_seed(2)
    X, Cc, T, ks = 3, 4, 9, 3
    g3 = torch.randn(X, Cc, T)
    w = torch.randn(2 * Cc, Cc, 1, ks)
    b = torch.randn(2 * Cc)
    u = torch.zeros(X, 2 * Cc, T - ks + 1)
    _run("xctd,ocd->xot", g3.unfold(2, ks, 1), w.view(2 * Cc, Cc, ks), u, c_add=("o", b))
    ref = torch.nn.functional.conv2d(g3.permute(1, 0, 2).unsqueeze(0), w, b)[0].permute(1, 0, 2)
    assert torch.allclose(u, ref, atol=1e-5)
    # input gradient scattered through the overlapping view (atomic accumulation)
    du = torch.randn_like(u)
    dg = torch.zeros_like(g3)
    _run("xot,ocd->xctd", du, w.view(2 * Cc, Cc, ks), dg.unfold(2, ks, 1), accumulate=_lib.ACC_ATOMIC)
    g3r = g3.clone().requires_grad_(True)
    torch.nn.functional.conv2d(g3r.permute(1, 0, 2).unsqueeze(0), w, b)[0].permute(1, 0, 2).backward(du)
    assert torch.allclose(dg, g3r.grad, atol=1e-4)


def test_descriptor_scores_with_packed_qk_and_bias_matrix():
    torch.manual_seed(3)
    Bz, N, K, dk, ld = 2, 6, 2, 4, 8
    qk = torch.randn(Bz, N, 2, K, dk)
    bias = torch.randn(K, N, ld)
    P = torch.zeros(Bz, K, N, ld)
    _run("bjkd,bikd->bkji", qk[:, :, 1], qk[:, :, 0], P[..., :N], alpha=0.5, c_add=("kji", bias[..., :N]))
    ref = 0.5 * torch.einsum("bjkd,bikd->bkji", qk[:, :, 1], qk[:, :, 0]) + bias[..., :N]
    assert torch.allclose(P[..., :N], ref, atol=1e-5)


def test_plan_rejects_unsupported_specs():
    with pytest.raises(ValueError):
        ct.plan("abcd,de->abce")        # m would span 3 dims
    with pytest.raises(ValueError):
        ct.plan("ab,bc->ad")
    with pytest.raises(ValueError):
        ct.plan("abz,bc->ac")           # z summed over a single operand


def test_auto_split_k():
    assert ct.auto_split_k(512, 384, 56512, 1) > 1      # pre_conv weight gradient at PEMS07
    assert ct.auto_split_k(56512, 512, 384, 1) == 1


# ---- module API ------------------------------------------------------------------------
@pytest.mark.parametrize("name", GOLDEN)
def test_make_model_matches_reference_state_dict_and_init(name):
    """Same seed -> bit-identical weights as the reference's make_model (fixture weights were
    produced by the reference under torch.manual_seed(1)); same keys, shapes, no buffers."""
    d, cfg, w, g, cheb = load_golden(name)
    torch.manual_seed(1)
    net = make_model("cpu", 1, cfg["nb"], 1, cfg["K"], 32, 32, 1, torch.from_numpy(d["adj_tmd"]),
                     torch.from_numpy(d["adj_pa"]), torch.from_numpy(d["adj_tmd"]), cfg["P"], cfg["T"],
                     cfg["N"], cfg["D"], 32, 32, cfg["H"])
    sd = net.state_dict()
    assert list(sd.keys()) == list(w.keys())
    for k in w:
        assert sd[k].shape == w[k].shape, k
        assert torch.equal(sd[k], w[k]), k
    net.load_state_dict(w)                                   # reference checkpoint loads
    # Chebyshev constants equal the reference's (transposed, padded)
    blk = net.BlockList[0]
    for k, t in enumerate(cheb[:cfg["K"]]):
        assert torch.allclose(blk.cheb_T[k, :, :cfg["N"]], t.t(), atol=2e-5, rtol=1e-4)
        assert float(blk.cheb_T[k, :, cfg["N"]:].abs().max()) == 0.0 if blk.ld > cfg["N"] else True


def test_forward_on_cpu_fails_loudly():
    d, cfg, w, g, cheb = load_golden("g_n20")
    net = make_model("cpu", 1, cfg["nb"], 1, cfg["K"], 32, 32, 1, d["adj_tmd"], d["adj_pa"], d["adj_tmd"],
                     cfg["P"], cfg["T"], cfg["N"], cfg["D"], 32, 32, cfg["H"])
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.from_numpy(d["x"]))
