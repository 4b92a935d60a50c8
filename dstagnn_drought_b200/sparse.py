"""Sparse Chebyshev path of ``cheb_conv_withSAt`` (model/DSTAGNN_my.py:117-133): host side.

The reference builds its polynomials with an ELEMENT-wise recurrence (lib/utils.py:198-203), so T_0 = I and
every T_k has the sparsity pattern of the adjacency matrix -- about 1 % with the STAG graphs the model is trained
on (data/STAG_gen.py:108-116).  ``T_k (.) softmax(S_k)`` and the part of its gradient that survives the Hadamard
product therefore live on the graph's edges.  ``EdgeLists`` turns the (constant) polynomials and ``adj_pa`` into
CSR index arrays once, at model-build time; ``SparseChebConv`` is a drop-in for ``ops.SpatialChebConv`` (same
inputs, same output, same gradients) that touches only those edges -- see csrc/sparse.cu for the formulas.

Which path a block takes is decided when it is built (``choose_path``): the sparse one when every T_k is at most
``SPARSE_MAX_DENSITY`` dense, the dense tensor-core contraction otherwise; ``DSTAG_CHEB=dense|sparse`` overrides.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Sequence

import numpy as np
import torch

from . import _lib
from .contraction import _check, contract_raw, zeros

SPARSE_MAX_DENSITY = 0.05
SHORT_CHAIN = int(os.environ.get("DSTAG_SHORT_CHAIN", "64"))
EXACT_BELOW = 128                # graphs up to this many nodes: P Qc and P^T (r K) on the exact fp32 engine


def choose_path(cheb_polynomials: Sequence, d_k: int) -> str:
    """'sparse' or 'dense' for a block with these polynomials (``DSTAG_CHEB`` overrides the density rule)."""
    mode = os.environ.get("DSTAG_CHEB", "auto").lower()
    if mode not in ("auto", "dense", "sparse"):
        raise ValueError(f"DSTAG_CHEB must be auto, dense or sparse (got {mode!r})")
    if d_k > 32 or len(cheb_polynomials) * d_k > 128:
        return "dense"                               # dstag_dqk_combine keeps one lane per d_k component
    if mode != "auto":
        return mode
    n = np.asarray(cheb_polynomials[0]).shape[0]
    dens = max(float(np.count_nonzero(np.asarray(t))) / float(n * n) for t in cheb_polynomials)
    return "sparse" if dens <= SPARSE_MAX_DENSITY else "dense"


class EdgeLists:
    """CSR / transposed-CSR edge lists of the K polynomials plus the non-zeros of ``adj_pa`` (int32 / fp32 CPU
    tensors; the block registers them as non-persistent buffers so that ``.to(device)`` moves them).

    Edge e of order k joins OUTPUT node j to INPUT node i, weight ``tval[e] = T_k[i, j]``:
      row CSR   rows r = k N + j: ``rowptr`` (K N + 1), ``col`` = i, ``erow`` = r, ``tval``
      column CSR rows k N + i:     ``t_rowptr``, ``t_col`` = j, ``t_eid`` = position of the edge in the row CSR
      adj_pa    ``pa_j``, ``pa_i`` (npa), ``pa_eid`` (K, npa) = position of edge (k, j, i) in the row CSR or -1
    """

    NAMES = ("rowptr", "col", "erow", "tval", "t_rowptr", "t_col", "t_eid", "pa_j", "pa_i", "pa_eid")

    def __init__(self, cheb_polynomials: Sequence, adj_pa):
        polys = [np.asarray(t, dtype=np.float32) for t in cheb_polynomials]
        K, N = len(polys), polys[0].shape[0]
        cols, rows, vals, rowptr = [], [], [], [0]
        t_cols, t_keys = [], []
        for k, T in enumerate(polys):
            jj, ii = np.nonzero(T.T)                              # sorted by j, then i
            cols.append(ii)
            rows.append(k * N + jj)
            vals.append(T[ii, jj])
            counts = np.bincount(jj, minlength=N)
            rowptr.extend((rowptr[-1] + np.cumsum(counts)).tolist())
        col = np.concatenate(cols).astype(np.int32)
        erow = np.concatenate(rows).astype(np.int32)
        tval = np.concatenate(vals).astype(np.float32)
        E = col.shape[0]
        # transposed lists: sort the edges by (k, i, j)
        k_of = erow // N
        j_of = erow % N
        order = np.lexsort((j_of, col, k_of))
        t_eid = order.astype(np.int32)
        t_col = j_of[order].astype(np.int32)
        t_counts = np.bincount(k_of[order] * N + col[order], minlength=K * N)
        t_rowptr = np.concatenate([[0], np.cumsum(t_counts)]).astype(np.int32)
        # adj_pa non-zeros and their edge ids per order
        pa = np.asarray(adj_pa)
        pi, pj = np.nonzero(pa)
        key = erow.astype(np.int64) * N + col                     # unique, ascending (row CSR order)
        pa_eid = np.full((K, pi.shape[0]), -1, np.int32)
        for k in range(K):
            want = (k * N + pj).astype(np.int64) * N + pi
            pos = np.searchsorted(key, want)
            pos = np.minimum(pos, max(E - 1, 0))
            hit = key[pos] == want if E else np.zeros_like(want, bool)
            pa_eid[k, hit] = pos[hit]
        self.K, self.N, self.E, self.npa = K, N, int(E), int(pi.shape[0])
        self.tensors: Dict[str, torch.Tensor] = {
            "rowptr": torch.from_numpy(np.asarray(rowptr, np.int32)), "col": torch.from_numpy(col),
            "erow": torch.from_numpy(erow), "tval": torch.from_numpy(tval),
            "t_rowptr": torch.from_numpy(t_rowptr), "t_col": torch.from_numpy(t_col), "t_eid": torch.from_numpy(t_eid),
            "pa_j": torch.from_numpy(pj.astype(np.int32)), "pa_i": torch.from_numpy(pi.astype(np.int32)),
            "pa_eid": torch.from_numpy(pa_eid)}


class SparseChebConv(torch.autograd.Function):
    """Drop-in for ``ops.SpatialChebConv`` on a sparse graph: qk (B,N,2,K,dk), biasT / tkT (K,N,ld), x3 (B,N,F T),
    ``g`` = dict of the ``EdgeLists`` tensors on the device -> Z (K,B,N,F T)."""

    @staticmethod
    def forward(ctx, qk, biasT, tkT, x3, g):
        from .ops import _stream, spatial_maps
        lib = _lib.load()
        for t, n in ((qk, "qk"), (biasT, "biasT"), (x3, "x")):
            _check(t, n)
        qk, x3 = qk.contiguous(), x3.contiguous()
        B, N, _, K, dk = qk.shape
        L = x3.shape[-1]
        if L % 4:
            raise RuntimeError(f"sparse Chebyshev path: F*T = {L} must be a multiple of 4")
        E = g["col"].numel()
        P = spatial_maps(qk, biasT)                               # (B, K, N, ld) softmax maps, no T_k (.) P copy
        ld = P.shape[-1]
        a = torch.empty((B, E), device=qk.device, dtype=torch.float32)
        st = _stream(qk)
        _lib.check(lib.dstag_edge_probs(P.data_ptr(), g["erow"].data_ptr(), g["col"].data_ptr(), g["tval"].data_ptr(),
                                        a.data_ptr(), B, K, N, ld, E, st), "dstag_edge_probs")
        Z = torch.empty((K, B, N, L), device=qk.device, dtype=torch.float32)
        _lib.check(lib.dstag_spmm_expand(x3.data_ptr(), a.data_ptr(), g["rowptr"].data_ptr(), g["col"].data_ptr(), None,
                                         Z.data_ptr(), B, K, N, L, E, st), "dstag_spmm_expand")
        ctx.save_for_backward(qk, P, x3, a)
        ctx.g = g
        return Z

    @staticmethod
    def backward(ctx, dZ):
        from .ops import _stream
        lib = _lib.load()
        qk, P, x3, a = ctx.saved_tensors
        g = ctx.g
        B, N, _, K, dk = qk.shape
        ld, L, E = P.shape[-1], x3.shape[-1], a.shape[1]
        dZ = dZ.contiguous()
        st = _stream(qk)
        dev = qk.device
        dx = None
        if ctx.needs_input_grad[3]:
            dx = torch.empty_like(x3)
            _lib.check(lib.dstag_spmm_reduce(dZ.data_ptr(), a.data_ptr(), g["t_rowptr"].data_ptr(), g["t_col"].data_ptr(),
                                             g["t_eid"].data_ptr(), dx.data_ptr(), B, K, N, L, E, st), "dstag_spmm_reduce")
        h = torch.empty((B, E), device=dev, dtype=torch.float32)
        r = torch.empty((B, K, N), device=dev, dtype=torch.float32)
        _lib.check(lib.dstag_edge_dots(dZ.data_ptr(), x3.data_ptr(), a.data_ptr(), g["rowptr"].data_ptr(),
                                       g["col"].data_ptr(), h.data_ptr(), r.data_ptr(), B, K, N, L, E, st), "dstag_edge_dots")
        rK = torch.empty((B, N, K, dk), device=dev, dtype=torch.float32)
        _lib.check(lib.dstag_scale_keys(qk.data_ptr(), r.data_ptr(), rK.data_ptr(), B, N, K, dk, st), "dstag_scale_keys")
        Qc = torch.empty((B, N, K, dk), device=dev, dtype=torch.float32)
        qsum = torch.empty((B, K * dk), device=dev, dtype=torch.float32)
        _lib.check(lib.dstag_center_queries(qk.data_ptr(), qsum.data_ptr(), Qc.data_ptr(), B, N, K, dk, st),
                   "dstag_center_queries")
        Pv = P[..., :N]
        # the only dense work of the backward: two N x N x d_k products with the (non-negative) softmax maps
        # (accumulated in chains of 64 terms: the tensor core's accumulator truncates, and the bias of a chain
        #  of N terms reached 1e-3 of the projection-bias gradients at N = 4096)
        # Small graphs take the exact fp32 engine: the products are tiny there, and with one or two edges per row the
        # edge half of the key gradient is exact while a 3xTF32 P Qc is not -- their difference showed at 1.2e-4.
        prec = _lib.PREC_FP32 if N <= EXACT_BELOW else None
        O = contract_raw("bkji,bikd->bjkd", Pv, Qc, tag="sat_PQ", chain=SHORT_CHAIN, precision=prec)     # (B, N, K, dk)
        W = contract_raw("bkji,bjkd->bikd", Pv, rK, tag="sat_PtrK", chain=SHORT_CHAIN, precision=prec)   # (B, N, K, dk)
        dqk = torch.empty_like(qk)
        _lib.check(lib.dstag_dqk_combine(qk.data_ptr(), Qc.data_ptr(), h.data_ptr(), r.data_ptr(), O.data_ptr(), W.data_ptr(),
                                         g["rowptr"].data_ptr(), g["col"].data_ptr(), g["t_rowptr"].data_ptr(),
                                         g["t_col"].data_ptr(), g["t_eid"].data_ptr(), dqk.data_ptr(), B, N, K, dk, E,
                                         1.0 / math.sqrt(dk), st), "dstag_dqk_combine")
        dbias = None
        if ctx.needs_input_grad[1]:
            dbias = zeros((K, N, ld), dev)
            npa = g["pa_j"].numel()
            if npa:
                _lib.check(lib.dstag_dbias_sparse(P.data_ptr(), h.data_ptr(), r.data_ptr(), g["pa_j"].data_ptr(),
                                                  g["pa_i"].data_ptr(), g["pa_eid"].data_ptr(), dbias.data_ptr(), B, K, N,
                                                  ld, E, npa, st), "dstag_dbias_sparse")
        return dqk, dbias, None, dx, None
