"""TEST INFRASTRUCTURE ONLY: a numpy emulation of the C ABI in include/dstagnn_b200.h.

It lets the CPU test-suite drive the *host-side wiring* (einsum specs, argument order of
every ctypes call, the hand-written backward formulas, the res_att gradient chain) end to end
against the golden fixtures without a GPU.  It is installed by monkeypatching
``dstagnn_drought_b200._lib.load`` inside a pytest fixture; the product package never imports
this file and has no CPU path of its own.
"""
import ctypes as C
import math

import numpy as np


def _arr(ptr, n):
    return np.ctypeslib.as_array((C.c_float * int(n)).from_address(int(ptr)))


def _axoff(ax, idx):
    idx = np.asarray(idx, dtype=np.int64)
    if ax.d1 == 1:
        return idx * ax.s0
    return (idx // ax.d1) * ax.s0 + (idx % ax.d1) * ax.s1


def _offsets(op, nb, nr, nc):
    return (_axoff(op.batch, np.arange(nb))[:, None, None] + _axoff(op.row, np.arange(nr))[None, :, None]
            + _axoff(op.col, np.arange(nc))[None, None, :])


def _read(op, nb, nr, nc):
    off = _offsets(op, nb, nr, nc)
    lo, hi = int(off.min()), int(off.max())
    base = _arr(op.ptr + 4 * lo, hi - lo + 1) if lo >= 0 else None
    if base is None:
        raise ValueError("negative offset")
    return base[off - lo]


class FakeLib:
    def __init__(self):
        self.launches = 0
        self.err = b""

    # ---- plumbing ----
    def dstag_abi_version(self):
        return 19

    def dstag_last_error(self):
        return self.err

    def dstag_launch_count(self):
        return self.launches

    def dstag_reset_launch_count(self):
        self.launches = 0

    def dstag_gemm_uses_tensor_cores(self, gref):
        return 0

    # ---- contraction ----
    def dstag_gemm(self, gref, stream):
        g = gref._obj
        self.launches += 1
        A = _read(g.A, g.batch, g.M, g.K).astype(np.float64)
        if g.Amul.ptr:
            A = A * _read(g.Amul, g.batch, g.M, g.K)
        B = _read(g.B, g.batch, g.K, g.N).astype(np.float64)
        R = g.alpha * np.einsum("bmk,bkn->bmn", A, B)
        if g.Cadd.ptr:
            R = R + _read(g.Cadd, g.batch, g.M, g.N)
        if g.Cmul.ptr:
            R = R * _read(g.Cmul, g.batch, g.M, g.N)
        if g.relu:
            R = np.maximum(R, 0)
        off = _offsets(g.C, g.batch, g.M, g.N)
        lo, hi = int(off.min()), int(off.max())
        out = _arr(g.C.ptr + 4 * lo, hi - lo + 1)
        if g.accumulate == 0:
            assert len(np.unique(off)) == off.size, "store-mode output aliases"
            out[off - lo] = R
            if g.split_k > 1:
                # two-stage reduction: one slice per K split.  The emulation puts the whole product into
                # slice 0 and zeros into the others (their fixed-order sum is the same result).
                assert g.split_stride != 0, "split_k > 1 with plain stores needs split_stride"
                for s in range(1, g.split_k):
                    _arr(g.C.ptr + 4 * (lo + s * g.split_stride), hi - lo + 1)[off - lo] = 0.0
        else:
            np.add.at(out, (off - lo).ravel(), R.ravel().astype(np.float32))
        return 0

    def dstag_sum_slices(self, ws, slice_stride, nslices, out, rows, cols, out_ld, stream):
        self.launches += 1
        acc = np.zeros((rows, cols), np.float32)
        for s in range(nslices):
            acc = acc + _arr(ws + 4 * s * slice_stride, rows * cols).reshape(rows, cols)
        o = _arr(out, (rows - 1) * out_ld + cols)
        idx = (np.arange(rows, dtype=np.int64) * out_ld)[:, None] + np.arange(cols, dtype=np.int64)[None, :]
        o[idx] = acc
        return 0

    # ---- row kernels ----
    @staticmethod
    def _ln_pre(a, a_s0, a_s1, a_div, a_sc, tab, tab_mod, tab_ld, rows, L):
        r = np.arange(rows, dtype=np.int64)
        off = ((r // a_div) * a_s0 + (r % a_div) * a_s1)[:, None] + (np.arange(L, dtype=np.int64) * a_sc)[None, :]
        pre = _arr(a, off.max() + 1)[off].astype(np.float64)
        if tab:
            pre = pre + _arr(tab, tab_mod * tab_ld).reshape(tab_mod, tab_ld)[r % tab_mod, :L]
        return pre

    def dstag_layernorm_fwd(self, a, a_s0, a_s1, a_div, a_sc, tab, tab_mod, tab_ld, gamma, beta, eps, y, y_ld,
                            mean, rstd, rows, L, drop_p, seed, seed_dev, stream):
        assert drop_p == 0.0, "emulation has no dropout"
        self.launches += 1
        pre = self._ln_pre(a, a_s0, a_s1, a_div, a_sc, tab, tab_mod, tab_ld, rows, L)
        mu = pre.mean(1)
        rs = 1.0 / np.sqrt(pre.var(1) + eps)
        _arr(y, rows * y_ld).reshape(rows, y_ld)[:, :L] = ((pre - mu[:, None]) * rs[:, None] * _arr(gamma, L)
                                                          + _arr(beta, L))
        _arr(mean, rows)[:] = mu
        _arr(rstd, rows)[:] = rs
        return 0

    def dstag_layernorm_bwd(self, dy, dy_ld, a, a_s0, a_s1, a_div, a_sc, tab, tab_mod, tab_ld, gamma, mean, rstd,
                            dpre, dpre_ld, dgamma, dbeta, rows, L, drop_p, seed, seed_dev, stream):
        assert drop_p == 0.0
        self.launches += 1
        pre = self._ln_pre(a, a_s0, a_s1, a_div, a_sc, tab, tab_mod, tab_ld, rows, L)
        mu = _arr(mean, rows).astype(np.float64)[:, None]
        rs = _arr(rstd, rows).astype(np.float64)[:, None]
        xh = (pre - mu) * rs
        d = _arr(dy, rows * dy_ld).reshape(rows, dy_ld)[:, :L].astype(np.float64)
        g = d * _arr(gamma, L)
        _arr(dpre, rows * dpre_ld).reshape(rows, dpre_ld)[:, :L] = rs * (g - g.mean(1, keepdims=True)
                                                                        - xh * (g * xh).mean(1, keepdims=True))
        _arr(dgamma, L)[:] += (d * xh).sum(0)
        _arr(dbeta, L)[:] += d.sum(0)
        return 0

    def dstag_colsum(self, inp, out, rows, cols, ld, stream):
        self.launches += 1
        m = _arr(inp, (rows - 1) * ld + cols)
        acc = np.zeros(cols)
        for r in range(rows):
            acc += m[r * ld:r * ld + cols]
        _arr(out, cols)[:] += acc
        return 0

    def dstag_transpose(self, inp, in_ld, out, out_ld, batch, R, Cc, r_d, c_d, stream):
        self.launches += 1
        src = _arr(inp, batch * R * in_ld).reshape(batch, R, in_ld)[:, :, :Cc]
        pr = (np.arange(R) % r_d) * (R // r_d) + np.arange(R) // r_d
        pc = (np.arange(Cc) % c_d) * (Cc // c_d) + np.arange(Cc) // c_d
        dst = _arr(out, batch * Cc * out_ld).reshape(batch, Cc, out_ld)
        dst[:, pc[:, None], pr[None, :]] = src.transpose(0, 2, 1)
        return 0

    def dstag_softmax_fwd(self, x, rows, L, ld, stream):
        self.launches += 1
        m = _arr(x, rows * ld).reshape(rows, ld)
        s = m[:, :L].astype(np.float64)
        e = np.exp(s - s.max(1, keepdims=True))
        m[:, :L] = e / e.sum(1, keepdims=True)
        m[:, L:] = 0
        return 0

    def dstag_softmax_bwd(self, p, dp, mul, mul_rows, rows, L, ld, stream):
        self.launches += 1
        P = _arr(p, rows * ld).reshape(rows, ld)[:, :L].astype(np.float64)
        D = _arr(dp, rows * ld).reshape(rows, ld)
        d = D[:, :L].astype(np.float64)
        if mul:
            d = d * _arr(mul, mul_rows * ld).reshape(mul_rows, ld)[np.arange(rows) % mul_rows, :L]
        D[:, :L] = P * (d - (P * d).sum(1, keepdims=True))
        D[:, L:] = 0
        return 0

    def dstag_sat_softmax_fwd(self, qk, biasT, tkT, P, TP, B, N, K, dk, ld, stream):
        self.launches += 1
        q = _arr(qk, B * N * 2 * K * dk).reshape(B, N, 2, K, dk).astype(np.float64)
        s = np.einsum("bjkd,bikd->bkji", q[:, :, 1], q[:, :, 0]) / math.sqrt(dk)
        s = s + _arr(biasT, K * N * ld).reshape(K, N, ld)[None, :, :, :N]
        e = np.exp(s - s.max(-1, keepdims=True))
        out = _arr(P, B * K * N * ld).reshape(B, K, N, ld)
        out[..., :N] = e / e.sum(-1, keepdims=True)
        out[..., N:] = 0
        if TP:
            _arr(TP, B * K * N * ld).reshape(B, K, N, ld)[:] = out * _arr(tkT, K * N * ld).reshape(1, K, N, ld)
        return 0

    def dstag_relu_bwd(self, y, dy, dx, n, stream):
        self.launches += 1
        _arr(dx, n)[:] = np.where(_arr(y, n) > 0, _arr(dy, n), 0)
        return 0

    # ---- temporal attention ----
    def dstag_tattn_fwd(self, qkv, res, res_F, scores, ctx, B, F, H, T, dk, stream):
        self.launches += 1
        q = _arr(qkv, B * F * T * 3 * H * dk).reshape(B, F, T, 3, H, dk).astype(np.float64)
        Q, K, V = (q[:, :, :, i].transpose(0, 1, 3, 2, 4) for i in range(3))            # (B,F,H,T,dk)
        S = Q @ K.transpose(0, 1, 2, 4, 3) / math.sqrt(dk)
        if res_F:
            S = S + _arr(res, B * res_F * H * T * T).reshape(B, res_F, H, T, T)
        _arr(scores, B * F * H * T * T).reshape(B, F, H, T, T)[:] = S
        e = np.exp(S - S.max(3, keepdims=True))
        P = e / e.sum(3, keepdims=True)                                                  # over the QUERY axis
        _arr(ctx, B * F * T * H * dk).reshape(B, F, T, H, dk)[:] = (P @ V).transpose(0, 1, 3, 2, 4)
        return 0

    def dstag_tattn_bwd(self, qkv, scores, dctx, dsext, dqkv, dsout, B, F, H, T, dk, stream):
        self.launches += 1
        q = _arr(qkv, B * F * T * 3 * H * dk).reshape(B, F, T, 3, H, dk).astype(np.float64)
        Q, K, V = (q[:, :, :, i].transpose(0, 1, 3, 2, 4) for i in range(3))
        S = _arr(scores, B * F * H * T * T).reshape(B, F, H, T, T).astype(np.float64)
        e = np.exp(S - S.max(3, keepdims=True))
        P = e / e.sum(3, keepdims=True)
        dC = _arr(dctx, B * F * T * H * dk).reshape(B, F, T, H, dk).transpose(0, 1, 3, 2, 4).astype(np.float64)
        dV = P.transpose(0, 1, 2, 4, 3) @ dC
        dP = dC @ V.transpose(0, 1, 2, 4, 3)
        dS = P * (dP - (P * dP).sum(3, keepdims=True))
        if dsext:
            dS = dS + _arr(dsext, B * F * H * T * T).reshape(B, F, H, T, T)
        _arr(dsout, B * F * H * T * T).reshape(B, F, H, T, T)[:] = dS
        dQ = dS @ K / math.sqrt(dk)
        dK = dS.transpose(0, 1, 2, 4, 3) @ Q / math.sqrt(dk)
        out = _arr(dqkv, B * F * T * 3 * H * dk).reshape(B, F, T, 3, H, dk)
        for i, t in enumerate((dQ, dK, dV)):
            out[:, :, :, i] = t.transpose(0, 1, 3, 2, 4)
        return 0

    # ---- GTU gate + tail ----
    def dstag_gtu_gate_fwd(self, u, out, X, Cc, Tp, S, off, stream):
        self.launches += 1
        U = _arr(u, X * Tp * 2 * Cc).reshape(X, Tp, 2 * Cc).astype(np.float64)
        O = _arr(out, X * S * Cc).reshape(X, S, Cc)                                   # channels-last
        O[:, off:off + Tp, :] = np.tanh(U[:, :, :Cc]) / (1 + np.exp(-U[:, :, Cc:]))
        return 0

    def dstag_gtu_tc_supported(self, Cc, T, ks, backward):
        return int(Cc == 32 and ks in (3, 5, 7) and T >= ks and (not backward or T - ks + 1 <= 128))

    def dstag_gtu_conv_gate_fwd(self, g, w, bias, cat, sv, X, Cc, T, ks, S, off, stream):
        self.launches += 1
        Tp = T - ks + 1
        G = _arr(g, X * T * Cc).reshape(X, T, Cc).astype(np.float64)
        W = _arr(w, 2 * Cc * Cc * ks).reshape(2 * Cc, Cc, ks).astype(np.float64)
        win = np.stack([G[:, j:j + Tp, :] for j in range(ks)], 3)                         # (X, Tp, C, ks)
        U = np.einsum("xtcj,ocj->xto", win, W) + _arr(bias, 2 * Cc).astype(np.float64)
        th, sg = np.tanh(U[:, :, :Cc]), 1 / (1 + np.exp(-U[:, :, Cc:]))
        _arr(cat, X * S * Cc).reshape(X, S, Cc)[:, off:off + Tp, :] = th * sg
        if sv:
            SV = _arr(sv, X * Tp * 2 * Cc).reshape(X, Tp, 2 * Cc)
            SV[:, :, :Cc], SV[:, :, Cc:] = th, sg
        return 0

    def dstag_gram_slices(self):
        return 3

    def dstag_gram_supported(self, taps, NB, T, Tp):
        return int(1 <= taps <= 8 and NB in (32, 64) and 1 <= Tp <= min(T, 32))

    def dstag_gram_rows(self, a, b, partial, X, T, Tp, taps, tap_stride, NB, stream):
        self.launches += 1
        MT = 1 if taps <= 4 else 2
        Bm = _arr(b, X * Tp * NB).reshape(X * Tp, NB).astype(np.float64)
        P = _arr(partial, 3 * MT * 128 * NB).reshape(3, MT * 128, NB)
        P[:] = 0
        rows = (np.arange(X)[:, None] * T + np.arange(Tp)[None, :]).reshape(-1)
        span = int(rows.max()) * 32 + (taps - 1) * tap_stride + 32
        A = _arr(a, span).astype(np.float64)
        for j in range(taps):
            Aj = A[(rows[:, None] * 32 + j * tap_stride + np.arange(32)[None, :])]          # (R, 32)
            P[1, 32 * j:32 * j + 32, :] = Aj.T @ Bm
        return 0

    def dstag_gtu_bwd_ws_rows(self):
        return 8

    def dstag_gtu_gate_conv_bwd(self, sv, dcat, w, du, dg, db_ws, X, Cc, T, ks, S, off, accumulate, stream):
        self.launches += 1
        Tp = T - ks + 1
        SV = _arr(sv, X * Tp * 2 * Cc).reshape(X, Tp, 2 * Cc).astype(np.float64)
        th, sg = SV[:, :, :Cc], SV[:, :, Cc:]
        d = _arr(dcat, X * S * Cc).reshape(X, S, Cc)[:, off:off + Tp, :].astype(np.float64)
        DU = np.concatenate([d * sg * (1 - th * th), d * th * sg * (1 - sg)], 2)          # (X, Tp, 2C)
        _arr(du, X * Tp * 2 * Cc).reshape(X, Tp, 2 * Cc)[:] = DU
        if db_ws:
            _arr(db_ws, 8 * 2 * Cc).reshape(8, 2 * Cc)[0] = DU.reshape(-1, 2 * Cc).sum(0)
        W = _arr(w, 2 * Cc * Cc * ks).reshape(2 * Cc, Cc, ks).astype(np.float64)
        Y = np.einsum("xto,ocj->xtjc", DU, W)
        DG = np.zeros((X, T, Cc))
        for j in range(ks):
            DG[:, j:j + Tp, :] += Y[:, :, j, :]
        out = _arr(dg, X * T * Cc).reshape(X, T, Cc)
        out[:] = (out + DG) if accumulate else DG
        return 0

    # ---- thin-feature contractions (csrc/thin.cu) ----
    def dstag_thin_supported(self, J, Cc):
        Q = Cc // 4
        return int(1 <= J <= 4 and Cc % 4 == 0 and 1 <= Q <= 32 and Q & (Q - 1) == 0)

    def dstag_thin_slices(self):
        return 3

    @staticmethod
    def _thin_operands(a, a_plane, taps, F, w, w_js, w_cs, R, Cc):
        J = taps * F
        span_a = (taps - 1) * a_plane + R * F
        A = _arr(a, span_a)
        idx = [(j // F) * a_plane + np.arange(R) * F + j % F for j in range(J)]
        Wm = _arr(w, (J - 1) * w_js + (Cc - 1) * w_cs + 1)
        W = np.array([[Wm[j * w_js + c * w_cs] for c in range(Cc)] for j in range(J)], dtype=np.float64)
        return A, idx, W

    def dstag_thin_expand(self, a, a_plane, taps, F, w, w_js, w_cs, bias, out, R, Cc, relu, stream):
        self.launches += 1
        A, idx, W = self._thin_operands(a, a_plane, taps, F, w, w_js, w_cs, R, Cc)
        Z = np.stack([A[i] for i in idx], 1).astype(np.float64)                  # (R, J)
        O = Z @ W
        if bias:
            O = O + _arr(bias, Cc).astype(np.float64)
        _arr(out, R * Cc).reshape(R, Cc)[:] = np.maximum(O, 0) if relu else O
        return 0

    def dstag_thin_bwd(self, a, a_plane, taps, F, w, w_js, w_cs, d, y, dz, partial, R, Cc, stream):
        self.launches += 1
        J = taps * F
        A, idx, W = self._thin_operands(a, a_plane, taps, F, w, w_js, w_cs, R, Cc)
        D = _arr(d, R * Cc).reshape(R, Cc).astype(np.float64)
        if y:
            D = D * (_arr(y, R * Cc).reshape(R, Cc) > 0)
        if dz:
            DZ = _arr(dz, (taps - 1) * a_plane + R * F)
            P = D @ W.T                                                          # (R, J)
            for j, i in enumerate(idx):
                DZ[i] = P[:, j]
        if partial:
            Pm = _arr(partial, 3 * (J + 1) * Cc).reshape(3, J + 1, Cc)
            Pm[:] = 0
            Z = np.stack([A[i] for i in idx], 1).astype(np.float64)
            Pm[1, :J] = Z.T @ D
            Pm[1, J] = D.sum(0)
        return 0

    def dstag_outer_sum_supported(self, m, n, Cc):
        return int(m > 0 and n > 0 and n % 4 == 0 and Cc % 4 == 0 and m * (n // 4) <= 256)

    def dstag_outer_sum(self, a, b, out, asum, X, m, n, Cc, stream):
        self.launches += 1
        A = _arr(a, X * m * Cc).reshape(X, m, Cc).astype(np.float64)
        Bm = _arr(b, X * n * Cc).reshape(X, n, Cc).astype(np.float64)
        _arr(out, m * n).reshape(m, n)[:] += np.einsum("xmc,xnc->mn", A, Bm)
        if asum:
            _arr(asum, m)[:] += A.sum((0, 2))
        return 0

    def dstag_gtu_gate_bwd_sv(self, sv, dout, dup, db, X, Cc, Tp, pad, S, off, stream):
        return self.dstag_gtu_gate_bwd(sv, dout, dup, db, X, Cc, Tp, pad, S, off, stream, from_sv=True)

    def dstag_gtu_gate_bwd(self, u, dout, dup, db, X, Cc, Tp, pad, S, off, stream, from_sv=False):
        self.launches += 1
        U = _arr(u, X * Tp * 2 * Cc).reshape(X, Tp, 2 * Cc).astype(np.float64)
        d = _arr(dout, X * S * Cc).reshape(X, S, Cc)[:, off:off + Tp, :].astype(np.float64)   # (X, Tp, C)
        th, sg = (U[:, :, :Cc], U[:, :, Cc:]) if from_sv else (np.tanh(U[:, :, :Cc]), 1 / (1 + np.exp(-U[:, :, Cc:])))
        D = _arr(dup, X * (Tp + 2 * pad) * 2 * Cc).reshape(X, Tp + 2 * pad, 2 * Cc)
        D[:] = 0
        D[:, pad:pad + Tp, :Cc] = d * sg * (1 - th * th)
        D[:, pad:pad + Tp, Cc:] = d * th * sg * (1 - sg)
        if db:
            _arr(db, 2 * Cc)[:] += D.reshape(-1, 2 * Cc).astype(np.float64).sum(0)
        return 0

    # ---- tail: every tensor (X, T, C) channels-last ----
    def dstag_tail_fwd(self, tc, g, r, gamma, beta, eps, out, mean, rstd, X, Cc, T, drop_p, seed, seed_dev, stream):
        assert drop_p == 0.0
        self.launches += 1
        n = X * Cc * T
        s1 = _arr(tc, n).reshape(X, T, Cc).astype(np.float64)
        if g:
            s1 = s1 + _arr(g, n).reshape(X, T, Cc)
        v = np.maximum(_arr(r, n).reshape(X, T, Cc) + np.maximum(s1, 0), 0)
        mu = v.mean(2)
        rs = 1 / np.sqrt(v.var(2) + eps)
        _arr(out, n).reshape(X, T, Cc)[:] = (v - mu[..., None]) * rs[..., None] * _arr(gamma, Cc) + _arr(beta, Cc)
        _arr(mean, X * T).reshape(X, T)[:] = mu
        _arr(rstd, X * T).reshape(X, T)[:] = rs
        return 0

    def dstag_tail_bwd(self, dout, tc, g, r, gamma, mean, rstd, dtc, dg, dr, dgamma, dbeta, X, Cc, T, drop_p,
                       seed, seed_dev, stream):
        assert drop_p == 0.0
        self.launches += 1
        n = X * Cc * T
        s1 = _arr(tc, n).reshape(X, T, Cc).astype(np.float64)
        if g:
            s1 = s1 + _arr(g, n).reshape(X, T, Cc)
        s2 = _arr(r, n).reshape(X, T, Cc) + np.maximum(s1, 0)
        v = np.maximum(s2, 0)
        mu = _arr(mean, X * T).reshape(X, T, 1).astype(np.float64)
        rs = _arr(rstd, X * T).reshape(X, T, 1).astype(np.float64)
        xh = (v - mu) * rs
        d = _arr(dout, n).reshape(X, T, Cc).astype(np.float64)
        gy = d * _arr(gamma, Cc)
        dv = rs * (gy - gy.mean(2, keepdims=True) - xh * (gy * xh).mean(2, keepdims=True))
        ds2 = dv * (s2 > 0)
        ds1 = ds2 * (s1 > 0)
        _arr(dr, n).reshape(X, T, Cc)[:] = ds2
        _arr(dtc, n).reshape(X, T, Cc)[:] = ds1
        if dg:
            _arr(dg, n).reshape(X, T, Cc)[:] = ds1
        _arr(dgamma, Cc)[:] += (d * xh).sum((0, 1))
        _arr(dbeta, Cc)[:] += d.sum((0, 1))
        return 0

    def dstag_tail_fused_supported(self, Cc, T):
        return 1 if (Cc == 32 and T == 12) else 0

    def _fcmy(self, cat, wf, bf, X, Cc, T):
        S = 3 * T - 12
        W = _arr(wf, T * S).reshape(T, S).astype(np.float64)
        tc = np.einsum("xsc,ts->xtc", _arr(cat, X * S * Cc).reshape(X, S, Cc).astype(np.float64), W)
        return W, np.ascontiguousarray((tc + _arr(bf, T)[None, :, None]).astype(np.float32))

    def dstag_tail_fused_fwd(self, cat, wf, bf, g, r, gamma, beta, eps, out, mean, rstd, X, Cc, T, drop_p, seed,
                             seed_dev, stream):
        _, self._tc = self._fcmy(cat, wf, bf, X, Cc, T)
        return self.dstag_tail_fwd(self._tc.ctypes.data, g, r, gamma, beta, eps, out, mean, rstd, X, Cc, T, drop_p,
                                   seed, seed_dev, stream)

    def dstag_tail_fused_bwd(self, dout, cat, wf, bf, g, r, gamma, mean, rstd, dcat, dtc, dg, dr, dgamma, dbeta, X,
                             Cc, T, drop_p, seed, seed_dev, stream):
        S = 3 * T - 12
        W, tc = self._fcmy(cat, wf, bf, X, Cc, T)
        rc = self.dstag_tail_bwd(dout, tc.ctypes.data, g, r, gamma, mean, rstd, dtc, dg, dr, dgamma, dbeta, X, Cc, T,
                                 drop_p, seed, seed_dev, stream)
        _arr(dcat, X * S * Cc).reshape(X, S, Cc)[:] = np.einsum(
            "xtc,ts->xsc", _arr(dtc, X * Cc * T).reshape(X, T, Cc).astype(np.float64), W)
        return rc


def _emu_optim(cls):
    """The step either side of the block stack (csrc/optim.cu)."""
    def dstag_smooth_l1_fwd(self, pred, target, loss, workspace, n, stream):
        self.launches += 1
        d = np.abs(_arr(pred, n).astype(np.float64) - _arr(target, n))
        _arr(loss, 1)[0] = np.mean(np.where(d < 1, 0.5 * d * d, d - 0.5))
        return 0

    def dstag_smooth_l1_bwd(self, pred, target, gout, dpred, n, stream):
        self.launches += 1
        d = _arr(pred, n).astype(np.float64) - _arr(target, n)
        g = _arr(gout, 1)[0] if gout else 1.0
        _arr(dpred, n)[:] = g * np.clip(d, -1, 1) / n
        return 0

    def dstag_adam_step(self, p, g, m, v, n, lr, b1, b2, eps, gscale, step_dev, step_host, stream):
        self.launches += 1
        t = float(np.ctypeslib.as_array((C.c_int64 * 1).from_address(int(step_dev)))[0]) if step_dev else float(step_host)
        P, G, M_, V = _arr(p, n), _arr(g, n).astype(np.float64) * gscale, _arr(m, n), _arr(v, n)
        mm = b1 * M_.astype(np.float64) + (1 - b1) * G
        vv = b2 * V.astype(np.float64) + (1 - b2) * G * G
        M_[:], V[:] = mm, vv
        P[:] = P - (lr / (1 - b1 ** t)) * mm / (np.sqrt(vv) / math.sqrt(1 - b2 ** t) + eps)
        return 0

    def dstag_gather_flat(self, srcs, offsets, sizes, n, flat, stream):
        self.launches += 1
        for t in range(n):
            sz = int(sizes[t])
            _arr(int(flat) + 4 * int(offsets[t]), sz)[:] = _arr(int(srcs[t]), sz)
        return 0

    def dstag_multi_copy(self, copies, n, stream):
        self.launches += 1
        for t in range(n):
            c = copies[t]
            idx = np.indices(tuple(c.n)).reshape(4, -1).astype(np.int64)
            so = sum(idx[k] * c.src_stride[k] for k in range(4))
            do = sum(idx[k] * c.dst_stride[k] for k in range(4))
            lo_s, lo_d = int(so.min()), int(do.min())
            src = _arr(int(c.src) + 4 * lo_s, int(so.max()) - lo_s + 1)
            dst = _arr(int(c.dst) + 4 * lo_d, int(do.max()) - lo_d + 1)
            dst[do - lo_d] = src[so - lo_s]
        return 0

    def dstag_tick(self, counters, n, stream):
        self.launches += 1
        np.ctypeslib.as_array((C.c_int64 * int(n)).from_address(int(counters)))[:] += 1
        return 0

    def dstag_bias_transpose(self, adj_pa, masks, K, biasT, N, ld, stream):
        self.launches += 1
        pa = _arr(adj_pa, N * N).reshape(N, N)
        out = _arr(biasT, K * N * ld).reshape(K, N, ld)
        out[:] = 0
        for k in range(K):
            out[k, :, :N] = (pa * _arr(masks[k], N * N).reshape(N, N)).T
        return 0

    def dstag_bias_transpose_bwd(self, adj_pa, dbiasT, dmask, K, N, ld, stream):
        self.launches += 1
        pa = _arr(adj_pa, N * N).reshape(N, N)
        d = _arr(dbiasT, K * N * ld).reshape(K, N, ld)
        out = _arr(dmask, K * N * N).reshape(K, N, N)
        for k in range(K):
            out[k] = pa * d[k, :, :N].T
        return 0

    for f in (dstag_smooth_l1_fwd, dstag_smooth_l1_bwd, dstag_adam_step, dstag_gather_flat, dstag_multi_copy, dstag_tick,
              dstag_bias_transpose, dstag_bias_transpose_bwd):
        setattr(cls, f.__name__, f)
    return cls


FakeLib = _emu_optim(FakeLib)


def _iarr(ptr, n):
    return np.ctypeslib.as_array((C.c_int32 * int(n)).from_address(int(ptr)))


def _emu_sparse(cls):
    """Sparse Chebyshev path (csrc/sparse.cu), edge by edge in numpy."""
    def dstag_edge_probs(self, P, erow, col, tval, a, B, K, N, ld, E, stream):
        self.launches += 1
        Pm = _arr(P, B * K * N * ld).reshape(B, K * N, ld)
        a_ = _arr(a, B * E).reshape(B, E)
        a_[:] = _arr(tval, E)[None, :] * Pm[:, _iarr(erow, E), _iarr(col, E)]
        return 0

    def _rows(rowptr, KN):
        rp = _iarr(rowptr, KN + 1)
        return np.repeat(np.arange(KN), np.diff(rp))              # row id of every CSR slot

    def dstag_spmm_expand(self, src, a, rowptr, col, eid, dst, B, K, N, L, E, stream):
        self.launches += 1
        slot_row = _rows(rowptr, K * N)
        c = _iarr(col, E)
        ids = _iarr(eid, E) if eid else np.arange(E)
        S = _arr(src, B * N * L).reshape(B, N, L).astype(np.float64)
        A = _arr(a, B * E).reshape(B, E).astype(np.float64)
        out = np.zeros((B, K * N, L))
        for b in range(B):
            np.add.at(out[b], slot_row, A[b, ids][:, None] * S[b, c])
        _arr(dst, K * B * N * L).reshape(K, B, N, L)[:] = out.reshape(B, K, N, L).transpose(1, 0, 2, 3)
        return 0

    def dstag_spmm_reduce(self, src, a, rowptr, col, eid, dst, B, K, N, L, E, stream):
        self.launches += 1
        slot_row = _rows(rowptr, K * N)
        c = _iarr(col, E)
        ids = _iarr(eid, E) if eid else np.arange(E)
        S = _arr(src, K * B * N * L).reshape(K, B, N, L).astype(np.float64)
        A = _arr(a, B * E).reshape(B, E).astype(np.float64)
        out = np.zeros((B, N, L))
        k_of, n_of = slot_row // N, slot_row % N
        for b in range(B):
            np.add.at(out[b], n_of, A[b, ids][:, None] * S[k_of, b, c])
        _arr(dst, B * N * L).reshape(B, N, L)[:] = out
        return 0

    def dstag_edge_dots(self, dZ, X, a, rowptr, col, h, r, B, K, N, L, E, stream):
        self.launches += 1
        slot_row = _rows(rowptr, K * N)
        c = _iarr(col, E)
        D = _arr(dZ, K * B * N * L).reshape(K, B, N, L).astype(np.float64)
        Xm = _arr(X, B * N * L).reshape(B, N, L).astype(np.float64)
        A = _arr(a, B * E).reshape(B, E).astype(np.float64)
        k_of, j_of = slot_row // N, slot_row % N
        H = np.zeros((B, E))
        R = np.zeros((B, K * N))
        for b in range(B):
            H[b] = A[b] * np.einsum("el,el->e", D[k_of, b, j_of], Xm[b, c])
            np.add.at(R[b], slot_row, H[b])
        _arr(h, B * E).reshape(B, E)[:] = H
        _arr(r, B * K * N).reshape(B, K * N)[:] = R
        return 0

    def dstag_scale_keys(self, qk, r, rK, B, N, K, dk, stream):
        self.launches += 1
        Q = _arr(qk, B * N * 2 * K * dk).reshape(B, N, 2, K, dk)
        R = _arr(r, B * K * N).reshape(B, K, N)
        _arr(rK, B * N * K * dk).reshape(B, N, K, dk)[:] = R.transpose(0, 2, 1)[..., None] * Q[:, :, 1]
        return 0

    def dstag_center_queries(self, qk, sums, Qc, B, N, K, dk, stream):
        self.launches += 1
        Q = _arr(qk, B * N * 2 * K * dk).reshape(B, N, 2, K, dk)[:, :, 0].astype(np.float64)
        _arr(Qc, B * N * K * dk).reshape(B, N, K, dk)[:] = Q - Q.mean(axis=1, keepdims=True)
        return 0

    def dstag_dqk_combine(self, qk, Qc, h, r, O, W, rowptr, col, t_rowptr, t_col, t_eid, dqk, B, N, K, dk, E, alpha,
                          stream):
        self.launches += 1
        Q = _arr(qk, B * N * 2 * K * dk).reshape(B, N, 2, K, dk).astype(np.float64)
        Qcm = _arr(Qc, B * N * K * dk).reshape(B, N, K, dk).astype(np.float64)
        H = _arr(h, B * E).reshape(B, E).astype(np.float64)
        R = _arr(r, B * K * N).reshape(B, K, N).astype(np.float64)
        Om = _arr(O, B * N * K * dk).reshape(B, N, K, dk).astype(np.float64)
        Wm = _arr(W, B * N * K * dk).reshape(B, N, K, dk).astype(np.float64)
        row = _rows(rowptr, K * N); c = _iarr(col, E)
        trow = _rows(t_rowptr, K * N); tc = _iarr(t_col, E); te = _iarr(t_eid, E)
        out = _arr(dqk, B * N * 2 * K * dk).reshape(B, N, 2, K, dk)
        for b in range(B):
            dq = np.zeros((K * N, dk)); dk_ = np.zeros((K * N, dk))
            np.add.at(dq, trow, H[b, te][:, None] * Q[b, tc, 1, trow // N])
            np.add.at(dk_, row, H[b][:, None] * Qcm[b, c, row // N])
            dq = dq.reshape(K, N, dk).transpose(1, 0, 2) - Wm[b]
            dk_ = dk_.reshape(K, N, dk).transpose(1, 0, 2) - R[b].T[..., None] * Om[b]
            out[b, :, 0] = alpha * dq
            out[b, :, 1] = alpha * dk_
        return 0

    def dstag_dbias_sparse(self, P, h, r, pa_j, pa_i, pa_eid, dbiasT, B, K, N, ld, E, npa, stream):
        self.launches += 1
        Pm = _arr(P, B * K * N * ld).reshape(B, K, N, ld).astype(np.float64)
        H = _arr(h, B * E).reshape(B, E).astype(np.float64)
        R = _arr(r, B * K * N).reshape(B, K, N).astype(np.float64)
        pj, pi = _iarr(pa_j, npa), _iarr(pa_i, npa)
        pe = _iarr(pa_eid, K * npa).reshape(K, npa)
        out = _arr(dbiasT, K * N * ld).reshape(K, N, ld)
        for k in range(K):
            v = -(R[:, k, pj] * Pm[:, k, pj, pi]).sum(0)
            hit = pe[k] >= 0
            v[hit] += H[:, pe[k][hit]].sum(0)
            out[k, pj, pi] = v
        return 0

    for f in (dstag_edge_probs, dstag_spmm_expand, dstag_spmm_reduce, dstag_edge_dots, dstag_scale_keys,
              dstag_center_queries, dstag_dqk_combine, dstag_dbias_sparse):
        setattr(cls, f.__name__, f)
    return cls


FakeLib = _emu_sparse(FakeLib)


def install(monkeypatch):
    """Route the package's ctypes calls to the emulation (CPU tensors)."""
    from dstagnn_drought_b200 import _lib, contraction, modules, ops, sparse, step
    fake = FakeLib()
    monkeypatch.setattr(modules, "_require_cuda", lambda x: None)
    monkeypatch.setattr(_lib, "load", lambda: fake)
    monkeypatch.setattr(_lib, "stream_ptr", lambda t: None)
    monkeypatch.setattr(contraction, "_check", lambda t, name: None)
    monkeypatch.setattr(ops, "_check", lambda t, name: None)
    monkeypatch.setattr(step, "_check", lambda t, name: None)
    monkeypatch.setattr(sparse, "_check", lambda t, name: None)
    return fake
