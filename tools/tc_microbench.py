#!/usr/bin/env python
"""Time individual contractions of the PEMS07 training step (CUDA events, L2 flushed between
repeats).  usage: tc_microbench.py [--batch 64] [--precision tf32x3] [--only cheb]"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dstagnn_drought_b200 import _lib, contraction as ct  # noqa: E402
from dstagnn_drought_b200._lib import ACC_ATOMIC  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--precision", default="tf32x3")
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, N, K, F, T, D, H, C, dk = a.batch, 883, 3, 32, 12, 512, 4, 32, 32
    ld = (N + 31) // 32 * 32
    FT, X = F * T, B * N
    prec = ct._PRECISION[a.precision]
    r = lambda *s: torch.randn(*s, device=dev)  # noqa: E731
    rp = lambda *s: ct.empty_rows(s, dev).normal_()  # noqa: E731   (row-padded like the model's node-axis tensors)
    P = torch.zeros(B, K, N, ld, device=dev); P[..., :N] = torch.softmax(r(B, K, N, N), -1)
    Tk = torch.zeros(K, N, ld, device=dev); Tk[..., :N] = r(K, N, N)
    x3, Z, dS = r(B, N, FT), r(K, B, N, FT), torch.zeros(B, K, N, ld, device=dev)
    qk = r(B, N, 2, K, dk)
    E, Wqkv = rp(B * FT, N), rp(3 * H * dk, N)
    O, Wpre, Zs = rp(B, F, T, N), r(D, T, 1, F), r(B, N, D)
    g = r(X, T, C)                                     # channels-last graph-conv output
    theta = r(K, F, C)
    cat, dtc = r(X * C, 3 * T - 12), r(X * C, T)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    cases = [
        # (the model feeds TP = T_k (.) P, written by the map kernel, and applies T_k in the softmax backward)
        ("cheb_fwd", "bkji,bie->kbje", P[..., :N], x3, {}),
        ("cheb_bwd_dx", "bkji,kbje->bie", P[..., :N], Z, {}),
        ("cheb_bwd_dmap", "kbje,bie->bkji", Z, x3, dict(out=dS[..., :N])),
        ("cheb_fwd_amul", "bkji,bie->kbje", P[..., :N], x3, dict(a_mul=("kji", Tk[..., :N]))),
        # orientation experiment: the row-contiguous operand first (MN-major A by TMA through TMEM, K-major B by TMA)
        ("cheb_fwd_swap", "bie,bkji->kbje", x3, P[..., :N], dict(noswap=True)),
        ("qkv_dx_swap", "on,ro->rn", Wqkv, r(B * FT, 3 * H * dk), dict(out=rp(B * FT, N), noswap=True)),
        ("sat_scores", "bjkd,bikd->bkji", qk[:, :, 1], qk[:, :, 0], dict(out=dS[..., :N], alpha=1 / math.sqrt(dk))),
        ("sat_dq", "bkji,bjkd->bikd", dS[..., :N], qk[:, :, 1], {}),
        ("sat_dk", "bkji,bikd->bjkd", dS[..., :N], qk[:, :, 0], {}),
        # the sparse path's two dense products with the softmax maps (chains of 64 terms, dstagnn_drought_b200/sparse.py)
        ("sat_PQ", "bkji,bikd->bjkd", P[..., :N], r(B, N, K, dk), dict(chain=64)),
        ("sat_PtrK", "bkji,bjkd->bikd", P[..., :N], r(B, N, K, dk), dict(chain=64)),
        ("qkv", "rn,on->ro", E, Wqkv, {}),
        ("qkv_dx", "ro,on->rn", r(B * FT, 3 * H * dk), Wqkv, dict(out=rp(B * FT, N))),
        ("qkv_dw", "ro,rn->on", r(B * FT, 3 * H * dk), E, dict(out=rp(3 * H * dk, N))),
        ("pre_conv", "bftn,dtf->bnd", O, Wpre[:, :, 0, :], {}),
        ("pre_conv_dx", "bnd,dtf->bftn", Zs, Wpre[:, :, 0, :], dict(out=rp(B, F, T, N))),
        ("pre_conv_dw", "bftn,bnd->dtf", O, Zs, {}),
        ("sat_proj", "rd,od->ro", r(X, D), r(2 * K * dk, D), {}),
        ("theta_mix", "kxft,kfc->xtc", Z.view(K, X, F, T), theta, dict(relu=True)),
        ("theta_dz", "xtc,kfc->kxft", g, theta, {}),
        ("theta_dth", "kxft,xtc->kfc", Z.view(K, X, F, T), g, {}),
        ("fcmy_dw", "rt,rs->ts", dtc, cat, {}),
    ]
    # epilogue calibration: same GEMM (M=96, N=X*T, K=32), three output layouts
    th2, g2 = r(96, 32), r(X * T, 32)
    obuf = torch.empty(96 * X * T + 8, device=dev)
    cases += [
        ("epi_vec", "mk,nk->mn", th2, g2, dict(out=obuf[:96 * X * T].view(96, X * T))),            # (a) float4 rows
        ("epi_unal", "mk,nk->mn", th2, g2, dict(out=obuf[1:96 * X * T + 1].view(96, X * T))),       # (b) unaligned rows
        ("epi_2lvl", "mk,xtk->xmt", th2, g2.view(X, T, 32), dict(out=obuf[:96 * X * T].view(X, 96, T))),  # (b) runs of T
    ]
    for ks in (3, 5, 7):
        tp, pad = T - ks + 1, ks - 1
        tpad = tp + 2 * pad
        gv = g.as_strided((X, tp, ks * C), (T * C, C, 1))
        dup = r(X, tpad, 2 * C)
        dview = dup.as_strided((X, T, ks * 2 * C), (tpad * 2 * C, 2 * C, 1))
        cases += [
            (f"gtu{ks}_fwd", "xtk,ok->xto", gv, r(2 * C, ks * C), {}),
            (f"gtu{ks}_dw", "xto,xtk->ok", dup[:, pad:pad + tp, :], gv, {}),
            (f"gtu{ks}_dx", "xtk,ck->xtc", dview, r(C, ks * 2 * C), dict(split_k=1)),
        ]
    print(f"{'case':14s} {'engine':6s} {'ms':>9s} {'TFLOP/s':>9s}   (precision {a.precision}, batch {B})")
    for name, spec, A, Bm, kw in cases:
        if a.only and not any(o == name or (o.endswith("*") and name.startswith(o[:-1])) for o in a.only.split(",")):
            continue
        kw = dict(kw)
        noswap = kw.pop("noswap", False)
        out = kw.get("out")
        if out is None:
            out = torch.empty(ct.output_shape(spec, A, Bm), device=dev)
            kw["out"] = out
        import ctypes
        gd = ct.build_descriptor(spec, A, Bm, out, **{k: v for k, v in kw.items() if k in ("a_mul", "c_add", "c_mul")},
                                 precision=prec)
        tc = bool(_lib.load().dstag_gemm_uses_tensor_cores(ctypes.byref(gd)))
        flops = 2.0 * gd.M * gd.N * gd.K * gd.batch
        kw.setdefault("split_k", None)
        ts = []
        for _ in range(a.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if noswap:
                saved = ct._prefers_swap
                ct._prefers_swap = lambda *a, **k: False
                ct._swap_cache.clear()
            ct.contract_raw(spec, A, Bm, precision=prec, **kw)
            if noswap:
                ct._prefers_swap = saved
                ct._swap_cache.clear()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print(f"{name:14s} {'tc' if tc else 'simt':6s} {ms:9.3f} {flops / ms / 1e9:9.2f}   M={gd.M} N={gd.N} K={gd.K} b={gd.batch}")


if __name__ == "__main__":
    main()
