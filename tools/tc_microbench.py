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
    P = torch.zeros(B, K, N, ld, device=dev); P[..., :N] = torch.softmax(r(B, K, N, N), -1)
    Tk = torch.zeros(K, N, ld, device=dev); Tk[..., :N] = r(K, N, N)
    x3, Z, dS = r(B, N, FT), r(K, B, N, FT), torch.zeros(B, K, N, ld, device=dev)
    qk = r(B, N, 2, K, dk)
    E, Wqkv = r(B * FT, N), r(3 * H * dk, N)
    O, Wpre, Zs = r(B, F, T, N), r(D, T, 1, F), r(B, N, D)
    g3 = r(X, C, T); w5 = r(2 * C, C, 5); u5 = r(X, 2 * C, T - 4)
    cat, Wf = r(X * C, 3 * T - 12), r(T, 3 * T - 12)
    theta = r(K, F, C)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    cases = [
        ("cheb_fwd", "bkji,bie->kbje", P[..., :N], x3, dict(a_mul=("kji", Tk[..., :N]))),
        ("cheb_bwd_dx", "bkji,kbje->bie", P[..., :N], Z, dict(a_mul=("kji", Tk[..., :N]))),
        ("cheb_bwd_dmap", "kbje,bie->bkji", Z, x3, dict(out=dS[..., :N], c_mul=("kji", Tk[..., :N]))),
        ("sat_scores", "bjkd,bikd->bkji", qk[:, :, 1], qk[:, :, 0], dict(out=dS[..., :N], alpha=1 / math.sqrt(dk))),
        ("sat_dq", "bkji,bjkd->bikd", dS[..., :N], qk[:, :, 1], {}),
        ("sat_dk", "bkji,bikd->bjkd", dS[..., :N], qk[:, :, 0], {}),
        ("qkv", "rn,on->ro", E, Wqkv, {}),
        ("qkv_dx", "ro,on->rn", r(B * FT, 3 * H * dk), Wqkv, {}),
        ("qkv_dw", "rn,ro->on", E, r(B * FT, 3 * H * dk), {}),
        ("pre_conv", "bftn,dtf->bnd", O, Wpre[:, :, 0, :], {}),
        ("pre_conv_dx", "bnd,dtf->bftn", Zs, Wpre[:, :, 0, :], {}),
        ("pre_conv_dw", "bftn,bnd->dtf", O, Zs, {}),
        ("theta_mix", "kxft,kfc->xct", Z.view(K, X, F, T), theta, dict(relu=True)),
        ("theta_dz", "xct,kfc->kxft", g3, theta, {}),
        ("theta_dth", "kxft,xct->kfc", Z.view(K, X, F, T), g3, {}),
        ("gtu5_fwd", "xctd,ocd->xot", g3.unfold(2, 5, 1), w5, {}),
        ("gtu5_dw", "xot,xctd->ocd", u5, g3.unfold(2, 5, 1), {}),
        ("gtu5_dx", "xot,ocd->xctd", u5, w5, dict(out=torch.zeros_like(g3).unfold(2, 5, 1), accumulate=ACC_ATOMIC, split_k=1)),
        ("fcmy", "rs,ts->rt", cat, Wf, {}),
        ("fcmy_dx", "rt,ts->rs", r(X * C, T), Wf, {}),
        ("fcmy_dw", "rs,rt->ts", cat, r(X * C, T), {}),
    ]
    print(f"{'case':14s} {'engine':6s} {'ms':>9s} {'TFLOP/s':>9s}   (precision {a.precision}, batch {B})")
    for name, spec, A, Bm, kw in cases:
        if a.only and a.only not in name:
            continue
        kw = dict(kw)
        out = kw.get("out")
        if out is None:
            out = torch.empty(ct.output_shape(spec, A, Bm), device=dev)
            kw["out"] = out
        import ctypes
        gd = ct.build_descriptor(spec, A, Bm, out, **{k: v for k, v in kw.items() if k in ("a_mul", "c_add", "c_mul")},
                                 precision=prec)
        tc = bool(_lib.load().dstag_gemm_uses_tensor_cores(ctypes.byref(gd)))
        flops = 2.0 * gd.M * gd.N * gd.K * gd.batch
        ts = []
        for _ in range(a.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ct.contract_raw(spec, A, Bm, precision=prec, **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print(f"{name:14s} {'tc' if tc else 'simt':6s} {ms:9.3f} {flops / ms / 1e9:9.2f}   M={gd.M} N={gd.N} K={gd.K} b={gd.batch}")


if __name__ == "__main__":
    main()
