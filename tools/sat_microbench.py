"""Times dstag_sat_softmax_fwd alone (spatial attention maps, forward) at a dataset's graph size and checks
it against a float64 torch evaluation.  The kernel variant is chosen by the library's environment switches
(DSTAG_SAT_MMA=0 -> SIMT variants), so A/B runs are separate processes:

    python tools/sat_microbench.py --N 883 --B 64 --K 3
"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dstagnn_drought_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=883)
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--K", type=int, default=3)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--ponly", action="store_true", help="P only (no T_k (.) P output): what the sparse Chebyshev path calls")
    ap.add_argument("--ld32", action="store_true", help="leading dimension padded to 32 floats as in the model")
    a = ap.parse_args()
    B, N, K, dk = a.B, a.N, a.K, 32
    ld = (N + 31) // 32 * 32 if a.ld32 else (N + 3) // 4 * 4
    dev = torch.device("cuda:0")
    lib = _lib.load()
    torch.manual_seed(0)
    qk = torch.randn(B, N, 2, K, dk, device=dev)
    bias = torch.zeros(K, N, ld, device=dev)
    bias[..., :N] = torch.randn(K, N, N, device=dev)
    tk = torch.zeros(K, N, ld, device=dev)
    tk[..., :N] = torch.randn(K, N, N, device=dev)
    P = torch.full((B, K, N, ld), float("nan"), device=dev)
    TP = torch.full((B, K, N, ld), float("nan"), device=dev)
    st = _lib.stream_ptr(P)

    def run():
        _lib.check(lib.dstag_sat_softmax_fwd(qk.data_ptr(), bias.data_ptr(), None if a.ponly else tk.data_ptr(), P.data_ptr(),
                                             None if a.ponly else TP.data_ptr(), B, N, K, dk, ld, st), "dstag_sat_softmax_fwd")
    run()
    torch.cuda.synchronize()
    nb = min(B, 4)
    s = torch.einsum("bjkd,bikd->bkji", qk[:nb, :, 1].double(), qk[:nb, :, 0].double()) / math.sqrt(dk) + bias[..., :N].double()
    ref = torch.softmax(s, dim=-1)
    errP = float((P[:nb, ..., :N].double() - ref).abs().max())
    errT = 0.0 if a.ponly else float((TP[:nb, ..., :N].double() - ref * tk[..., :N].double()).abs().max())
    pad_ok = bool((P[..., N:] == 0).all()) and (a.ponly or bool((TP[..., N:] == 0).all()))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    gb = (1.0 if a.ponly else 2.0) * B * K * N * ld * 4 / 1e9
    print(f"sat_softmax_fwd B={B} N={N} K={K}: {ms:.3f} ms  ({gb / ms * 1e3:.0f} GB/s of P+TP writes)  "
          f"max|dP|={errP:.2e} max|dTP|={errT:.2e} pad_zero={pad_ok} "
          f"[MMA={os.environ.get('DSTAG_SAT_MMA', '1')}]")


if __name__ == "__main__":
    main()
