#!/usr/bin/env python
"""Measure the dense TF32 tensor-core throughput of this box next to the bf16 figure of MEASURED_PEAKS.json:
cuBLAS TF32 and bf16 (torch.matmul, 8192^3) and this library's engine on the same shape in its single-pass
`tf32` and fp32-grade `tf32x3` modes.  The 3xTF32 ceiling quoted in DESIGN.md (measured TF32 / 3) comes from here.

    python tools/tf32_peak.py > profiles/r02_tf32_peak.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dstagnn_drought_b200 import _lib  # noqa: E402
from dstagnn_drought_b200.contraction import contract_raw  # noqa: E402


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    n = 8192
    dev = torch.device("cuda:0")
    A = torch.randn(n, n, device=dev)
    B = torch.randn(n, n, device=dev)
    C = torch.empty(n, n, device=dev)
    flops = 2.0 * n ** 3
    out = {"shape": f"{n}^3", "unit": "TFLOP/s"}
    torch.backends.cuda.matmul.allow_tf32 = True
    out["cublas_tf32"] = flops / timed(lambda: torch.matmul(A, B, out=C)) / 1e9
    torch.backends.cuda.matmul.allow_tf32 = False
    out["cublas_fp32"] = flops / timed(lambda: torch.matmul(A, B, out=C), reps=3, warm=1) / 1e9
    Ah, Bh = A.bfloat16(), B.bfloat16()
    Ch = torch.empty(n, n, device=dev, dtype=torch.bfloat16)
    out["cublas_bf16"] = flops / timed(lambda: torch.matmul(Ah, Bh, out=Ch)) / 1e9
    _lib.load()
    for mode in ("tf32", "tf32x3"):
        prec = {"tf32": _lib.PREC_TF32, "tf32x3": _lib.PREC_TF32X3}[mode]
        # K-major operands on both sides ("mk,nk->mn"): the TMA-fed path of the engine
        out[f"engine_{mode}"] = flops / timed(lambda: contract_raw("mk,nk->mn", A, B, out=C, precision=prec)) / 1e9
    out["tf32x3_ceiling_from_cublas_tf32"] = out["cublas_tf32"] / 3.0
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
