#!/usr/bin/env python
"""Time the GTU stack (forward, backward) at the PEMS07 B=64 shape: fused tcgen05 kernels (csrc/gtu_tc.cu)
against the contraction-engine + gate-kernel path.  CUDA events, L2 flushed between repeats."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dstagnn_drought_b200 import _lib  # noqa: E402
from dstagnn_drought_b200.ops import GtuStack  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    X, T, Cc = int(os.environ.get("X", 64 * 883)), int(os.environ.get("T", 12)), 32
    torch.manual_seed(0)
    g = torch.randn(X, T, Cc, device=dev, requires_grad=True)
    ws = [(torch.randn(2 * Cc, Cc, 1, ks, device=dev) * 0.2).requires_grad_() for ks in (3, 5, 7)]
    bs = [torch.randn(2 * Cc, device=dev, requires_grad=True) for _ in range(3)]
    dout = torch.randn(X, 3 * T - 12, Cc, device=dev)
    args = [g, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2]]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for mode in ("gemm", "fused", "gemm", "fused"):
        os.environ["DSTAG_GTU"] = mode
        tf, tb = [], []
        for rep in range(6):
            flush.zero_()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            out = GtuStack.apply(*args)
            e[1].record()
            out.backward(dout)
            e[2].record()
            torch.cuda.synchronize()
            if rep:
                tf.append(e[0].elapsed_time(e[1])), tb.append(e[1].elapsed_time(e[2]))
            for a in args:
                a.grad = None
        tf.sort(), tb.sort()
        print(f"{mode:6s} forward {tf[len(tf) // 2]:.3f} ms  backward {tb[len(tb) // 2]:.3f} ms   (X={X} T={T})", flush=True)
    # per entry point
    _lib.prof_enable(True)
    os.environ["DSTAG_GTU"] = "fused"
    for _ in range(3):
        out = GtuStack.apply(*args)
        out.backward(dout)
    torch.cuda.synchronize()
    for k, v in sorted(_lib.prof_results().items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:32s} {v[0] / 3:5.1f} calls {v[1] / 3:8.3f} ms  {v[2] / max(v[1], 1e-9) / 1e9:7.1f} TF/s {v[3] / max(v[1], 1e-9) / 1e6:8.1f} GB/s")


if __name__ == "__main__":
    main()
