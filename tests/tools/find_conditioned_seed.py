#!/usr/bin/env python
"""Find, with the fp64 CPU oracle ONLY, a well-conditioned data seed for each oracle-only parity case of
tests/test_gpu_parity.py (ORACLE_CASES).

    python tests/tools/find_conditioned_seed.py [case ...]

"Well conditioned" = the fp64 gradients move by < 5e-5 (relative to each tensor's maximum) when all weights
and the input are perturbed by 1e-6: no ReLU pre-activation sits within rounding distance of its kink, so the
comparison tests the kernels and not the position of one rounding error (tests/helpers.conditioned_case).
The seeds found are written into ORACLE_CASES by hand; the test verifies each one again before using it.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from helpers import conditioned_case, oracle_fp32_grad_err  # noqa: E402
from oracle import dstagnn_oracle as O  # noqa: E402
from test_gpu_parity import COND_THR, ORACLE_CASES  # noqa: E402


def main():
    for name in sys.argv[1:] or sorted(ORACLE_CASES):
        shape = ORACLE_CASES[name][0]
        N, T, K, H, D, nb, B = (shape[k] for k in ("N", "T", "K", "H", "D", "nb", "B"))
        t0 = time.time()
        net = O.build_oracle(N, T, K, H, D, nb, dtype=torch.float64, faithful=False)
        x, tgt, _, g64, seed = conditioned_case(net, B, N, T, verbose=True, thr=COND_THR)
        e = oracle_fp32_grad_err(net, x, tgt, g64)
        print(f"{name}: data seed {seed}  ({time.time() - t0:.1f} s; fp32 oracle's worst gradient error "
              f"{max(e.values()):.2e} at {max(e, key=e.get)})", flush=True)


if __name__ == "__main__":
    main()
