#!/usr/bin/env python
"""Debug aid: run a golden fixture through the CUDA path and, for EVERY contraction, also run the
exact-fp32 SIMT kernel on the same operands; report the calls whose tensor-core result differs."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dstagnn_drought_b200 import _lib, contraction as ct, make_model  # noqa: E402
from helpers import load_golden, rel_err  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "g_n33"
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 2e-5
orig = ct.contract_raw
bad = []
stats = []


def checked(spec, A, B, **kw):
    if kw.get("accumulate", 0) != 0:
        return orig(spec, A, B, **kw)
    out = kw.get("out")
    pre = out.clone() if out is not None else None
    res = orig(spec, A, B, **kw)
    kw2 = dict(kw)
    kw2["precision"] = _lib.PREC_FP32
    kw2.pop("tag", None)
    if pre is not None:
        ref_out = torch.empty_like(pre)          # preserves strides of dense tensors
        ref_out = pre.clone()
        kw2["out"] = ref_out
    ref = orig(spec, A, B, **kw2)
    e = rel_err(res, ref)
    stats.append((e, spec, tuple(A.shape), tuple(B.shape)))
    if e > thr:
        bad.append((spec, tuple(A.shape), tuple(B.shape), e))
        print(f"MISMATCH {spec:22s} A{tuple(A.shape)} B{tuple(B.shape)} err={e:.3e} kw={[k for k in kw if k != 'out']}")
    return res


ct.contract_raw = checked
import dstagnn_drought_b200.ops as ops  # noqa: E402
ops.contract_raw = checked

d, cfg, w, g, cheb = load_golden(name)
dev = torch.device("cuda:0")
net = make_model(dev, 1, cfg["nb"], 1, cfg["K"], 32, 32, 1, torch.from_numpy(d["adj_tmd"]), torch.from_numpy(d["adj_pa"]),
                 torch.from_numpy(d["adj_tmd"]), cfg["P"], cfg["T"], cfg["N"], cfg["D"], 32, 32, cfg["H"])
net.load_state_dict(w)
net.eval()
y = net(torch.from_numpy(d["x"]).to(dev))
torch.nn.SmoothL1Loss()(y, torch.from_numpy(d["target"]).to(dev)).backward()
torch.cuda.synchronize()
print("pred err", rel_err(y.detach(), d["y64"]))
worst = sorted(((rel_err(p.grad, g[k]), k) for k, p in net.named_parameters() if p.grad is not None), reverse=True)[:5]
print("worst grads", worst)
print("mismatching contractions:", len(bad))
for e, spec, sa, sb in sorted(stats, reverse=True)[:12]:
    print(f"  {e:.2e} {spec:24s} {sa} {sb}")
