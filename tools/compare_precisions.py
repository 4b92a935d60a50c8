#!/usr/bin/env python
"""Run the same net / batch with the exact-fp32 SIMT path and with a tensor-core precision and report,
per tensor, max|diff|/max|ref| for block outputs, block-output gradients and parameter gradients."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dstagnn_drought_b200 as D
from dstagnn_drought_b200 import make_model
from helpers import load_golden, rel_err

name = sys.argv[1] if len(sys.argv) > 1 else "g_n33"
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
d, cfg, w, g, cheb = load_golden(name)
dev = torch.device("cuda:0")


def run(precision):
    D.set_precision(precision)
    net = make_model(dev, 1, cfg["nb"], 1, cfg["K"], 32, 32, 1, torch.from_numpy(d["adj_tmd"]), torch.from_numpy(d["adj_pa"]),
                     torch.from_numpy(d["adj_tmd"]), cfg["P"], cfg["T"], cfg["N"], cfg["D"], 32, 32, cfg["H"])
    net.load_state_dict(w); net.eval()
    rec = {}
    def fhook(i):
        def h(m, inp, out):
            rec[f"fwd.blk{i}"] = out[0].detach().clone()
            out[0].register_hook(lambda gr: rec.__setitem__(f"dout.blk{i}", gr.detach().clone()))
            out[1].register_hook(lambda gr: rec.__setitem__(f"dlogits.blk{i}", gr.detach().clone()) if gr is not None else None)
        return h
    for i, b in enumerate(net.BlockList):
        b.register_forward_hook(fhook(i))
    y = net(torch.from_numpy(d["x"]).to(dev))
    torch.nn.SmoothL1Loss()(y, torch.from_numpy(d["target"]).to(dev)).backward()
    rec["y"] = y.detach().clone()
    for k, p in net.named_parameters():
        if p.grad is not None:
            rec["g." + k] = p.grad.detach().clone()
    return rec

a = run("fp32")
b = run(prec)
rows = sorted(((rel_err(b[k], a[k]), k, float(a[k].abs().max())) for k in a if k in b), reverse=True)
for e, k, mx in rows[:30]:
    gold = f" vs golden fp32={rel_err(a[k], g[k[2:]]):.1e} {prec}={rel_err(b[k], g[k[2:]]):.1e}" if k.startswith("g.") else ""
    print(f"{e:9.2e}  max|ref|={mx:9.2e}  {k}{gold}")
