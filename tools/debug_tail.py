import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dstagnn_drought_b200 as D
from dstagnn_drought_b200 import make_model, ops
from helpers import load_golden, rel_err
d, cfg, w, g, cheb = load_golden("g_n33")
dev = torch.device("cuda:0")
rec = {}
cur = {"mode": None, "n": 0}
of, ob = ops.BlockTail.forward, ops.BlockTail.backward
def fwd(ctx, tc, g_, r, gamma, beta, p, seed):
    out = of(ctx, tc, g_, r, gamma, beta, p, seed)
    i = cur["n"]; cur["n"] += 1
    ctx.dbg = i
    rec[(cur["mode"], f"blk{i}.tc")] = tc.detach().clone(); rec[(cur["mode"], f"blk{i}.r")] = r.detach().clone()
    rec[(cur["mode"], f"blk{i}.out")] = out.detach().clone()
    return out
def bwd(ctx, dout):
    res = ob(ctx, dout)
    i = ctx.dbg
    rec[(cur["mode"], f"blk{i}.dout")] = dout.detach().clone()
    rec[(cur["mode"], f"blk{i}.dtc")] = res[0].detach().clone()
    rec[(cur["mode"], f"blk{i}.dr")] = res[2].detach().clone()
    rec[(cur["mode"], f"blk{i}.dgamma")] = res[3].detach().clone()
    return res
ops.BlockTail.forward = staticmethod(fwd); ops.BlockTail.backward = staticmethod(bwd)
for mode in ("fp32", "tf32x3"):
    D.set_precision(mode); cur["mode"] = mode; cur["n"] = 0
    net = make_model(dev, 1, cfg["nb"], 1, cfg["K"], 32, 32, 1, torch.from_numpy(d["adj_tmd"]), torch.from_numpy(d["adj_pa"]),
                     torch.from_numpy(d["adj_tmd"]), cfg["P"], cfg["T"], cfg["N"], cfg["D"], 32, 32, cfg["H"])
    net.load_state_dict(w); net.eval()
    y = net(torch.from_numpy(d["x"]).to(dev))
    torch.nn.SmoothL1Loss()(y, torch.from_numpy(d["target"]).to(dev)).backward()
keys = sorted(k for m, k in rec if m == "fp32")
for k in keys:
    a, b = rec[("fp32", k)], rec[("tf32x3", k)]
    diff = (a - b).abs()
    print(f"{k:14s} rel={rel_err(b, a):.2e} max|ref|={float(a.abs().max()):.2e} n_big_diff={(diff > 1e-3 * a.abs().max()).sum().item()} of {a.numel()}")
a, b = rec[("fp32", "blk0.tc")], rec[("tf32x3", "blk0.tc")]
print("tc sign flips:", ((a > 0) != (b > 0)).sum().item(), "min|tc|", float(a.abs().min()))
