#!/usr/bin/env python
"""Stress the tcgen05 engine on the small / ragged shapes the golden nets produce; repeat each call
to expose races (result must be bit-identical across repeats when no atomics are involved)."""
import os, sys, ctypes
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from dstagnn_drought_b200 import _lib, contraction as ct
from helpers import rel_err
dev = torch.device("cuda:0")
torch.manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev)
P = ct._PRECISION["tf32x3"]
cases = []
X, C, T = 99, 32, 12
g = r(X, T, C)                                            # channels-last graph-conv output
for ks in (3, 5, 7):
    Tp, pad = T - ks + 1, ks - 1
    gv = g.as_strided((X, Tp, ks * C), (T * C, C, 1))     # overlapping conv windows
    dup = r(X, Tp + 2 * pad, 2 * C)
    cases += [(f"conv_fwd{ks}", "xtk,ok->xto", gv, r(2 * C, ks * C), {}),
              (f"conv_dw{ks}", "xto,xtk->ok", dup[:, pad:pad + Tp, :], gv, {}),
              (f"conv_dx{ks}", "xtk,ck->xtc", dup.as_strided((X, T, ks * 2 * C), ((Tp + 2 * pad) * 2 * C, 2 * C, 1)),
               r(C, ks * 2 * C), {})]
N, D, Bz, K, FT = 33, 32, 3, 3, 384
O, Wp, dZ = ct.empty_rows((Bz, 32, 12, N), dev).normal_(), r(D, 12, 1, 32), r(Bz, N, D)
cases += [("pre_conv", "bftn,dtf->bnd", O, Wp[:, :, 0, :], {}),
          ("pre_dx", "bnd,dtf->bftn", dZ, Wp[:, :, 0, :], {}),
          ("pre_dw", "bftn,bnd->dtf", O, dZ, {})]
E, Wq = ct.empty_rows((Bz * FT, N), dev).normal_(), ct.empty_rows((288, N), dev).normal_()
cases += [("qkv", "rn,on->ro", E, Wq, {}), ("qkv_dx", "ro,on->rn", r(Bz * FT, 288), Wq, {}),
          ("qkv_dw", "rn,ro->on", E, r(Bz * FT, 288), {})]
ld = 64
Pm = torch.zeros(Bz, K, N, ld, device=dev); Pm[..., :N] = torch.softmax(r(Bz, K, N, N), -1)
Tk = torch.zeros(K, N, ld, device=dev); Tk[..., :N] = r(K, N, N)
x3, Z = r(Bz, N, FT), r(K, Bz, N, FT)
cases += [("cheb_fwd", "bkji,bie->kbje", Pm[..., :N], x3, dict(a_mul=("kji", Tk[..., :N]))),
          ("cheb_dx", "bkji,kbje->bie", Pm[..., :N], Z, dict(a_mul=("kji", Tk[..., :N]))),
          ("cheb_dmap", "kbje,bie->bkji", Z, x3, dict(c_mul=("kji", Tk[..., :N])))]
th = r(K, 32, 32)
cases += [("theta", "kxtf,kfc->xtc", Z.view(K, Bz * N, 12, 32), th, dict(relu=True)),
          ("theta_dz", "xtc,kfc->kxtf", r(Bz * N, 12, 32), th, {}),
          ("theta_dth", "kxtf,xtc->kfc", Z.view(K, Bz * N, 12, 32), r(Bz * N, 12, 32), {})]
cat, Wf = r(Bz * N, 24, 32), r(12, 24)
cases += [("fcmy", "xsc,ts->xtc", cat, Wf, {}), ("fcmy_dx", "xtc,ts->xsc", r(Bz * N, 12, 32), Wf, {})]
fin, Wfin = r(Bz * N, 12, 32), r(128, 12, 32)
cases += [("final", "xtc,otc->xo", fin, Wfin, {}), ("final_dx", "xo,otc->xtc", r(Bz * N, 128), Wfin, {}),
          ("final_dw", "xo,xtc->otc", r(Bz * N, 128), fin, {})]
SE, Ws = r(Bz * N, D), r(192, D)
cases += [("sat", "rd,od->ro", SE, Ws, {}), ("sat_dx", "ro,od->rd", r(Bz * N, 192), Ws, {}), ("sat_dw", "rd,ro->od", SE, r(Bz * N, 192), {})]
for name, spec, A, B, kw in cases:
    letters = spec.replace(" ", "").split("->")[0].split(",")
    ein = spec
    ops = [A.double(), B.double()]
    if "a_mul" in kw:
        ein = f"{letters[0]},{kw['a_mul'][0]},{letters[1]}->{spec.split('->')[1]}"; ops = [A.double(), kw["a_mul"][1].double(), B.double()]
    ref = torch.einsum(ein, *ops)
    if "c_mul" in kw:
        ref = ref * kw["c_mul"][1].double()
    if kw.get("relu"):
        ref = torch.relu(ref)
    outs = []
    for rep in range(6):
        out = torch.full(ct.output_shape(spec, A, B), float("nan"), device=dev)
        ct.contract_raw(spec, A, B, out=out, precision=P, split_k=1, **kw)
        outs.append(out.clone())
    gd = ct.build_descriptor(spec, A, B, outs[0], precision=P, **{k: v for k, v in kw.items() if k in ("a_mul", "c_mul")})
    tc = bool(_lib.load().dstag_gemm_uses_tensor_cores(ctypes.byref(gd)))
    errs = [rel_err(o, ref) for o in outs]
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    flag = "" if max(errs) < 2e-5 and same else "   <<<<<<"
    print(f"{name:10s} {'tc' if tc else 'simt':4s} M={gd.M:6d} N={gd.N:4d} K={gd.K:6d} b={gd.batch:3d} max_err={max(errs):.2e} min_err={min(errs):.2e} deterministic={same}{flag}")
