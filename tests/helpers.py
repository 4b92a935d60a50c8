"""Shared test helpers: golden-fixture loading and the per-tensor error metric."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    N, T, K, H, D, nb, B, P = (int(v) for v in d["cfg"])
    cfg = dict(N=N, T=T, K=K, H=H, D=D, nb=nb, B=B, P=P)
    w = {k[2:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("w.")}
    g = {k[4:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("g64.")}
    cheb = [torch.from_numpy(d[f"cheb.{k}"]) for k in range(max(K, 2)) if f"cheb.{k}" in d.files]
    return d, cfg, w, g, cheb


def rel_err(a, ref):
    """max-abs-err / max-abs-ref (SURVEY.md Q16 metric)."""
    a = torch.as_tensor(a).double().cpu()
    ref = torch.as_tensor(ref).double().cpu()
    assert a.shape == ref.shape, (a.shape, ref.shape)
    den = ref.abs().max().item()
    return (a - ref).abs().max().item() / (den if den > 0 else 1.0)


# ------------------------------------------------------------------------------------------------
# oracle-side test cases at shapes the reference fixtures do not cover
# ------------------------------------------------------------------------------------------------
def _oracle_run(onet, x, tgt):
    for v in onet.named().values():
        v.grad = None
    y = onet(x)
    torch.nn.SmoothL1Loss()(y, tgt).backward()
    return y.detach(), {k: (v.grad.detach().clone() if v.grad is not None else None) for k, v in onet.named().items()}


def conditioned_case(onet64, B, N, T, P=12, seeds=range(3, 64), thr=2e-5, trials=2, verbose=False):
    """Inputs for a parity test at an oracle-only shape, chosen WITHOUT looking at the implementation
    under test: the first data seed whose fp64 oracle gradients move by less than ``thr`` (relative to each
    tensor's maximum, per 1e-6 of noise) when every weight and the input are perturbed by ~1e-6.

    Why: the net has ReLUs.  A pre-activation within fp32 rounding of zero lands on either side of the kink
    in ANY fp32 implementation (torch's own CPU kernels included) and moves a handful of gradients by 1e-3;
    such an input tests the position of one rounding error, not the kernels.  tests/golden/make_golden.py
    conditions the reference fixtures the same way.  Returns (x, target, y64, grads64, data_seed)."""
    import copy
    for data_seed in seeds:
        g = torch.Generator().manual_seed(data_seed)
        x = torch.randn(B, N, 1, T, generator=g)
        tgt = torch.randn(B, N, P, generator=g)
        y64, g64 = _oracle_run(onet64, x.double(), tgt.double())
        cond = 0.0
        for trial in range(trials):
            eps = (1e-6, 3e-6, 3e-7, 1e-7)[trial % 4]
            netp = copy.deepcopy(onet64)
            with torch.no_grad():
                for prm in netp.parameters():
                    prm.mul_(1 + eps * torch.randn(prm.shape, generator=g, dtype=torch.float64))
            xp = x.double() * (1 + eps * torch.randn(x.shape, generator=g, dtype=torch.float64))
            _, gp = _oracle_run(netp, xp, tgt.double())
            cond = max(cond, max(float((gp[k] - g64[k]).abs().max() / g64[k].abs().max()) / (eps / 1e-6)
                                 for k in g64 if g64[k] is not None))
            if cond >= thr:
                break
        if verbose:
            print(f"  data seed {data_seed}: gradient sensitivity per 1e-6 noise = {cond:.2e}")
        if cond < thr:
            return x, tgt, y64, g64, data_seed
    raise RuntimeError("no well-conditioned data seed found")


def oracle_fp32_grad_err(onet64, x, tgt, g64):
    """The error the SAME algorithm makes in plain fp32 on the CPU against its fp64 run, per gradient
    (max-abs-err / max-abs-ref): the yardstick for the ill-conditioned spatial-path gradients (Q16)."""
    import copy
    net32 = copy.deepcopy(onet64).float()
    net32.adj_pa = onet64.adj_pa.float()
    net32.cheb = [c.float() for c in onet64.cheb]
    _, g32 = _oracle_run(net32, x.float(), tgt.float())
    return {k: (rel_err(g32[k], g64[k]) if g64[k] is not None else 0.0) for k in g64}


def kink_slack(onet64, x, tgt, g64, tau=1e-5):
    """Per gradient tensor: how far the fp64 oracle's gradient moves (max-abs, relative to the tensor's
    maximum) between "every ReLU unit within tau * max|pre-activation| of its kink counts as active" and
    "... as inactive" (oracle.KINK_SHIFT).  Zero for a conditioned input; at the full benchmark shapes
    (millions of units, always a few within 1e-5 of zero) it is the part of a gradient difference that says
    nothing about the kernels: any correct fp32 implementation may take either side there."""
    from oracle import dstagnn_oracle as O
    try:
        O.KINK_SHIFT = -tau
        _, g_on = _oracle_run(onet64, x.double(), tgt.double())
        O.KINK_SHIFT = tau
        _, g_off = _oracle_run(onet64, x.double(), tgt.double())
    finally:
        O.KINK_SHIFT = 0.0
    out = {}
    for k in g64:
        if g64[k] is None:
            out[k] = 0.0
            continue
        den = g64[k].abs().max().item() or 1.0
        out[k] = (g_on[k] - g_off[k]).abs().max().item() / den
    return out
