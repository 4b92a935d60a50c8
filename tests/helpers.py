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
