"""Generate golden fixtures by running the UNMODIFIED reference model on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For each small configuration it builds ``make_model('cpu', ...)`` from the
reference's ``model/DSTAGNN_my.py`` (matplotlib is stubbed because
``lib/utils.py:14-16`` imports it and it is not installed), runs forward and
backward in eval mode (dropout off) in fp32 and in fp64, and
stores inputs, graph matrices, the reference's own Chebyshev polynomials,
weights, per-block outputs, attention logits, predictions and every parameter
gradient in ``tests/golden/<name>.npz``.  The fixtures travel to the GPU box;
the reference does not.
"""
import copy
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: N, T, K, H, D, nb_block, B, P
    "g_n20": dict(N=20, T=12, K=3, H=2, D=16, nb=2, B=2, P=12),
    "g_n33": dict(N=33, T=12, K=3, H=3, D=32, nb=3, B=3, P=12),
    "g_t18": dict(N=24, T=18, K=2, H=2, D=16, nb=2, B=2, P=6),
    # the one real graph the reference ships (data/PEMS04/stag_001_PEMS04.csv, strg_001_PEMS04.csv; N = 307,
    # 3 ones per STAG row, 4 non-zeros per STRG row), loaded with the reference's own lib/dataloader.py;
    # d_model and nb_block are reduced from the conf's 512 / 4 to keep the fixture small
    "g_pems04_real": dict(N=307, T=12, K=3, H=3, D=64, nb=2, B=2, P=12, real="PEMS04"),
}


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from model.DSTAGNN_my import make_model  # noqa
    return make_model


def graph(N, seed):
    rng = np.random.RandomState(seed)
    m = max(1, N // 8)
    tmd = np.zeros((N, N), np.float32)
    pa = np.zeros((N, N), np.float32)
    for i in range(N):
        cols = rng.permutation(N)[: m + 1]
        tmd[i, cols[:m]] = 1.0
        pa[i, cols] = 1.0
    return tmd, pa


def real_graph(name):
    """STAG / STRG of a dataset shipped with the reference, through the reference's own loaders
    (lib/dataloader.py:5-15, called at train_DSTAGNN_my.py:71-73; graph = AG => adj_mx = adj_TMD)."""
    from lib.dataloader import load_PA, load_weighted_adjacency_matrix
    base = os.path.join(REF, "data", name)
    tmd = load_weighted_adjacency_matrix(os.path.join(base, f"stag_001_{name}.csv"), None)
    pa = load_PA(os.path.join(base, f"strg_001_{name}.csv"))
    return np.float32(tmd), np.float32(pa)


def run(net, x, target):
    net.zero_grad(set_to_none=True)
    blocks, atts = [], []
    def grab(_m, _i, o):
        blocks.append(o[0].detach().clone())
        atts.append(o[1].detach().clone())

    hooks = [b.register_forward_hook(grab) for b in net.BlockList]
    y = net(x)
    loss = torch.nn.SmoothL1Loss()(y, target)
    loss.backward()
    for h in hooks:
        h.remove()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else None)
             for k, p in net.named_parameters()}
    return y.detach(), loss.detach(), blocks, atts, grads


def main():
    make_model = import_reference()
    only = set(sys.argv[1:])
    for name, c in CASES.items():
        if only and name not in only:
            continue
        N, T, K, H, D, nb, B, P = (c[k] for k in ("N", "T", "K", "H", "D", "nb", "B", "P"))
        tmd, pa = real_graph(c["real"]) if "real" in c else graph(N, seed=len(name))
        assert tmd.shape == (N, N) and pa.shape == (N, N)
        torch.manual_seed(1)
        net = make_model("cpu", 1, nb, 1, K, 32, 32, 1, torch.FloatTensor(tmd), torch.FloatTensor(pa),
                         torch.FloatTensor(tmd), P, T, N, D, 32, 32, H)
        net.eval()
        cheb = [t.detach().clone() for t in net.BlockList[0].cheb_conv_SAt.cheb_polynomials]
        net64 = copy.deepcopy(net).double()
        for blk in net64.BlockList:      # plain attributes are not moved by .double() (Q15)
            blk.adj_pa = blk.adj_pa.double()
            blk.cheb_conv_SAt.cheb_polynomials = [t.double() for t in cheb]

        # Conditioning check: the model has ReLUs, so an input whose pre-activation sits within fp32
        # noise of a kink has a gradient that ANY fp32 implementation may or may not reproduce.  Pick
        # the first data seed whose fp64 gradients are stable under a 1e-6 relative input perturbation.
        for data_seed in range(7, 64):
            g = torch.Generator().manual_seed(data_seed)
            x = torch.randn(B, N, 1, T, generator=g)
            target = torch.randn(B, N, P, generator=g)
            y64, l64, blk64, att64, g64 = run(net64, x.double(), target.double())
            cond = 0.0
            for trial in range(8):              # several draws: each flips a near-kink unit with p ~ 1/2
                eps = (1e-7, 3e-7, 1e-6, 3e-6)[trial % 4]
                netp = copy.deepcopy(net64)            # perturb every weight AND the input
                with torch.no_grad():
                    for prm in netp.parameters():
                        prm.mul_(1 + eps * torch.randn(prm.shape, generator=g, dtype=torch.float64))
                xp = x.double() * (1 + eps * torch.randn(x.shape, generator=g, dtype=torch.float64))
                _, _, _, _, gp = run(netp, xp, target.double())
                cond = max(cond, max(float((gp[k] - g64[k]).abs().max() / g64[k].abs().max()) / (eps / 1e-6)
                                     for k in g64 if g64[k] is not None))
            print(f"  {name}: data seed {data_seed}: gradient sensitivity per 1e-6 weight+input noise (max of 8 draws) = {cond:.2e}")
            if cond < 2e-5:
                break
        else:
            raise SystemExit("no well-conditioned data seed found")

        y32, l32, blk32, att32, g32 = run(net, x, target)

        out = {"cfg": np.array([N, T, K, H, D, nb, B, P], np.int64), "data_seed": np.array(data_seed),
               "x": x.numpy(), "target": target.numpy(), "adj_tmd": tmd, "adj_pa": pa,
               "y32": y32.numpy(), "y64": y64.numpy(),
               "loss32": l32.numpy(), "loss64": l64.numpy()}
        for k, t in enumerate(cheb):
            out[f"cheb.{k}"] = t.numpy()
        for k, v in net.state_dict().items():
            out["w." + k] = v.numpy()
        for i in range(nb):
            # fp64 results are stored rounded to fp32 (6e-8 relative: far below any gate)
            out[f"blk64.{i}"] = blk64[i].float().numpy()
            out[f"att64.{i}"] = att64[i].float().numpy()
        dead, err32 = [], []
        for k in g64:
            if g64[k] is None:
                dead.append(k)
            else:
                out["g64." + k] = g64[k].float().numpy()
                # the reference's own fp32 error on this gradient (max-abs-err / max-abs-ref), Q16
                err32.append(float((g32[k].double() - g64[k]).abs().max() / g64[k].abs().max()))
        out["dead"] = np.array(dead)
        out["ref_fp32_grad_err"] = np.array(err32)
        out["ref_fp32_grad_err_keys"] = np.array([k for k in g64 if g64[k] is not None])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB; dead params:", len(dead),
              "loss", float(l64))


if __name__ == "__main__":
    main()
