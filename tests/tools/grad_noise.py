"""Run-to-run spread of the parameter gradients at the GAMBIA-layout test shape (N=37, T=144) against the fp64
oracle, and the tooling that traced its one discrete outlier to a ReLU kink rather than to a kernel fault.

    python tests/tools/grad_noise.py 40 quiet            # per repetition: worst gradient error; outliers listed
    python tests/tools/grad_noise.py 40 quiet check      # + every contraction issued twice, results compared
    python tests/tools/grad_noise.py 40 quiet trace      # + element-wise comparison of every autograd Function's
                                                   #   outputs against a good repetition, ReLU sign flips
    python tests/tools/grad_noise.py 40 quiet fp32       # CUDA-core GEMMs (no atomically summed slices)

Finding (B200, round 1): in ~10 % of the repetitions one of the 340 992 outputs of block 1's Theta contraction
(relu fused) is 0 instead of 3.4e-9 -- the forward sums its K-outer slices atomically, the pre-activation is
within fp32 rounding of zero -- and the blocked gradient moves mask.1 by 2.4e-3 and TAt.W_Q of block 0 by
8.9e-4.  Every contraction reproduces itself exactly (check), inputs of the first diverging op are identical
(trace).  Needs the oracle: test infrastructure only."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import dstagnn_oracle as O  # noqa: E402
from dstagnn_drought_b200 import make_model  # noqa: E402


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def install_rerun_check(tol=2e-5):
    """Every STORE contraction is issued twice into the same buffer; the two results must agree (bitwise when
    the GEMM is not split-K, to rounding noise otherwise).  Localises timing-dependent faults to one GEMM."""
    import dstagnn_drought_b200.contraction as ct
    import dstagnn_drought_b200.ops as ops
    import dstagnn_drought_b200.modules as mods
    orig = ct.contract_raw
    state = {"busy": False, "n": 0, "bad": 0}

    def wrapped(spec, A, B, **kw):
        acc = kw.get("accumulate", ct.ACC_STORE)
        if state["busy"]:
            return orig(spec, A, B, **kw)
        state["busy"] = True
        try:
            before = kw["out"].clone() if acc != ct.ACC_STORE else None
            res = orig(spec, A, B, **kw)
            first = res.clone()
            kw2 = dict(kw)
            kw2["out"] = res
            if before is not None:
                res.copy_(before)
            orig(spec, A, B, **kw2)
            state["n"] += 1
            d = float((res - first).abs().max())
            scale = float(first.abs().max())
            if d > tol * max(scale, 1e-30):
                state["bad"] += 1
                print(f"   RERUN MISMATCH {d / max(scale, 1e-30):.2e}  spec={spec} tag={kw.get('tag')} acc={acc} A={tuple(A.shape)}/{A.stride()} "
                      f"B={tuple(B.shape)}/{B.stride()} out={tuple(res.shape)}/{res.stride()}")
            return res
        finally:
            state["busy"] = False

    for m in (ct, ops, mods):
        if hasattr(m, "contract_raw"):
            m.contract_raw = wrapped
    return state


def install_trace():
    """Checksums (sum |x| in float64) of every autograd Function's forward outputs and backward gradients, in
    execution order; two repetitions' traces are compared to find the first op whose result moved."""
    import dstagnn_drought_b200.contraction as ct
    import dstagnn_drought_b200.ops as ops
    trace = []

    def csum(t):                                  # a full copy: the comparison is element-wise
        return t.detach().clone() if isinstance(t, torch.Tensor) else None

    def wrap(cls):
        fwd, bwd = cls.forward, cls.backward

        def forward(ctx, *a):
            out = fwd(ctx, *a)
            outs = out if isinstance(out, tuple) else (out,)
            extra = [csum(t) for t in a if isinstance(t, torch.Tensor)]
            trace.append((cls.__name__ + ".fwd", [csum(t) for t in outs], extra))
            return out

        def backward(ctx, *g):
            gin = [csum(t) for t in g]
            saved = [csum(t) for t in ctx.saved_tensors]
            out = bwd(ctx, *g)
            outs = out if isinstance(out, tuple) else (out,)
            trace.append((cls.__name__ + ".bwd", [csum(t) for t in outs], gin + ["saved"] + saved))
            return out

        cls.forward, cls.backward = staticmethod(forward), staticmethod(backward)

    orig_cr, orig_cs = ops.contract_raw, ops.colsum

    keep = {}
    trace_keep = keep

    def cr(spec, A, B, **kw):
        out = orig_cr(spec, A, B, **kw)
        if kw.get("tag") == "cheb_bwd_dmap" and "dmap" not in keep:
            keep["dmap"] = out.detach().clone()
            keep["A"], keep["B"] = A, B
        trace.append((f"  contract_raw {spec} {kw.get('tag')}", [csum(out)], [csum(A), csum(B)]))
        return out

    def cs(t2d, rows, cols, ld=None):
        out = orig_cs(t2d, rows, cols, ld)
        trace.append(("  colsum", [csum(out)], [csum(t2d)]))
        return out

    ops.contract_raw, ops.colsum = cr, cs
    for cls in (ct._Contract, ops.Transpose, ops.LayerNormAdd, ops.TemporalAttention, ops.SpatialChebConv,
                ops.GtuStack, ops.BlockTail, ops.FusedTail):
        wrap(cls)
    trace.append(keep)
    return trace


def diff_traces(a, b, tol=2e-6):
    n = 0
    for i, (x, y) in enumerate(zip(a, b)):          # ReLU kinks: activations whose sign differs between the repetitions
        if x[0].endswith(".fwd") and x[1] and x[1][0] is not None and x[1][0].shape == y[1][0].shape:
            u, v = x[1][0], y[1][0]
            flip = (u > 0) != (v > 0)
            if bool(flip.any()) and bool((u >= 0).all()):
                print(f"   trace[{i}] {x[0]}: {int(flip.sum())} of {u.numel()} relu outputs flip; values there: good "
                      f"{u[flip].flatten()[:4].tolist()} outlier {v[flip].flatten()[:4].tolist()} (max |out| {float(u.max()):.3e})")
    for i, (x, y) in enumerate(zip(a, b)):
        def rel(u, v):
            if u is None or v is None or isinstance(u, str) or u.shape != v.shape:
                return 0.0
            return float((u - v).abs().max()) / max(float(u.abs().max()), 1e-30)
        dout = max([rel(u, v) for u, v in zip(x[1], y[1])] or [0.0])
        din = max([rel(u, v) for u, v in zip(x[2], y[2])] or [0.0])
        if dout > tol or din > tol:
            print(f"   trace[{i}] {x[0]}: outputs moved {dout:.2e}, inputs/saved moved {din:.2e}")
            n += 1
            if n >= 10:
                break


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    trace = install_trace() if "trace" in sys.argv else None
    keep = trace.pop() if trace is not None else None
    good_trace = None
    good_dmap = None
    chk = install_rerun_check() if "check" in sys.argv else None
    quiet = len(sys.argv) > 2 and sys.argv[2] == "quiet"          # only outliers and a summary line
    if len(sys.argv) > 3 and sys.argv[3] in ("fp32", "tf32x3", "tf32"):
        import dstagnn_drought_b200 as D
        D.set_precision(sys.argv[3])
    N, T, K, H, Dm, nb, B = 37, 144, 2, 2, 64, 2, 2
    dev = torch.device("cuda:0")
    onet = O.build_oracle(N, T, K, H, Dm, nb, dtype=torch.float64, faithful=False)
    tmd, pa = O.synthetic_graph(N, 0)
    net = make_model(dev, 1, nb, 1, K, 32, 32, 1, torch.from_numpy(tmd), torch.from_numpy(pa),
                     torch.from_numpy(tmd), 12, T, N, Dm, 32, 32, H)
    net.load_state_dict({k: v.detach().float() for k, v in onet.named().items()})
    net.eval()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, N, 1, T, generator=g)
    tgt = torch.randn(B, N, 12, generator=g)
    torch.nn.SmoothL1Loss()(onet(x.double()), tgt.double()).backward()
    og = {k: v.grad for k, v in onet.named().items()}
    base, ybase = None, None
    for r in range(reps):
        net.zero_grad(set_to_none=True)
        if trace is not None:
            trace.clear()
            keep.clear()
        y = net(x.to(dev))
        torch.nn.SmoothL1Loss()(y, tgt.to(dev)).backward()
        cur = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
        if base is None:
            base, ybase = cur, y.detach().clone()
        errs = {k: rel_err(p.grad, og[k]) for k, p in net.named_parameters() if og[k] is not None}
        worst = max(errs, key=errs.get)
        outl = errs[worst] > 1e-4
        if trace is not None:
            if not outl and good_trace is None:
                good_trace = list(trace)
                good_dmap = keep["dmap"]
            if outl and good_trace is not None:
                print(f"rep {r} is an outlier; first ops that moved against a good repetition:")
                diff_traces(good_trace, list(trace))
                d = (keep["dmap"] - good_dmap).abs()
                ref = keep["A"].double(), keep["B"].double()
                exact = torch.einsum("kbje,bie->bkji", *ref)
                print("   dmap out shape", tuple(d.shape), "max|d|", float(d.max()), "max|out|", float(good_dmap.abs().max()),
                      "n(|d|>1e-6 max)", int((d > 1e-6 * good_dmap.abs().max()).sum()), "of", d.numel())
                print("   err vs fp64: good", float((good_dmap.double() - exact).abs().max()), "outlier",
                      float((keep["dmap"].double() - exact).abs().max()))
                idx = (d > 1e-6 * good_dmap.abs().max()).nonzero()
                print("   b:", sorted(set(idx[:, 0].tolist())), "k:", sorted(set(idx[:, 1].tolist())),
                      "j range", int(idx[:, 2].min()), int(idx[:, 2].max()), "i range", int(idx[:, 3].min()), int(idx[:, 3].max()))
        nout = (nout if r else 0) + int(outl)
        if quiet and not outl:
            continue
        if outl and base is not cur:                                  # which gradients moved against repetition 0?
            print("   forward output vs rep 0:", rel_err(y, ybase))
            for k in cur:
                d = rel_err(cur[k], base[k])
                if d > 1e-6:
                    print(f"   moved {d:.2e}  {k}")
        print(f"rep {r}: worst {errs[worst]:.3e} ({worst})  TAt0.W_Q {errs['BlockList.0.TAt.W_Q.weight']:.3e} "
              f"|g|max {float(og['BlockList.0.TAt.W_Q.weight'].abs().max()):.3e}")
    print(f"outliers (worst > 1e-4): {nout} of {reps}")
    if chk:
        print(f"rerun check: {chk['bad']} mismatching of {chk['n']} contractions")


if __name__ == "__main__":
    main()
