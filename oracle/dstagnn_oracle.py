"""CPU oracle for the DSTAGNN block path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch (CPU, any float dtype) restatement of the algorithm in the
reference's ``model/DSTAGNN_my.py``.  Only ``tests/``, ``__graft_entry__.smoke()``
and the CPU-baseline / reference arms of ``bench.py`` may import this file; the
product package ``dstagnn_drought_b200`` never does.

The restatement is *functional*: every stage is a free function over a flat
``{state_dict key: tensor}`` mapping, so the same weights drive the reference
model, this oracle and the CUDA path.  Each function cites the reference lines
it restates (paths relative to the reference checkout).

Pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference itself:
``tests/golden/make_golden.py`` imports the unmodified reference classes in the
build container, runs them (fp64 and fp32, eval mode) and commits small
input/weight/output/gradient fixtures; ``tests/test_oracle_golden.py`` checks
this file against them (and, whenever ``/root/reference`` is present, against
a live run of the reference).

Two execution orders are provided for the Chebyshev stage:
  * ``faithful=True``  -- the reference's own order of operations (a Python loop
    over time steps and Chebyshev orders, the spatial softmax recomputed every
    iteration).  This is what the CPU baseline times.
  * ``faithful=False`` -- the same mathematics with the softmax hoisted out of
    the time loop (one einsum); used by tests because it is ~T times faster.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# ----------------------------------------------------------------------------
# graph preprocessing (host side; reference: lib/utils.py:149-203)
# ----------------------------------------------------------------------------
def scaled_laplacian(W: np.ndarray) -> np.ndarray:
    """L~ = 2 L / lambda_max - I with L = D - W (lib/utils.py:149-177).

    lambda_max is the eigenvalue of largest real part of the (generally
    non-symmetric) Laplacian, obtained in the reference with ARPACK
    ``eigs(L, k=1, which='LR')`` (lib/utils.py:170) -- there in the precision of W (float32 from the
    driver, train_DSTAGNN_my.py:76) and from a random start vector, so the reference's own value moves
    by ~3e-6 between runs.  The oracle computes it in float64 from a fixed start vector (one value,
    inside that spread, the same on every machine -- the parity tests need reproducible polynomials)
    and then follows the reference's arithmetic in W's precision.
    """
    from scipy.sparse.linalg import eigs

    W = np.asarray(W)
    assert W.shape[0] == W.shape[1]
    lap = np.diag(W.sum(axis=1)) - W
    lam = eigs(lap.astype(np.float64), k=1, which="LR",
               v0=np.random.RandomState(0).uniform(-1.0, 1.0, W.shape[0]))[0].real
    if np.issubdtype(lap.dtype, np.floating):
        lam = lap.dtype.type(lam)
    return (2.0 * lap) / lam - np.identity(W.shape[0])


def cheb_polynomials(L_tilde: np.ndarray, K: int) -> List[np.ndarray]:
    """T_0..T_{K-1}; NOTE the recurrence is element-wise (Hadamard), not a matrix
    product: T_k = 2 * L~ (.) T_{k-1} - T_{k-2}  (lib/utils.py:198-203)."""
    n = L_tilde.shape[0]
    polys = [np.identity(n), L_tilde.copy()]
    for k in range(2, K):
        polys.append(2 * L_tilde * polys[k - 1] - polys[k - 2])
    return polys[:max(K, 2)]


# ----------------------------------------------------------------------------
# ReLU with a movable kink for the BACKWARD pass (test instrument)
# ----------------------------------------------------------------------------
# The net's three ReLU sites per block (lines 133, 245/247, 252 of model/DSTAGNN_my.py) make the
# gradients discontinuous in the inputs: a pre-activation within rounding distance of zero is "on" in one
# correct fp32 implementation and "off" in another.  With KINK_SHIFT = s the forward value is still
# relu(x) exactly, but the derivative is taken as [x > s * max|x|]: s = -tau treats every unit within tau of
# the kink as active, s = +tau as inactive.  The difference of the two gradients bounds what flipping those
# units can do; tests add it to the tolerance at shapes with millions of units, where no input avoids the
# kinks (tests/test_gpu_parity.py).  KINK_SHIFT = 0 (default) is the plain ReLU of the reference.
KINK_SHIFT = 0.0


class _ShiftedKinkRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, shift):
        ctx.save_for_backward(x > shift * x.detach().abs().max())
        return torch.relu(x)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * mask, None


def _relu(x: Tensor) -> Tensor:
    return torch.relu(x) if KINK_SHIFT == 0.0 else _ShiftedKinkRelu.apply(x, KINK_SHIFT)


# ----------------------------------------------------------------------------
# stage functions
# ----------------------------------------------------------------------------
def _layer_norm(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def embed_temporal(p: Params, pre: str, x: Tensor) -> Tensor:
    """Positional embedding over T + LayerNorm over N (model/DSTAGNN_my.py:162-182,
    'T' branch).  x: (B,N,F,T) -> (B,F,T,N)."""
    table = p[pre + "EmbedT.pos_embed.weight"]          # (T, N); index = arange(T)
    e = x.permute(0, 2, 3, 1) + table                   # broadcast over (B,F)
    return _layer_norm(e, p[pre + "EmbedT.norm.weight"], p[pre + "EmbedT.norm.bias"])


def temporal_attention(p: Params, pre: str, e: Tensor, res_att, n_heads: int,
                       d_k: int) -> Tuple[Tensor, Tensor]:
    """Multi-head attention over the T axis with d_model = N
    (model/DSTAGNN_my.py:69-100 and 24-42).

    e: (B,F,T,N).  The softmax runs over dim=3 of (B,F,H,Tq,Tk), i.e. over the
    *query* index (model/DSTAGNN_my.py:40).  Returns (LN(fc(ctx)+e), scores) where
    scores are the pre-softmax logits including ``res_att`` (lines 37, 42).
    """
    B, Fe, T, N = e.shape
    q = (e @ p[pre + "TAt.W_Q.weight"].t()).view(B, Fe, T, n_heads, d_k).transpose(2, 3)
    k = (e @ p[pre + "TAt.W_K.weight"].t()).view(B, Fe, T, n_heads, d_k).transpose(2, 3)
    v = (e @ p[pre + "TAt.W_V.weight"].t()).view(B, Fe, T, n_heads, d_k).transpose(2, 3)
    scores = q @ k.transpose(-1, -2) / math.sqrt(d_k) + res_att      # (B,F,H,Tq,Tk)
    attn = torch.softmax(scores, dim=3)
    ctx = (attn @ v).transpose(2, 3).reshape(B, Fe, T, n_heads * d_k)
    out = ctx @ p[pre + "TAt.fc.weight"].t() + e
    out = _layer_norm(out, p[pre + "TAt.layer_norm.weight"], p[pre + "TAt.layer_norm.bias"])
    return out, scores


def pre_conv(p: Params, pre: str, o: Tensor) -> Tensor:
    """Conv2d(T -> d_model, kernel (1,F)) on (B,T,N,F), last column kept
    (model/DSTAGNN_my.py:207,232): a per-node Linear over (t,f).  o: (B,F,T,N) -> (B,N,D)."""
    w = p[pre + "pre_conv.weight"]                      # (D, T, 1, F)
    return torch.einsum("bftn,dtf->bnd", o, w[:, :, 0, :]) + p[pre + "pre_conv.bias"]


def embed_spatial(p: Params, pre: str, z: Tensor) -> Tensor:
    """Positional embedding over N + LayerNorm over d_model
    (model/DSTAGNN_my.py:177-182, 'S' branch).  Dropout (line 234) is identity in eval."""
    s = z + p[pre + "EmbedS.pos_embed.weight"]          # (N, D) table, index = arange(N)
    return _layer_norm(s, p[pre + "EmbedS.norm.weight"], p[pre + "EmbedS.norm.bias"])


def spatial_scores(p: Params, pre: str, se: Tensor, K: int, d_k: int) -> Tensor:
    """K 'heads' of raw scaled scores, no softmax, no V
    (model/DSTAGNN_my.py:44-67, 8-22).  se: (B,N,D) -> (B,K,N,N)."""
    B, N, _ = se.shape
    q = (se @ p[pre + "SAt.W_Q.weight"].t()).view(B, N, K, d_k).transpose(1, 2)
    k = (se @ p[pre + "SAt.W_K.weight"].t()).view(B, N, K, d_k).transpose(1, 2)
    return q @ k.transpose(-1, -2) / math.sqrt(d_k)


def cheb_conv_with_sat(p: Params, pre: str, x: Tensor, sat: Tensor, adj_pa: Tensor,
                       cheb: Sequence[Tensor], faithful: bool = True) -> Tensor:
    """Chebyshev graph convolution with spatial attention
    (model/DSTAGNN_my.py:102-133).  x: (B,N,F,T), sat: (B,K,N,N) -> (B,N,C,T).

    A_k = T_k (.) softmax_{dim=1}(sat_k + adj_pa (.) mask_k)   (lines 126-128; the
    softmax normalises over the ROW index i), out_t = relu(sum_k A_k^T x_t Theta_k).
    """
    B, N, Fin, T = x.shape
    K = len(cheb)
    thetas = [p[f"{pre}cheb_conv_SAt.Theta.{k}"] for k in range(K)]
    masks = [p[f"{pre}cheb_conv_SAt.mask.{k}"] for k in range(K)]
    if faithful:
        steps = []
        for t in range(T):
            sig = x[:, :, :, t]                                        # (B,N,F)
            acc = torch.zeros(B, N, thetas[0].shape[1], dtype=x.dtype)
            for k in range(K):
                a = torch.softmax(sat[:, k] + adj_pa * masks[k], dim=1)
                a = cheb[k] * a
                acc = acc + (a.permute(0, 2, 1) @ sig) @ thetas[k]
            steps.append(acc.unsqueeze(-1))
        return _relu(torch.cat(steps, dim=-1))
    bias = torch.stack([adj_pa * m for m in masks])                    # (K,N,N)
    a = torch.softmax(sat + bias, dim=2) * torch.stack(list(cheb))     # (B,K,N,N)
    z = torch.einsum("bkij,bift->bkjft", a, x)
    return _relu(torch.einsum("bkjft,kfc->bjct", z, torch.stack(thetas)))


def gtu(p: Params, pre: str, name: str, xc: Tensor, C: int) -> Tensor:
    """Gated temporal unit: Conv2d(C -> 2C, (1,k)) then tanh(first C) * sigmoid(last C)
    (model/DSTAGNN_my.py:184-197).  xc: (B,C,N,T) -> (B,C,N,T-k+1)."""
    u = F.conv2d(xc, p[f"{pre}{name}.con2out.weight"], p[f"{pre}{name}.con2out.bias"])
    return torch.tanh(u[:, :C]) * torch.sigmoid(u[:, -C:])


def block_forward(p: Params, idx: int, x: Tensor, res_att, adj_pa: Tensor,
                  cheb: Sequence[Tensor], n_heads: int, d_k: int,
                  faithful: bool = True) -> Tuple[Tensor, Tensor]:
    """One ST block (model/DSTAGNN_my.py:225-253).  x: (B,N,F,T) -> ((B,N,C,T), scores)."""
    pre = f"BlockList.{idx}."
    B, N, Fin, T = x.shape
    K = len(cheb)
    C = p[pre + "ln.weight"].shape[0]
    e = embed_temporal(p, pre, x) if Fin == 1 else x.permute(0, 2, 3, 1)    # lines 227-230
    o, scores = temporal_attention(p, pre, e, res_att, n_heads, d_k)       # line 231
    se = embed_spatial(p, pre, pre_conv(p, pre, o))                        # lines 232-234
    sat = spatial_scores(p, pre, se, K, d_k)                               # line 235
    g = cheb_conv_with_sat(p, pre, x, sat, adj_pa, cheb, faithful)         # line 236
    xc = g.permute(0, 2, 1, 3)                                             # (B,C,N,T)
    tc = torch.cat([gtu(p, pre, n, xc, C) for n in ("gtu3", "gtu5", "gtu7")], dim=-1)
    tc = tc @ p[pre + "fcmy.0.weight"].t() + p[pre + "fcmy.0.bias"]        # line 243 (dropout = id)
    if Fin == 1:                                                           # lines 244-251
        tco = _relu(tc)
        r = F.conv2d(x.permute(0, 2, 1, 3), p[pre + "residual_conv.weight"],
                     p[pre + "residual_conv.bias"])
    else:
        tco = _relu(xc + tc)
        r = x.permute(0, 2, 1, 3)
    y = _relu(r + tco).permute(0, 3, 2, 1)                            # (B,T,N,C)
    y = _layer_norm(y, p[pre + "ln.weight"], p[pre + "ln.bias"]).permute(0, 2, 3, 1)
    return y, scores


def net_forward(p: Params, x: Tensor, adj_pa: Tensor, cheb: Sequence[Tensor],
                nb_block: int, n_heads: int, d_k: int, faithful: bool = True,
                return_blocks: bool = False):
    """Block stack + final_conv + final_fc (model/DSTAGNN_my.py:271-280).
    x: (B,N,F,T) -> (B,N,P)."""
    outs, res_att, att = [], 0, []
    for i in range(nb_block):
        x, res_att = block_forward(p, i, x, res_att, adj_pa, cheb, n_heads, d_k, faithful)
        outs.append(x)
        att.append(res_att)
    cat = torch.cat(outs, dim=-1)                                          # (B,N,C,T*nb)
    w = p["final_conv.weight"]                                             # (128, T*nb, 1, C)
    h = torch.einsum("bncs,osc->bno", cat, w[:, :, 0, :]) + p["final_conv.bias"]
    y = h @ p["final_fc.weight"].t() + p["final_fc.bias"]
    if return_blocks:
        return y, outs, att
    return y


# ----------------------------------------------------------------------------
# parameter construction (model/DSTAGNN_my.py:199-223, 255-269, 282-297)
# ----------------------------------------------------------------------------
def param_shapes(nb_block: int, in_channels: int, K: int, C: int, N: int, T: int,
                 d_model: int, d_k: int, n_heads: int, P: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict keys and shapes in registration order."""
    out: List[Tuple[str, Tuple[int, ...]]] = []
    for b in range(nb_block):
        pre = f"BlockList.{b}."
        fin = in_channels if b == 0 else C
        nd = in_channels if b == 0 else in_channels * C
        out += [
            (pre + "pre_conv.weight", (d_model, T, 1, nd)), (pre + "pre_conv.bias", (d_model,)),
            (pre + "EmbedT.pos_embed.weight", (T, N)),
            (pre + "EmbedT.norm.weight", (N,)), (pre + "EmbedT.norm.bias", (N,)),
            (pre + "EmbedS.pos_embed.weight", (N, d_model)),
            (pre + "EmbedS.norm.weight", (d_model,)), (pre + "EmbedS.norm.bias", (d_model,)),
            (pre + "TAt.W_Q.weight", (n_heads * d_k, N)), (pre + "TAt.W_K.weight", (n_heads * d_k, N)),
            (pre + "TAt.W_V.weight", (n_heads * d_k, N)), (pre + "TAt.fc.weight", (N, n_heads * d_k)),
            (pre + "TAt.layer_norm.weight", (N,)), (pre + "TAt.layer_norm.bias", (N,)),
            (pre + "SAt.W_Q.weight", (K * d_k, d_model)), (pre + "SAt.W_K.weight", (K * d_k, d_model)),
        ]
        out += [(f"{pre}cheb_conv_SAt.Theta.{k}", (fin, C)) for k in range(K)]
        out += [(f"{pre}cheb_conv_SAt.mask.{k}", (N, N)) for k in range(K)]
        for ks in (3, 5, 7):
            out += [(f"{pre}gtu{ks}.con2out.weight", (2 * C, C, 1, ks)),
                    (f"{pre}gtu{ks}.con2out.bias", (2 * C,))]
        out += [
            (pre + "residual_conv.weight", (C, fin, 1, 1)), (pre + "residual_conv.bias", (C,)),
            (pre + "fcmy.0.weight", (T, 3 * T - 12)), (pre + "fcmy.0.bias", (T,)),
            (pre + "ln.weight", (C,)), (pre + "ln.bias", (C,)),
        ]
    out += [("final_conv.weight", (128, T * nb_block, 1, C)), ("final_conv.bias", (128,)),
            ("final_fc.weight", (P, 128)), ("final_fc.bias", (P,))]
    return out


def init_params(shapes, seed: int = 1, dtype=torch.float32) -> Params:
    """Reference init rule (model/DSTAGNN_my.py:292-296): xavier_uniform_ for every
    tensor with dim>1, uniform_(0,1) for every 1-D tensor (biases AND LayerNorm
    gamma/beta).  The RNG stream differs from make_model's (which also burns the
    stock nn.Module constructors' draws), so values are equal in distribution only."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}
    for name, shp in shapes:
        t = torch.empty(shp, dtype=torch.float32)
        if t.dim() > 1:
            fan_in, fan_out = torch.nn.init._calculate_fan_in_and_fan_out(t)
            bound = math.sqrt(6.0 / (fan_in + fan_out))
            t.uniform_(-bound, bound, generator=g)
        else:
            t.uniform_(0.0, 1.0, generator=g)
        p[name] = t.to(dtype)
    return p


def synthetic_graph(N: int, seed: int = 0, density: float = 0.01):
    """STAG/STRG-like {0,1} matrices: per row the top
    max(1, floor(density*N)) random entries are ones (data/STAG_gen.py:108-116 +
    lib/dataloader.py:5-15); adj_pa gets one more per row; adj_mx = adj_TMD (graph=AG)."""
    rng = np.random.RandomState(seed)
    m = max(1, int(density * N))
    tmd = np.zeros((N, N), np.float32)
    pa = np.zeros((N, N), np.float32)
    for i in range(N):
        cols = rng.permutation(N)[: m + 1]
        tmd[i, cols[:m]] = 1.0
        pa[i, cols] = 1.0
    return tmd, pa


class OracleNet(torch.nn.Module):
    """nn.Module shell so autograd / optimisers / ``state_dict`` work on the oracle.
    Parameter names equal the reference's state_dict keys ('.' replaced for storage)."""

    def __init__(self, params: Params, adj_pa: Tensor, cheb: Sequence[Tensor], nb_block: int,
                 n_heads: int, d_k: int, faithful: bool = True):
        super().__init__()
        self._names = list(params.keys())
        self.store = torch.nn.ParameterDict(
            {k.replace(".", "/"): torch.nn.Parameter(v.clone()) for k, v in params.items()})
        self.adj_pa = adj_pa
        self.cheb = list(cheb)
        self.nb_block, self.n_heads, self.d_k, self.faithful = nb_block, n_heads, d_k, faithful

    def named(self) -> Params:
        return {k: self.store[k.replace(".", "/")] for k in self._names}

    def forward(self, x: Tensor) -> Tensor:
        return net_forward(self.named(), x, self.adj_pa, self.cheb, self.nb_block,
                           self.n_heads, self.d_k, self.faithful)


def build_oracle(N: int, T: int, K: int, n_heads: int, d_model: int, nb_block: int,
                 P: int = 12, C: int = 32, d_k: int = 32, in_channels: int = 1,
                 seed: int = 1, graph_seed: int = 0, dtype=torch.float32,
                 faithful: bool = True) -> OracleNet:
    """Synthetic-graph, reference-init oracle net for one of the benchmark shapes."""
    tmd, pa = synthetic_graph(N, graph_seed)
    cheb = [torch.from_numpy(c).to(dtype) for c in cheb_polynomials(scaled_laplacian(tmd), K)]
    shapes = param_shapes(nb_block, in_channels, K, C, N, T, d_model, d_k, n_heads, P)
    params = init_params(shapes, seed, dtype)
    return OracleNet(params, torch.from_numpy(pa).to(dtype), cheb, nb_block, n_heads, d_k, faithful)
