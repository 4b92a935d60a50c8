"""Einsum-style front end of ``dstag_gemm`` (include/dstagnn_b200.h).

``contract("bftn,dtf->bnd", O, W)`` classifies every index letter by where it
occurs -- in both inputs and the output (batch), first input and output (m),
second input and output (n), both inputs only (k: summed) -- and hands the
kernel the operands' *native strides*, so permuted / sliced / unfolded views are
consumed in place: no ``.contiguous()``, no im2col buffer.  Each role may span at
most two tensor dimensions (the C ABI's two-level axis).

``contract`` is differentiable (both operands and the additive epilogue term);
its backward is two more ``contract_raw`` calls with the roles swapped, which is
how every weight- and input-gradient of the dense products in
model/DSTAGNN_my.py:62-63, 92-99, 128-131, 193, 207, 221, 232, 243, 266-267 is
produced.  ``contract_raw`` is the non-differentiable primitive used inside the
hand-written autograd Functions of ``ops.py``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ACC_ADD, ACC_ATOMIC, ACC_STORE, Axis, Gemm, Operand

_PRECISION = {"fp32": _lib.PREC_FP32, "tf32x3": _lib.PREC_TF32X3, "tf32": _lib.PREC_TF32}
_precision = _lib.PREC_TF32X3
NUM_SMS = 148


def set_precision(name: str):
    """'fp32' (CUDA-core FMA), 'tf32x3' (tensor cores, fp32-grade; default) or 'tf32'."""
    global _precision
    _precision = _PRECISION[name]


def get_precision() -> str:
    return {v: k for k, v in _PRECISION.items()}[_precision]


class Plan:
    """Index-role classification of one contraction spec."""

    def __init__(self, spec: str):
        ins, out = spec.replace(" ", "").split("->")
        a, b = ins.split(",")
        for s in (a, b, out):
            if len(set(s)) != len(s):
                raise ValueError(f"repeated index in '{s}' of '{spec}'")
        self.a, self.b, self.c = a, b, out
        sa, sb, sc = set(a), set(b), set(out)
        if not sc <= (sa | sb):
            raise ValueError(f"output index missing from inputs in '{spec}'")
        if (sa - sb - sc) or (sb - sa - sc):
            raise ValueError(f"index summed over a single operand in '{spec}' is not supported")
        self.batch = [l for l in out if l in sa and l in sb]
        self.m = [l for l in out if l in sa and l not in sb]
        self.n = [l for l in out if l in sb and l not in sa]
        self.k = [l for l in a if l in sb and l not in sc]
        for role, letters in (("batch", self.batch), ("m", self.m), ("n", self.n), ("k", self.k)):
            if len(letters) > 2:
                raise ValueError(f"role {role} spans {len(letters)} dims {letters} in '{spec}' (max 2); "
                                 "merge dims with .view() first")


_plans = {}


def plan(spec: str) -> Plan:
    p = _plans.get(spec)
    if p is None:
        p = _plans[spec] = Plan(spec)
    return p


def _sizes(pl: Plan, A, B, extra):
    size = {}
    for sub, t in ((pl.a, A), (pl.b, B)) + tuple(extra):
        if t.dim() != len(sub):
            raise ValueError(f"operand with subscripts '{sub}' has {t.dim()} dims {tuple(t.shape)}")
        for l, s in zip(sub, t.shape):
            if size.setdefault(l, s) != s:
                raise ValueError(f"size mismatch for index '{l}': {size[l]} vs {s}")
    return size


def _axis(letters: Sequence[str], sub: str, t: torch.Tensor, size) -> Axis:
    def stride(l):
        i = sub.find(l)
        return t.stride(i) if i >= 0 else 0

    if len(letters) == 0:
        return Axis(0, 0, 1, 0)
    if len(letters) == 1:
        return Axis(stride(letters[0]), 0, 1, 0)
    hi, lo = letters
    s_hi, s_lo, d_lo = stride(hi), stride(lo), size[lo]
    if s_hi == s_lo * d_lo:                      # the two dims flatten into one
        return Axis(s_lo, 0, 1, 0)
    return Axis(s_hi, s_lo, d_lo, 0)


def _operand(t: Optional[torch.Tensor], sub: str, batch, row, col, size) -> Operand:
    if t is None:
        z = Axis(0, 0, 1, 0)
        return Operand(None, z, z, z)
    return Operand(t.data_ptr(), _axis(batch, sub, t, size), _axis(row, sub, t, size),
                   _axis(col, sub, t, size))


def _prod(letters, size):
    r = 1
    for l in letters:
        r *= size[l]
    return r


def _check(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"dstagnn_drought_b200: {name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"dstagnn_drought_b200: {name} must be float32, got {t.dtype}")


def _k_order(pl: Plan, A, B):
    """Enumeration order of a two-letter reduction index.  Any order is correct; prefer the one whose
    inner letter is memory-contiguous in more operands (K-contiguous tiles load by vector / TMA)."""
    if len(pl.k) != 2:
        return pl.k

    def score(order):
        lo = order[1]
        return sum(1 for sub, t in ((pl.a, A), (pl.b, B)) if t.stride(sub.index(lo)) == 1)

    fwd, rev = pl.k, [pl.k[1], pl.k[0]]
    return rev if score(rev) > score(fwd) else fwd


_FP32_SPECS = [t for t in os.environ.get("DSTAG_FP32_SPECS", "").split(";") if t]
_desc_cache = {}
_swap_cache = {}
_dims_cache = {}
_reproducible = os.environ.get("DSTAG_REPRODUCIBLE", "1") != "0"      # A/B switch for measurements


def set_reproducible(on: bool):
    """Module default of ``contract_raw(reproducible=None)``: True (default) = split reductions go through
    a workspace and a fixed-order sum; False = atomic accumulation (results vary in the last bits from run
    to run).  The differentiable ``contract`` always runs its FORWARD product reproducibly."""
    global _reproducible
    _reproducible = bool(on)


def rows_view(t: torch.Tensor):
    """(rows, cols, ld) when ``t`` can be read as a rows x cols matrix with a constant row stride ld >= cols
    (contiguous tensors, row-padded tensors, one half of a packed [Q | K] buffer, ...), else None."""
    if t.dim() == 0:
        return (1, 1, 1)
    cols, d = 1, t.dim() - 1
    while d >= 0 and (t.shape[d] == 1 or t.stride(d) == cols):      # trailing contiguous block
        cols *= t.shape[d]
        d -= 1
    if d < 0:
        return (1, cols, cols)
    ld, rows = t.stride(d), t.shape[d]
    if ld < cols:
        return None
    expect = ld * t.shape[d]
    for e in range(d - 1, -1, -1):                                   # the leading dims must collapse into one
        if t.shape[e] != 1 and t.stride(e) != expect:
            return None
        rows *= t.shape[e]
        expect *= t.shape[e]
    return (rows, cols, ld)


def _workspace(n: int, out: torch.Tensor) -> torch.Tensor:
    """(n, *out.shape) contiguous scratch for the partial results of a two-stage reduction; the slice stride
    is kept a multiple of 4 floats so that the engine's 16-byte store paths stay available."""
    numel = out.numel()
    stride = (numel + 3) // 4 * 4
    buf = torch.empty(n * stride, device=out.device, dtype=torch.float32)
    return buf.as_strided((n,) + tuple(out.shape), (stride,) + tuple(torch.empty(out.shape, device="meta").stride()))


def _layout_key(t):
    return None if t is None else (t.shape, t.stride())


def _m_order(pl: Plan, A):
    """Enumeration order of a two-letter output-row index.  Rows of C may be visited in any order (the
    epilogue addresses every row by its strides); prefer the order in which the first operand's rows are
    memory-contiguous, so its tiles load by 16-byte copies (e.g. the (t, f) rows of a conv weight)."""
    if len(pl.m) != 2:
        return pl.m
    hi, lo = pl.m
    if A.stride(pl.a.index(lo)) != 1 and A.stride(pl.a.index(hi)) == 1:
        return [lo, hi]
    return pl.m


def build_descriptor(spec: str, A, B, out, a_mul=None, c_add=None, c_mul=None, alpha=1.0,
                     relu=False, accumulate=ACC_STORE, split_k=1, precision=None) -> Gemm:
    """Pure host logic (no CUDA call): used directly by the CPU tests.  A training step repeats the
    same ~100 contractions with the same layouts: the filled struct is cached per (spec, layouts,
    flags) and only the six data pointers are refreshed on a hit."""
    prec = _precision if precision is None else precision
    am_t = a_mul[1] if a_mul else None
    ca_t = c_add[1] if c_add else None
    cm_t = c_mul[1] if c_mul else None
    key = (spec, _layout_key(A), _layout_key(B), _layout_key(out),
           a_mul[0] if a_mul else None, _layout_key(am_t), c_add[0] if c_add else None, _layout_key(ca_t),
           c_mul[0] if c_mul else None, _layout_key(cm_t), float(alpha), bool(relu), int(accumulate), int(split_k), prec)
    hit = _desc_cache.get(key)
    if hit is None:
        g = _build_descriptor(spec, A, B, out, a_mul, c_add, c_mul, alpha, relu, accumulate, split_k, prec)
        if len(_desc_cache) < 4096:
            _desc_cache[key] = bytes(g)
        return g
    g = Gemm.from_buffer_copy(hit)
    g.A.ptr, g.B.ptr, g.C.ptr = A.data_ptr(), B.data_ptr(), out.data_ptr()
    if am_t is not None:
        g.Amul.ptr = am_t.data_ptr()
    if ca_t is not None:
        g.Cadd.ptr = ca_t.data_ptr()
    if cm_t is not None:
        g.Cmul.ptr = cm_t.data_ptr()
    return g


def _build_descriptor(spec: str, A, B, out, a_mul, c_add, c_mul, alpha, relu, accumulate, split_k, precision) -> Gemm:
    pl = plan(spec)
    extra = [(pl.c, out)]
    for opt in (a_mul, c_add, c_mul):
        if opt is not None:
            extra.append(opt)
    size = _sizes(pl, A, B, extra)
    if a_mul is not None and not set(a_mul[0]) <= set(pl.a):
        raise ValueError("a_mul subscripts must be a subset of the first operand's")
    for opt in (c_add, c_mul):
        if opt is not None and not set(opt[0]) <= set(pl.c):
            raise ValueError("c_add / c_mul subscripts must be a subset of the output's")
    g = Gemm()
    g.M, g.N, g.K, g.batch = (_prod(pl.m, size), _prod(pl.n, size), _prod(pl.k, size),
                              _prod(pl.batch, size))
    kl = _k_order(pl, A, B)
    ml = _m_order(pl, A)
    g.A = _operand(A, pl.a, pl.batch, ml, kl, size)
    g.B = _operand(B, pl.b, pl.batch, kl, pl.n, size)
    g.C = _operand(out, pl.c, pl.batch, ml, pl.n, size)
    g.Amul = _operand(a_mul[1], a_mul[0], pl.batch, ml, kl, size) if a_mul else _operand(None, "", [], [], [], size)
    g.Cmul = _operand(c_mul[1], c_mul[0], pl.batch, ml, pl.n, size) if c_mul else _operand(None, "", [], [], [], size)
    g.Cadd = _operand(c_add[1], c_add[0], pl.batch, ml, pl.n, size) if c_add else _operand(None, "", [], [], [], size)
    g.alpha, g.relu, g.accumulate, g.split_k = float(alpha), int(bool(relu)), int(accumulate), int(split_k)
    g.precision = precision
    return g


def _fold_k_outer_into_batch(g: Gemm, dry: bool = False) -> int:
    """A reduction index (b, n) whose inner level is contiguous but of odd length (N = 883 nodes) cannot
    be walked in aligned 32-wide k-blocks.  Such a problem equals a batch over the outer level whose
    members all accumulate into the same output: K' = d1, batch' = K / d1, C batch stride 0 with atomic
    adds -- or, for a reproducible result, one output slice per member (the caller sets C's batch stride
    and sums the slices afterwards).  Restores the 16-byte operand load paths for the pre_conv weight
    gradient.  Rewrites g in place (unless ``dry``); returns the number of members (0 = not applicable)."""
    if g.batch != 1 or g.relu or g.Cadd.ptr or g.Cmul.ptr or g.Amul.ptr:
        return 0
    ka, kb = g.A.col, g.B.row
    d1 = max(ka.d1, kb.d1)
    if d1 < 64 or g.K % d1 or any(ax.d1 not in (1, d1) for ax in (ka, kb)):
        return 0
    if not ((ka.d1 == d1 and ka.s1 == 1) or (kb.d1 == d1 and kb.s1 == 1)):
        return 0
    q = g.K // d1
    if q < 2:
        return 0
    if dry:
        return q

    def levels(ax):
        return (ax.s0, ax.s1) if ax.d1 == d1 else (ax.s0 * d1, ax.s0)

    (oa, ia), (ob, ib) = levels(ka), levels(kb)
    g.A.batch, g.A.col = Axis(oa, 0, 1, 0), Axis(ia, 0, 1, 0)
    g.B.batch, g.B.row = Axis(ob, 0, 1, 0), Axis(ib, 0, 1, 0)
    g.C.batch = Axis(0, 0, 1, 0)
    g.batch, g.K = q, d1
    g.accumulate, g.split_k = ACC_ATOMIC, 1
    return q


def even_split(K: int, split: int) -> int:
    """Largest split count <= ``split`` for which neither engine (k-blocks of 32 on the tensor cores, of 16
    on the CUDA cores) is left with an empty K slice -- every slice of a two-stage reduction is written."""
    def eff(bk, s):
        kb = (K + bk - 1) // bk
        s = min(s, kb)
        per = (kb + s - 1) // s
        return (kb + per - 1) // per
    split = max(1, split)
    while True:
        e = min(eff(32, split), eff(16, split))
        if e == split:
            return split
        split = e


def auto_split_k(M: int, N: int, K: int, batch: int) -> int:
    """Split the reduction when the output grid cannot fill the 148 SMs (weight gradients:
    tiny M x N, K = batch * nodes * time)."""
    tiles = ((M + 63) // 64) * ((N + 63) // 64) * batch
    if tiles >= 2 * NUM_SMS or K < 2048:
        return 1
    return max(1, min((4 * NUM_SMS + tiles - 1) // tiles, K // 512))


def output_shape(spec: str, A, B):
    pl = plan(spec)
    size = _sizes(pl, A, B, ())
    return tuple(size[l] for l in pl.c)


def _prefers_swap(spec: str, out, A, B) -> bool:
    """True when the output's memory-innermost index belongs to the FIRST operand only (it would be an
    m index: strided epilogue stores) -- the caller then swaps the operands."""
    pl = plan(spec)
    if not pl.c or not pl.m or not pl.n:
        return False
    # small outputs (weight gradients: K = all rows of the batch): the epilogue is irrelevant, what counts
    # is how much of the 128 x {32,64,128} MMA tile is padding -- e.g. (64 x 96) fills 37 % of a
    # 128 x 128 tile, (96 x 64) fills 75 % of a 128 x 64 one
    size = _sizes(pl, A, B, ())
    M, N = _prod(pl.m, size), _prod(pl.n, size)
    if M * N <= (1 << 16):
        def padded(m, n):
            bn = 128 if n > 64 else (64 if n > 32 else 32)
            return ((m + 127) // 128 * 128) * ((n + bn - 1) // bn * bn)
        if padded(N, M) * 4 <= padded(M, N) * 3:
            return True
        if padded(M, N) * 4 <= padded(N, M) * 3:
            return False
    if out is not None:
        strides = {l: out.stride(i) for i, l in enumerate(pl.c) if out.shape[i] > 1}
        if not strides:
            return False
        inner = min(strides, key=strides.get)
    else:
        inner = pl.c[-1]
    return inner in pl.m


def contract_raw(spec: str, A: torch.Tensor, B: torch.Tensor, *, out: Optional[torch.Tensor] = None,
                 a_mul: Optional[Tuple[str, torch.Tensor]] = None,
                 c_add: Optional[Tuple[str, torch.Tensor]] = None,
                 c_mul: Optional[Tuple[str, torch.Tensor]] = None, alpha: float = 1.0,
                 relu: bool = False, accumulate: int = ACC_STORE, split_k: Optional[int] = None,
                 precision=None, tag: Optional[str] = None, reproducible: Optional[bool] = None,
                 chain: int = 0) -> torch.Tensor:
    """Enqueue one contraction on the current CUDA stream.  No autograd.

    ``reproducible``: a reduction that is split over K (or whose outer K level is folded into a batch)
    writes one partial result per split into a workspace and sums the slices in a fixed order
    (``dstag_sum_slices``) instead of adding atomically.  Every forward product runs this way
    (``None`` = the module default, see ``set_reproducible``).

    ``chain``: longest accumulation chain (in k) the tensor-core engine may run in its truncating fp32
    accumulator before folding the partial result into registers (0 = unlimited; honoured for N <= 32)."""
    if a_mul is None:
        # a training step repeats the same ~100 contractions with the same layouts: the orientation decision
        # and the problem sizes are memoised per (spec, layouts) -- host time matters at the small graph
        # sizes, where the step is launch bound
        skey = (spec, A.shape, A.stride(), B.shape, B.stride(), _layout_key(out))
        swap = _swap_cache.get(skey)
        if swap is None:
            swap = _prefers_swap(spec, out, A, B)
            if len(_swap_cache) < 4096:
                _swap_cache[skey] = swap
        if swap:
            # C = A.B is symmetric in its operands; put the one that owns the output's innermost index
            # second, so that index becomes the GEMM's n (contiguous, coalesced epilogue stores)
            ins, res = spec.replace(" ", "").split("->")
            sa, sb = ins.split(",")
            return contract_raw(f"{sb},{sa}->{res}", B, A, out=out, c_add=c_add, c_mul=c_mul, alpha=alpha, relu=relu,
                                accumulate=accumulate, split_k=split_k, precision=precision, tag=tag,
                                reproducible=reproducible, chain=chain)
    if _FP32_SPECS and precision is None and any(t in spec or (tag and t in tag) for t in _FP32_SPECS):
        precision = _lib.PREC_FP32          # debugging aid: DSTAG_FP32_SPECS=substring,... forces the exact engine
    lib = _lib.load()
    _check(A, "A"), _check(B, "B")
    for opt in (a_mul, c_add, c_mul):
        if opt is not None:
            _check(opt[1], "epilogue/prologue operand")
    _lib.same_device(A, B, out, *(opt[1] for opt in (a_mul, c_add, c_mul) if opt is not None))
    if out is None:
        if accumulate != ACC_STORE:
            raise ValueError("accumulating contraction needs an explicit out=")
        out = torch.empty(output_shape(spec, A, B), device=A.device, dtype=torch.float32)
    _check(out, "out")
    dkey = (spec, A.shape, B.shape, out.shape)
    dims = _dims_cache.get(dkey)
    if dims is None:
        pl = plan(spec)
        size = _sizes(pl, A, B, ((pl.c, out),))                    # also validates the shapes against the spec
        dims = (_prod(pl.m, size), _prod(pl.n, size), _prod(pl.k, size), _prod(pl.batch, size))
        if len(_dims_cache) < 4096:
            _dims_cache[dkey] = dims
    M, N, K, nb = dims
    if split_k is None:
        split_k = 1 if relu else auto_split_k(M, N, K, nb)
    if reproducible is None:
        reproducible = _reproducible
    # two-stage (workspace + fixed-order sum) instead of atomics: plain stores into a dense output only
    rv = rows_view(out) if (reproducible and accumulate == ACC_STORE and c_mul is None) else None
    two_stage = rv is not None
    ws = None
    if split_k > 1 and two_stage:
        split_k = even_split(K, split_k)
    if split_k > 1:
        if two_stage:
            ws = _workspace(split_k, out)
            g = build_descriptor(spec, A, B, ws[0], a_mul, c_add, c_mul, alpha, relu, ACC_STORE, split_k, precision)
            g.split_stride = ws.stride(0)
        else:
            if accumulate == ACC_STORE:
                out.zero_()
            # (ACC_ADD: existing contents are kept; partial sums are added atomically)
            accumulate = ACC_ATOMIC
            g = build_descriptor(spec, A, B, out, a_mul, c_add, c_mul, alpha, relu, accumulate, split_k, precision)
    else:
        g = build_descriptor(spec, A, B, out, a_mul, c_add, c_mul, alpha, relu, accumulate, split_k, precision)
        q = _fold_k_outer_into_batch(g, dry=True)
        if q:
            if two_stage:
                ws = _workspace(q, out)
                g = build_descriptor(spec, A, B, ws[0], a_mul, c_add, c_mul, alpha, relu, ACC_STORE, 1, precision)
                _fold_k_outer_into_batch(g)
                g.accumulate = ACC_STORE
                g.C.batch = Axis(ws.stride(0), 0, 1, 0)            # one output slice per batch member
            else:
                _fold_k_outer_into_batch(g)
                if accumulate == ACC_STORE:
                    out.zero_()
    g.chain = int(chain)
    st = _lib.stream_ptr(A)
    with _lib.on_device(st):
        if _lib._prof is None:
            _lib.check(lib.dstag_gemm(C.byref(g), st), f"dstag_gemm('{spec}')")
        else:
            nbytes = 4 * (A.numel() + B.numel() + out.numel())
            with _lib.prof_range(tag or ("gemm:" + spec), 2 * M * N * K * nb, nbytes):
                _lib.check(lib.dstag_gemm(C.byref(g), st), f"dstag_gemm('{spec}')")
        if ws is not None:
            rows, cols, ld = rv
            _lib.check(lib.dstag_sum_slices(ws.data_ptr(), ws.stride(0), ws.shape[0], out.data_ptr(), rows, cols, ld, st),
                       "dstag_sum_slices")
    return out


# ----------------------------------------------------------------------------------------
# zero-initialised scratch (accumulation targets of the backward kernels)
# ----------------------------------------------------------------------------------------
class ZeroArena:
    """One buffer per device, cleared ONCE per training step, from which the backward takes its zero-initialised
    accumulators (bias / LayerNorm / table gradients that kernels add into): ~80 separate fill kernels per step
    become one.  ``step.TrainStep`` resets it at the top of every step; when no arena is armed (plain autograd use
    of the modules) ``zeros`` is ``torch.zeros``.  Slices are handed out in order and never reused within a step."""

    ALIGN = 32                       # floats: every slice starts on a 128-byte boundary

    def __init__(self, device, numel: int):
        self.buf = torch.zeros(numel, device=device, dtype=torch.float32)
        self.used = 0
        self.peak_demand = 0

    def reset(self):
        if self.used:
            self.buf[:self.used].zero_()
        self.used = 0
        self.peak_demand = 0

    def take(self, shape):
        n = 1
        for d in shape:
            n *= d
        need = (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.peak_demand += need
        if self.used + need > self.buf.numel():
            return None
        out = self.buf[self.used:self.used + n].view(shape)
        self.used += need
        return out


_arenas = {}


def arm_zero_arena(device, numel: int) -> ZeroArena:
    a = _arenas[torch.device(device).index] = ZeroArena(device, numel)
    return a


def disarm_zero_arena(device):
    _arenas.pop(torch.device(device).index, None)


def zeros(shape, device) -> torch.Tensor:
    """fp32 zeros of ``shape``: a slice of the step's arena when one is armed on ``device``, else ``torch.zeros``."""
    if not isinstance(shape, (tuple, list)):
        shape = (shape,)
    a = _arenas.get(torch.device(device).index)
    if a is not None:
        t = a.take(tuple(shape))
        if t is not None:
            return t
    return torch.zeros(tuple(shape), device=device, dtype=torch.float32)


# ----------------------------------------------------------------------------------------
# small helpers shared with ops.py
# ----------------------------------------------------------------------------------------
def colsum(t2d: torch.Tensor, rows: int, cols: int, ld: Optional[int] = None) -> torch.Tensor:
    """sum over rows of a (rows, cols) matrix with leading dimension ld -> (cols,)."""
    lib = _lib.load()
    _check(t2d, "colsum input")
    out = zeros((cols,), t2d.device)
    _lib.check(lib.dstag_colsum(t2d.data_ptr(), out.data_ptr(), rows, cols, cols if ld is None else ld,
                                _lib.stream_ptr(t2d)), "dstag_colsum")
    return out


def padded_ld(n: int) -> int:
    """Leading dimension for rows of n floats: the next multiple of 4 (16-byte rows), so that the
    contraction engine can use its vector / TMA load paths on odd graph sizes (N = 883, 170, ...)."""
    return (n + 3) // 4 * 4


def empty_rows(shape, device) -> torch.Tensor:
    """Uninitialised fp32 tensor of ``shape`` whose innermost rows are padded to 16 bytes
    (a view ``buf[..., :n]``; contiguous when n % 4 == 0)."""
    n = shape[-1]
    ld = padded_ld(n)
    buf = torch.empty(tuple(shape[:-1]) + (ld,), device=device, dtype=torch.float32)
    return buf if ld == n else buf[..., :n]


def is_row_padded(t: torch.Tensor) -> bool:
    """True for the layout ``empty_rows`` produces (or plain contiguous): inner stride 1, outer dims
    contiguous over a leading dimension >= the row length."""
    if t.dim() == 0 or t.stride(-1) != 1:
        return False
    if t.dim() == 1:
        return True
    ld = t.stride(-2)
    if ld < t.shape[-1]:
        return False
    expect = ld
    for d in range(t.dim() - 2, -1, -1):
        if t.shape[d] != 1 and t.stride(d) != expect:
            return False
        expect *= t.shape[d]
    return True


def dense_rows(t: torch.Tensor) -> torch.Tensor:
    """``t`` itself when it is contiguous or row padded, else a contiguous copy."""
    return t if is_row_padded(t) else t.contiguous()


def empty_like_rows(t: torch.Tensor) -> torch.Tensor:
    """Gradient buffer with the same (possibly row-padded) layout as ``t``."""
    if t.is_contiguous() or not is_row_padded(t):
        return torch.empty(t.shape, device=t.device, dtype=torch.float32)
    return torch.empty_strided(t.shape, t.stride(), device=t.device, dtype=torch.float32)


def relu_bwd(y: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    dy = dy.contiguous()
    dx = torch.empty_like(dy)
    _lib.check(lib.dstag_relu_bwd(y.data_ptr(), dy.data_ptr(), dx.data_ptr(), y.numel(),
                                  _lib.stream_ptr(y)), "dstag_relu_bwd")
    return dx


def reduce_to(sub_out: str, grad: torch.Tensor, sub_keep: str) -> torch.Tensor:
    """Sum a contiguous gradient with subscripts ``sub_out`` down to ``sub_keep`` (a subset,
    same relative order).  The big reduction (all leading dims) runs in dstag_colsum; only a
    tiny remainder, if any, is left to torch."""
    assert grad.is_contiguous()
    first = min(sub_out.index(l) for l in sub_keep)
    lead = 1
    for s in grad.shape[:first]:
        lead *= s
    tail_shape = grad.shape[first:]
    cols = 1
    for s in tail_shape:
        cols *= s
    red = colsum(grad, lead, cols).view(tail_shape) if lead > 1 else grad.reshape(tail_shape)
    drop = [i for i, l in enumerate(sub_out[first:]) if l not in sub_keep]
    if drop:
        red = red.sum(dim=drop)
    return red


def _gram_weight_grad(spec, A, dC, B, am, alpha):
    """Weight gradients whose reduction runs over the rows of two row-contiguous activations take the tall-skinny
    Gram kernel (csrc/gram_tc.cu) instead of a split-K contraction.  Today: the Theta mix of cheb_conv_withSAt,
    dTheta[k, f, c] = sum_{x,t} Z[k, x, t, f] dG[x, t, c] (model/DSTAGNN_my.py:131) with 32 features and channels.
    Returns None when the shape is not covered."""
    if spec != "kxtf,kfc->xtc" or am is not None or alpha != 1.0 or os.environ.get("DSTAG_GRAM", "1") == "0":
        return None
    if get_precision() == "fp32" or not (A.is_contiguous() and dC.is_contiguous()):
        return None
    K, X, T, Fin = A.shape
    Cc = dC.shape[-1]
    lib = _lib.load()
    if Fin != 32 or Cc != 32 or K * Fin > 128 or X * T * 32 >= 2 ** 31 - 64 or not lib.dstag_gram_supported(K, Cc, 1, 1):
        return None
    st = _lib.stream_ptr(A)
    with _lib.on_device(st):
        slices = lib.dstag_gram_slices()
    part = torch.empty((slices, 128, Cc), device=A.device, dtype=torch.float32)
    _lib.check(lib.dstag_gram_rows(A.data_ptr(), dC.data_ptr(), part.data_ptr(), X * T, 1, 1, K, X * T * Fin, Cc, st),
               "dstag_gram_rows")
    dt = torch.empty((128, Cc), device=A.device, dtype=torch.float32)
    _lib.check(lib.dstag_sum_slices(part.data_ptr(), part.stride(0), slices, dt.data_ptr(), 128, Cc, Cc, st),
               "dstag_sum_slices")
    return dt[:K * Fin].view(K, Fin, Cc)


def _thin_args(spec, A, B, amul, alpha, cadd, cadd_sub, pad_rows):
    """Operands of the thin-feature kernels (csrc/thin.cu) for the first block's Theta mix and 1x1 residual conv
    (model/DSTAGNN_my.py:131, 228 with in_channels = 1): (a_plane, taps, F, w strides, R, C), or None when the
    contraction is not one of them (the engine takes it then)."""
    if amul is not None or alpha != 1.0 or pad_rows or os.environ.get("DSTAG_THIN", "1") == "0":
        return None
    if cadd is not None and (cadd_sub != "c" or not cadd.is_contiguous() or cadd.data_ptr() % 16):
        return None
    if not (A.is_contiguous() and A.dtype == torch.float32 and B.dtype == torch.float32):
        return None
    if spec == "kxtf,kfc->xtc" and B.is_contiguous():
        K, X, T, Fin = A.shape
        Cc = B.shape[-1]
        args = (X * T * Fin, K, Fin, Cc, 1, X * T, Cc)
    elif spec == "xtf,cf->xtc":
        X, T, Fin = A.shape
        Cc = B.shape[0]
        args = (0, 1, Fin, B.stride(1), B.stride(0), X * T, Cc)
    else:
        return None
    if not _lib.load().dstag_thin_supported(args[1] * args[2], Cc):
        return None
    return args


def _thin_forward(args, A, B, cadd, relu):
    a_plane, taps, Fin, w_js, w_cs, R, Cc = args
    out = torch.empty(A.shape[-3:-1] + (Cc,), device=A.device, dtype=torch.float32)
    _lib.check(_lib.load().dstag_thin_expand(A.data_ptr(), a_plane, taps, Fin, B.data_ptr(), w_js, w_cs,
                                             cadd.data_ptr() if cadd is not None else 0, out.data_ptr(), R, Cc,
                                             int(relu), _lib.stream_ptr(A)), "dstag_thin_expand")
    return out


def _thin_backward(args, A, B, dC, y, need_dA, need_dB, need_dadd):
    """One pass over dC: relu mask, dA (thin planes), and per-CTA partials of dB / the bias gradient."""
    a_plane, taps, Fin, w_js, w_cs, R, Cc = args
    J = taps * Fin
    lib = _lib.load()
    st = _lib.stream_ptr(A)
    if dC.data_ptr() % 16:                 # (an offset view: the kernel reads float4 rows)
        dC = dC.clone()
    dA = torch.empty_like(A) if need_dA else None
    part = None
    if need_dB or need_dadd:
        part = torch.empty((lib.dstag_thin_slices(), J + 1, Cc), device=A.device, dtype=torch.float32)
    _lib.check(lib.dstag_thin_bwd(A.data_ptr(), a_plane, taps, Fin, B.data_ptr(), w_js, w_cs, dC.data_ptr(),
                                  y.data_ptr() if y is not None else 0, dA.data_ptr() if need_dA else 0,
                                  part.data_ptr() if part is not None else 0, R, Cc, st), "dstag_thin_bwd")
    dB = dadd = None
    if part is not None:
        dt = torch.empty((J + 1, Cc), device=A.device, dtype=torch.float32)
        _lib.check(lib.dstag_sum_slices(part.data_ptr(), part.stride(0), part.shape[0], dt.data_ptr(), J + 1, Cc, Cc,
                                        st), "dstag_sum_slices")
        if need_dB:
            # (K, F, C) as stored; the conv weight is (C, F): a view when F == 1
            dB = dt[:J].view(B.shape) if (B.dim() == 3 or Fin == 1) else dt[:J].t().contiguous()
        if need_dadd:
            dadd = dt[J]
    return dA, dB, dadd


class _Contract(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, B, cadd, amul, spec, cadd_sub, amul_sub, alpha, relu, pad_rows):
        ctx.thin = _thin_args(spec, A, B, amul, alpha, cadd, cadd_sub, pad_rows)
        if ctx.thin is not None:
            out = _thin_forward(ctx.thin, A, B, cadd, relu)
            ctx.spec, ctx.cadd_sub, ctx.amul_sub, ctx.alpha, ctx.relu = spec, cadd_sub, amul_sub, alpha, relu
            ctx.has_cadd = cadd is not None
            ctx.save_for_backward(A, B, amul, out if relu else None)
            return out
        out = None
        if pad_rows:
            assert not relu
            out = empty_rows(output_shape(spec, A, B), A.device)
        out = contract_raw(spec, A, B, out=out, a_mul=(amul_sub, amul) if amul is not None else None,
                           c_add=(cadd_sub, cadd) if cadd is not None else None, alpha=alpha, relu=relu,
                           reproducible=True)
        ctx.spec, ctx.cadd_sub, ctx.amul_sub, ctx.alpha, ctx.relu = spec, cadd_sub, amul_sub, alpha, relu
        ctx.has_cadd = cadd is not None
        ctx.save_for_backward(A, B, amul, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, dC):
        A, B, amul, out = ctx.saved_tensors
        pl = plan(ctx.spec)
        dC = dense_rows(dC)
        if ctx.thin is not None:
            dA, dB, dadd = _thin_backward(ctx.thin, A, B, dC.contiguous(), out if ctx.relu else None,
                                          ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                          ctx.has_cadd and ctx.needs_input_grad[2])
            return dA, dB, dadd, None, None, None, None, None, None, None
        if ctx.relu:
            dC = relu_bwd(out, dC)
        dA = dB = dadd = None
        am = (ctx.amul_sub, amul) if amul is not None else None
        if ctx.needs_input_grad[0]:
            dA = empty_like_rows(A)        # gradients keep the (row-padded) layout of their tensors
            contract_raw(f"{pl.c},{pl.b}->{pl.a}", dC, B, out=dA, c_mul=am, alpha=ctx.alpha)
        if ctx.needs_input_grad[1]:
            dB = _gram_weight_grad(ctx.spec, A, dC, B, am, ctx.alpha)
            if dB is None:
                dB = empty_like_rows(B)
                contract_raw(f"{pl.a},{pl.c}->{pl.b}", A, dC, out=dB, a_mul=am, alpha=ctx.alpha)
        if ctx.has_cadd and ctx.needs_input_grad[2]:
            dadd = reduce_to(pl.c, dC.contiguous(), ctx.cadd_sub)
        return dA, dB, dadd, None, None, None, None, None, None, None


def contract(spec: str, A: torch.Tensor, B: torch.Tensor, *, c_add=None, a_mul=None, alpha: float = 1.0,
             relu: bool = False, pad_rows: bool = False) -> torch.Tensor:
    """Differentiable contraction.  ``c_add=(subscripts, tensor)`` is a differentiable additive
    epilogue term (bias vector / bias matrix); ``a_mul=(subscripts, tensor)`` is a constant
    Hadamard factor applied to the first operand while it is loaded."""
    cadd_sub, cadd = c_add if c_add is not None else (None, None)
    amul_sub, amul = a_mul if a_mul is not None else (None, None)
    return _Contract.apply(A, B, cadd, amul, spec, cadd_sub, amul_sub, float(alpha), bool(relu), bool(pad_rows))
