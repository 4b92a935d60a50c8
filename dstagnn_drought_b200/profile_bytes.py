"""Algorithmic HBM bytes of every non-GEMM entry point (each operand read once, each result written once), as a
function of the raw ctypes arguments.  Used only by the CUDA-event profiler of ``_lib`` (bench.py --verbose /
profiles/): achieved GB/s of a kernel = these bytes / its measured time.  Gather traffic that is served by L2
(edge kernels re-reading rows of a per-sample matrix that fits the 126 MB L2) is NOT counted: the figure is the
roofline numerator, not the traffic."""


def _ln_fwd(a):
    rows, L, tab_mod = a[15], a[16], a[6]
    return 4 * rows * L * (2 + (1 if (a[5] and tab_mod == rows) else 0))


def _ln_bwd(a):
    rows, L, tab_mod = a[17], a[18], a[8]
    return 4 * rows * L * (3 + (1 if (a[7] and tab_mod == rows) else 0))


def _tattn(a, bwd):
    B, F, H, T, dk = (a[5:10] if not bwd else a[6:11])
    qkv, sc, ctx = 3 * T * H * dk, H * T * T, T * H * dk
    return 4 * B * F * ((qkv + sc * (1 + (1 if a[1] else 0)) + ctx) if not bwd else (2 * qkv + 3 * sc + ctx))


ALGO_BYTES = {
    "dstag_layernorm_fwd": _ln_fwd,
    "dstag_layernorm_bwd": _ln_bwd,
    "dstag_colsum": lambda a: 4 * a[2] * a[3],
    "dstag_sum_slices": lambda a: 4 * a[4] * a[5] * (a[2] + 1),
    "dstag_transpose": lambda a: 8 * a[4] * a[5] * a[6],
    "dstag_softmax_fwd": lambda a: 8 * a[1] * a[2],
    "dstag_softmax_bwd": lambda a: 12 * a[4] * a[5],
    "dstag_relu_bwd": lambda a: 12 * a[3],
    "dstag_sat_softmax_fwd": lambda a: 4 * (a[5] * a[7] * a[6] * a[9] * (2 if a[4] else 1) + a[5] * a[6] * 2 * a[7] * a[8]),
    "dstag_tattn_fwd": lambda a: _tattn(a, False),
    "dstag_tattn_bwd": lambda a: _tattn(a, True),
    "dstag_gtu_gate_fwd": lambda a: 4 * a[2] * a[4] * 3 * a[3],
    "dstag_gtu_gate_bwd": lambda a: 4 * a[4] * a[5] * (3 * a[6] + 2 * (a[6] + 2 * a[7])),
    "dstag_gtu_gate_bwd_sv": lambda a: 4 * a[4] * a[5] * (3 * a[6] + 2 * (a[6] + 2 * a[7])),
    # fused GTU: g in, cat (+ saved gate factors) out; saved factors + dcat in, du + dg out
    "dstag_gtu_conv_gate_fwd": lambda a: 4 * a[5] * a[6] * (a[7] + (a[7] - a[8] + 1) * (1 + (2 if a[4] else 0))),
    "dstag_gtu_gate_conv_bwd": lambda a: 4 * a[6] * a[7] * ((a[8] - a[9] + 1) * 5 + a[8] * (2 if a[12] else 1)),
    "dstag_gram_rows": lambda a: 4 * a[3] * (a[4] * 32 + a[5] * a[8]),
    # thin contractions: the (R, C) tensor once (+ the relu mask), the J thin planes once
    "dstag_thin_expand": lambda a: 4 * a[9] * (a[10] + a[2] * a[3]),
    "dstag_thin_bwd": lambda a: 4 * a[11] * (a[12] * (2 if a[8] else 1) + a[2] * a[3] * ((1 if a[10] else 0) + (1 if a[9] else 0))),
    "dstag_outer_sum": lambda a: 4 * a[4] * (a[5] + a[6]) * a[7],
    "dstag_tail_fwd": lambda a: 4 * a[9] * a[10] * a[11] * (3 + (1 if a[1] else 0)),
    "dstag_tail_bwd": lambda a: 4 * a[12] * a[13] * a[14] * (5 + (2 if a[2] else 0)),
    "dstag_tail_fused_fwd": lambda a: 4 * a[11] * a[12] * ((3 * a[13] - 12) + a[13] * (2 + (1 if a[3] else 0))),
    "dstag_tail_fused_bwd": lambda a: 4 * a[15] * a[16] * (2 * (3 * a[17] - 12) + a[17] * (4 + (2 if a[4] else 0))),
    "dstag_smooth_l1_fwd": lambda a: 8 * a[4],
    "dstag_smooth_l1_bwd": lambda a: 12 * a[4],
    "dstag_adam_step": lambda a: 28 * a[4],
    "dstag_bias_transpose": lambda a: 4 * (a[4] * a[4] * (1 + a[2]) + a[2] * a[4] * a[5]),
    "dstag_bias_transpose_bwd": lambda a: 4 * (a[4] * a[4] * (1 + a[3]) + a[3] * a[4] * a[5]),
    "dstag_edge_probs": lambda a: 8 * a[5] * a[9],
    "dstag_spmm_expand": lambda a: 4 * a[6] * a[8] * a[9] * (1 + a[7]),
    "dstag_spmm_reduce": lambda a: 4 * a[6] * a[8] * a[9] * (1 + a[7]),
    "dstag_edge_dots": lambda a: 4 * a[7] * a[9] * a[10] * (1 + a[8]),
    "dstag_scale_keys": lambda a: 8 * a[3] * a[4] * a[5] * a[6],
    "dstag_center_queries": lambda a: 12 * a[3] * a[4] * a[5] * a[6],
    "dstag_dqk_combine": lambda a: 24 * a[12] * a[13] * a[14] * a[15],
    "dstag_dbias_sparse": lambda a: 4 * a[12] * a[8] * (a[7] + 1),
    "dstag_tick": lambda a: 16 * a[1],
}


# multiply-add = 2 flops; only the entry points that run on the tensor cores outside dstag_gemm
ALGO_FLOPS = {
    "dstag_gram_rows": lambda a: 2 * a[3] * a[5] * a[6] * 32 * a[8],
    "dstag_gtu_conv_gate_fwd": lambda a: 2 * a[5] * (a[7] - a[8] + 1) * 2 * a[6] * a[8] * a[6],
    "dstag_gtu_gate_conv_bwd": lambda a: 2 * a[6] * (a[8] - a[9] + 1) * 2 * a[7] * a[9] * a[7],
}


def algo_flops(name, args):
    f = ALGO_FLOPS.get(name)
    if f is None:
        return 0
    try:
        vals = [(x.value if hasattr(x, "value") else x) for x in args]
        return int(f([0 if v is None else v for v in vals]))
    except Exception:
        return 0


def algo_bytes(name, args):
    f = ALGO_BYTES.get(name)
    if f is None:
        return 0
    try:
        vals = [(x.value if hasattr(x, "value") else x) for x in args]
        vals = [0 if v is None else v for v in vals]
        return int(f(vals))
    except Exception:
        return 0
