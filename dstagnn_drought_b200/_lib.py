"""ctypes binding of libdstagnn_b200.so (the C ABI declared in include/dstagnn_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call
fails, a RuntimeError is raised.  The library is built in-tree by
``__graft_entry__.build()`` (nvcc, sm_100a) so that it travels with the repo.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdstagnn_b200.so")
# A/B benchmarking of engine variants on the same box: DSTAG_LIB_PATH selects another build of the same ABI
LIB_PATH = os.environ.get("DSTAG_LIB_PATH", LIB_PATH)
ABI_VERSION = 19

ACC_STORE, ACC_ADD, ACC_ATOMIC = 0, 1, 2
PREC_FP32, PREC_TF32X3, PREC_TF32 = 0, 1, 2
ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_SMEM = -1, -2, -3      # DSTAG_ERR_* of include/dstagnn_b200.h


class Axis(C.Structure):
    _fields_ = [("s0", C.c_int64), ("s1", C.c_int64), ("d1", C.c_int32), ("pad_", C.c_int32)]


class Operand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("batch", Axis), ("row", Axis), ("col", Axis)]


class Gemm(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("batch", C.c_int32),
                ("A", Operand), ("B", Operand), ("C", Operand),
                ("Amul", Operand), ("Cmul", Operand), ("Cadd", Operand),
                ("alpha", C.c_float), ("relu", C.c_int32), ("accumulate", C.c_int32),
                ("split_k", C.c_int32), ("precision", C.c_int32), ("chain", C.c_int32),
                ("split_stride", C.c_int64)]


class Copy(C.Structure):     # dstag_copy_t
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("n", C.c_int32 * 4), ("src_stride", C.c_int64 * 4),
                ("dst_stride", C.c_int64 * 4)]


_F = C.c_void_p      # device pointers travel as integers (tensor.data_ptr())
_i32, _i64, _f32, _u64 = C.c_int32, C.c_int64, C.c_float, C.c_uint64

# name -> (restype, argtypes); mirrors include/dstagnn_b200.h one to one
SIGNATURES = {
    "dstag_abi_version": (C.c_int, []),
    "dstag_last_error": (C.c_char_p, []),
    "dstag_launch_count": (_i64, []),
    "dstag_reset_launch_count": (None, []),
    "dstag_gemm": (C.c_int, [C.POINTER(Gemm), _F]),
    "dstag_gemm_uses_tensor_cores": (C.c_int, [C.POINTER(Gemm)]),
    "dstag_layernorm_fwd": (C.c_int, [_F, _i64, _i64, _i32, _i64, _F, _i32, _i64, _F, _F, _f32, _F, _i64, _F, _F,
                                      _i32, _i32, _f32, _u64, _F, _F]),
    "dstag_layernorm_bwd": (C.c_int, [_F, _i64, _F, _i64, _i64, _i32, _i64, _F, _i32, _i64, _F, _F, _F, _F, _i64,
                                      _F, _F, _i32, _i32, _f32, _u64, _F, _F]),
    "dstag_sum_slices": (C.c_int, [_F, _i64, _i32, _F, _i64, _i64, _i64, _F]),
    "dstag_colsum": (C.c_int, [_F, _F, _i64, _i64, _i64, _F]),
    "dstag_transpose": (C.c_int, [_F, _i64, _F, _i64, _i32, _i32, _i32, _i32, _i32, _F]),
    "dstag_softmax_fwd": (C.c_int, [_F, _i64, _i32, _i32, _F]),
    "dstag_softmax_bwd": (C.c_int, [_F, _F, _F, _i64, _i64, _i32, _i32, _F]),
    "dstag_relu_bwd": (C.c_int, [_F, _F, _F, _i64, _F]),
    "dstag_sat_softmax_fwd": (C.c_int, [_F, _F, _F, _F, _F, _i32, _i32, _i32, _i32, _i32, _F]),
    "dstag_tattn_fwd": (C.c_int, [_F, _F, _i32, _F, _F, _i32, _i32, _i32, _i32, _i32, _F]),
    "dstag_tattn_bwd": (C.c_int, [_F, _F, _F, _F, _F, _F, _i32, _i32, _i32, _i32, _i32, _F]),
    "dstag_gtu_gate_fwd": (C.c_int, [_F, _F, _i64, _i32, _i32, _i32, _i32, _F]),
    "dstag_gtu_gate_bwd": (C.c_int, [_F, _F, _F, _F, _i64, _i32, _i32, _i32, _i32, _i32, _F]),
    "dstag_gtu_gate_bwd_sv": (C.c_int, [_F, _F, _F, _F, _i64, _i32, _i32, _i32, _i32, _i32, _F]),
    "dstag_gtu_tc_supported": (C.c_int, [_i32, _i32, _i32, _i32]),
    "dstag_gtu_conv_gate_fwd": (C.c_int, [_F, _F, _F, _F, _F, _i64, _i32, _i32, _i32, _i32, _i32, _F]),
    "dstag_gtu_gate_conv_bwd": (C.c_int, [_F, _F, _F, _F, _F, _F, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _F]),
    "dstag_gtu_bwd_ws_rows": (C.c_int, []),
    "dstag_gram_slices": (C.c_int, []),
    "dstag_gram_supported": (C.c_int, [_i32, _i32, _i32, _i32]),
    "dstag_gram_rows": (C.c_int, [_F, _F, _F, _i64, _i32, _i32, _i32, _i64, _i32, _F]),
    "dstag_thin_supported": (C.c_int, [_i32, _i32]),
    "dstag_thin_slices": (C.c_int, []),
    "dstag_thin_expand": (C.c_int, [_F, _i64, _i32, _i32, _F, _i64, _i64, _F, _F, _i64, _i32, _i32, _F]),
    "dstag_thin_bwd": (C.c_int, [_F, _i64, _i32, _i32, _F, _i64, _i64, _F, _F, _F, _F, _i64, _i32, _F]),
    "dstag_outer_sum_supported": (C.c_int, [_i32, _i32, _i32]),
    "dstag_outer_sum": (C.c_int, [_F, _F, _F, _F, _i64, _i32, _i32, _i32, _F]),
    "dstag_tail_fwd": (C.c_int, [_F, _F, _F, _F, _F, _f32, _F, _F, _F, _i64, _i32, _i32, _f32, _u64, _F, _F]),
    "dstag_tail_bwd": (C.c_int, [_F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _i64, _i32, _i32, _f32,
                                 _u64, _F, _F]),
    "dstag_tail_fused_supported": (C.c_int, [_i32, _i32]),
    "dstag_tail_fused_fwd": (C.c_int, [_F, _F, _F, _F, _F, _F, _F, _f32, _F, _F, _F, _i64, _i32, _i32, _f32, _u64, _F,
                                       _F]),
    "dstag_tail_fused_bwd": (C.c_int, [_F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _i64, _i32, _i32,
                                       _f32, _u64, _F, _F]),
    "dstag_smooth_l1_fwd": (C.c_int, [_F, _F, _F, _F, _i64, _F]),
    "dstag_smooth_l1_bwd": (C.c_int, [_F, _F, _F, _F, _i64, _F]),
    "dstag_adam_step": (C.c_int, [_F, _F, _F, _F, _i64, _f32, _f32, _f32, _f32, _f32, _F, _u64, _F]),
    "dstag_gather_flat": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(_i64), C.POINTER(_i64), _i32, _F, _F]),
    "dstag_multi_copy": (C.c_int, [C.POINTER(Copy), _i32, _F]),
    "dstag_tick": (C.c_int, [_F, _i32, _F]),
    "dstag_bias_transpose": (C.c_int, [_F, C.POINTER(C.c_void_p), _i32, _F, _i32, _i32, _F]),
    "dstag_bias_transpose_bwd": (C.c_int, [_F, _F, _F, _i32, _i32, _i32, _F]),
    "dstag_edge_probs": (C.c_int, [_F, _F, _F, _F, _F, _i32, _i32, _i32, _i32, _i32, _F]),
    "dstag_spmm_expand": (C.c_int, [_F, _F, _F, _F, _F, _F, _i32, _i32, _i32, _i32, _i32, _F]),
    "dstag_spmm_reduce": (C.c_int, [_F, _F, _F, _F, _F, _F, _i32, _i32, _i32, _i32, _i32, _F]),
    "dstag_edge_dots": (C.c_int, [_F, _F, _F, _F, _F, _F, _F, _i32, _i32, _i32, _i32, _i32, _F]),
    "dstag_scale_keys": (C.c_int, [_F, _F, _F, _i32, _i32, _i32, _i32, _F]),
    "dstag_center_queries": (C.c_int, [_F, _F, _F, _i32, _i32, _i32, _i32, _F]),
    "dstag_dqk_combine": (C.c_int, [_F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _i32, _i32, _i32, _i32, _i32, _f32,
                                    _F]),
    "dstag_dbias_sparse": (C.c_int, [_F, _F, _F, _F, _F, _F, _F, _i32, _i32, _i32, _i32, _i32, _i32, _F]),
}

_lib = None


def load():
    """Load the shared library once; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root "
            "(needs nvcc with sm_100a support). There is no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    ver = lib.dstag_abi_version()
    if ver != ABI_VERSION:
        raise RuntimeError(f"libdstagnn_b200.so ABI {ver} != expected {ABI_VERSION}; rebuild it")
    _lib = _Profiled(lib)
    return _lib


class _Profiled:
    """Thin proxy over the CDLL: identical calls; when the event profiler is enabled every
    kernel entry point (except dstag_gemm, which contraction.py tags itself) is bracketed by
    a CUDA-event pair named after the function."""

    _PLAIN = ("dstag_abi_version", "dstag_last_error", "dstag_launch_count", "dstag_reset_launch_count",
              "dstag_gemm", "dstag_gemm_uses_tensor_cores", "dstag_tail_fused_supported",
              "dstag_outer_sum_supported", "dstag_gtu_tc_supported", "dstag_gtu_bwd_ws_rows", "dstag_gram_slices",
              "dstag_gram_supported", "dstag_thin_supported", "dstag_thin_slices")

    def __init__(self, lib):
        self._cdll = lib
        for name in SIGNATURES:
            fn = getattr(lib, name)
            setattr(self, name, fn if name in self._PLAIN else self._wrap(name, fn))

    @staticmethod
    def _wrap(name, fn):
        def call(*args):
            with on_device(args[-1]):
                if _prof is None:
                    return fn(*args)
                from .profile_bytes import algo_bytes, algo_flops
                with prof_range(name, algo_flops(name, args), algo_bytes(name, args)):
                    return fn(*args)
        return call


class StreamHandle(C.c_void_p):
    """cudaStream_t as a void* that remembers which device it belongs to."""
    device_index = -1


def stream_ptr(t):
    """cudaStream_t of torch's current stream on t's device, as a void* tagged with the device."""
    import torch
    h = StreamHandle(torch.cuda.current_stream(t.device).cuda_stream)
    h.device_index = t.device.index if t.device.index is not None else torch.cuda.current_device()
    return h


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def on_device(stream):
    """Kernel launches go to the CURRENT CUDA device; the operands' device must be the current one while
    an entry point enqueues (a stream handle of device 1 launched on device 0 is an illegal access, or --
    with peer access enabled -- a silent launch on the wrong GPU).  Returns a no-op guard in the common
    case (already current) and ``torch.cuda.device(idx)`` otherwise, so models living on ``cuda:k`` work
    without the caller having to ``torch.cuda.set_device(k)``."""
    import torch
    idx = getattr(stream, "device_index", -1)
    if idx < 0 or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(idx)


def same_device(*tensors):
    """Raise unless all (non-None) tensors live on one CUDA device."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"dstagnn_drought_b200: operands on different devices ({dev} and {t.device})")
    return dev


# ---- optional per-kernel CUDA-event profiler (bench.py / profiling only) -----------------------
_prof = None          # None = disabled; else {tag: [(start_event, end_event, flops, bytes), ...]}


def prof_enable(on: bool = True):
    global _prof
    _prof = {} if on else None


def prof_results():
    """{tag: (calls, total_ms, flops, bytes)} -- call after torch.cuda.synchronize()."""
    out = {}
    for tag, evs in (_prof or {}).items():
        ms = sum(a.elapsed_time(b) for a, b, _, _ in evs)
        out[tag] = (len(evs), ms, sum(e[2] for e in evs), sum(e[3] for e in evs))
    return out


class prof_range:
    """Records a CUDA-event pair on the current stream around the enclosed launches."""

    def __init__(self, tag, flops=0, nbytes=0):
        self.tag, self.flops, self.nbytes = tag, flops, nbytes

    def __enter__(self):
        if _prof is not None:
            import torch
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _prof is not None:
            self.b.record()
            _prof.setdefault(self.tag, []).append((self.a, self.b, self.flops, self.nbytes))
        return False


def check(rc: int, what: str):
    if rc != 0:
        msg = load().dstag_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(load().dstag_launch_count())


def reset_launch_count():
    load().dstag_reset_launch_count()
