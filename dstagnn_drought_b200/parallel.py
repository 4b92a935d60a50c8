"""Batch-sharded data parallelism: one process per GPU, NCCL all-reduce of gradients.

The reference reduces gradients across replicas inside ``xm.optimizer_step``
(train_DSTAGNN_my.py:148,158).  Here every rank holds a full replica and B_local
samples; the loss is a mean, so gradients are *averaged* (SURVEY.md section 8e).

Gradients live in ONE flat fp32 buffer per block (``p.grad`` are views into it), so a
bucket is reduced in place with a single ``all_reduce`` -- no flatten / unflatten copies.
Blocks finish their backward in reverse order; each bucket's collective is launched from a
post-accumulate hook as soon as its last gradient is written, overlapping NVLink traffic
with the remaining backward kernels.  Dead parameters (EmbedT / residual_conv of blocks
>= 1) are left out so their ``.grad`` stays ``None`` and Adam skips them
exactly as in the reference.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.distributed as dist


def live_parameters(net) -> Dict[str, torch.nn.Parameter]:
    """Parameters that receive a gradient (everything except the dead ones of blocks >= 1)."""
    out = {}
    for name, p in net.named_parameters():
        parts = name.split(".")
        if parts[0] == "BlockList" and int(parts[1]) >= 1 and parts[2] in ("EmbedT", "residual_conv"):
            continue
        out[name] = p
    return out


def bucket_of(name: str) -> str:
    parts = name.split(".")
    return ".".join(parts[:2]) if parts[0] == "BlockList" else "head"


ALIGN = 32      # floats: members of a flat bucket start on 128-byte boundaries


class GradBuckets:
    """Flat per-block gradient storage + overlapped all-reduce.

    At the top of a step the live parameters' ``.grad`` are set to ``None`` (``zero()`` -- or the reference
    loop's own ``optimizer.zero_grad()``, train_DSTAGNN_my.py:154: both work), so autograd simply hands every
    gradient tensor over instead of launching one accumulation kernel per parameter.  When the last gradient of
    a bucket has arrived (post-accumulate hooks; blocks finish their backward in reverse order) the bucket's
    gradients are COLLECTED into its flat buffer by one kernel (``dstag_gather_flat``), the flat buffer is
    all-reduced in place -- overlapping NVLink traffic with the remaining backward kernels -- and each ``p.grad``
    is pointed at its slice of the reduced buffer, so any optimizer that reads ``p.grad`` sees the averaged
    gradients.  ``average=False`` leaves the SUM over ranks in the buckets (``step.FlatAdam`` folds 1 / world
    into its gradient read instead of a separate scaling kernel).
    """

    def __init__(self, net, process_group=None, overlap: bool = True, average: bool = True):
        self.group = process_group
        self.average = average          # False: buckets hold the SUM over ranks (the optimizer scales by 1 / world)
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.overlap = overlap and self.world > 1
        self.buckets: Dict[str, torch.Tensor] = {}
        self._members: Dict[str, List[torch.nn.Parameter]] = {}
        self._views: Dict[str, List[torch.Tensor]] = {}
        self._offsets: Dict[str, List[int]] = {}
        self._pending: Dict[str, int] = {}
        self._done: Dict[str, bool] = {}
        self._works = []
        groups: Dict[str, List] = {}
        for name, p in live_parameters(net).items():
            groups.setdefault(bucket_of(name), []).append((name, p))
        for b, items in groups.items():
            # every member starts on a 128-byte boundary: kernels that read parameters / write gradients through
            # these views (and the flat Adam that mirrors the layout) keep their 16-byte vector paths
            offs, n = [], 0
            for _, p in items:
                offs.append(n)
                n += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
            flat = torch.zeros(n, device=items[0][1].device, dtype=items[0][1].dtype)
            self.buckets[b] = flat
            self._members[b] = [p for _, p in items]
            self._views[b] = [flat[off:off + p.numel()].view_as(p) for (_, p), off in zip(items, offs)]
            self._offsets[b] = offs
            self._pending[b] = len(items)
            self._done[b] = False
        for b, ps in self._members.items():
            for p in ps:
                p.grad = None
                p.register_post_accumulate_grad_hook(self._make_hook(b))
        self.numel = sum(p.numel() for ps in self._members.values() for p in ps)      # without the alignment padding

    def _make_hook(self, b):
        def hook(_p):
            if self._pending[b] <= 0:          # a new step began without zero(): re-arm
                self._pending[b] = len(self._members[b])
                self._done[b] = False
            self._pending[b] -= 1
            if self._pending[b] == 0 and self.overlap:
                self._reduce(b)
        return hook

    def _collect(self, b):
        """Gather the bucket's gradients into its flat buffer (slots of parameters without a gradient are
        zeroed) and re-point every ``p.grad`` at its slice."""
        flat, views = self.buckets[b], self._views[b]
        srcs, offs, sizes = [], [], []
        for p, v, off in zip(self._members[b], views, self._offsets[b]):
            g = p.grad
            if g is None:
                v.zero_()
            elif g.data_ptr() != v.data_ptr():
                if not g.is_contiguous():
                    g = g.contiguous()
                srcs.append(g), offs.append(off), sizes.append(g.numel())
        if srcs:
            if flat.is_cuda:
                import ctypes as C

                from . import _lib
                n = len(srcs)
                _lib.check(_lib.load().dstag_gather_flat((C.c_void_p * n)(*[g.data_ptr() for g in srcs]),
                                                         (C.c_int64 * n)(*offs), (C.c_int64 * n)(*sizes), n,
                                                         flat.data_ptr(), _lib.stream_ptr(flat)), "dstag_gather_flat")
            else:                                  # CPU tensors (gloo tests of the bucket logic)
                for g, off, sz in zip(srcs, offs, sizes):
                    flat[off:off + sz].copy_(g.reshape(-1))
        for p, v in zip(self._members[b], views):
            p.grad = v

    def _reduce(self, b):
        self._collect(b)
        self._done[b] = True
        if self.world > 1:
            flat = self.buckets[b]
            if self.average:
                flat.mul_(1.0 / self.world)
            self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def zero(self):
        """Start of a step: drop the gradients (``.grad = None``: no kernel runs)."""
        for b, ps in self._members.items():
            for p in ps:
                p.grad = None
            self._pending[b] = len(ps)
            self._done[b] = False
        self._works = []

    def finish(self):
        """After ``loss.backward()``: every bucket collected (and averaged across ranks)."""
        for b in self.buckets:
            if not self._done[b]:
                self._reduce(b)
        for w in self._works:
            w.wait()
        self._works = []
        for b in self.buckets:
            self._pending[b] = 0
