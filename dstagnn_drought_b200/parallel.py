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
    """Flat per-block gradient storage + overlapped averaging all-reduce.

    ``p.grad`` of every live parameter is a view into its bucket's flat buffer.  The reference loop
    calls ``optimizer.zero_grad()`` (train_DSTAGNN_my.py:154), whose default ``set_to_none=True``
    DETACHES those views; the buckets therefore check the attachment whenever they are about to reduce
    (and in ``zero()``): a gradient that autograd allocated afresh is copied into its slot and
    re-attached, a gradient that is still ``None`` leaves a zero slot.  Either calling convention --
    ``buckets.zero()`` or ``optimizer.zero_grad()`` at the top of the step -- gives the same result.
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
            views = []
            for (_, p), off in zip(items, offs):
                v = flat[off:off + p.numel()].view_as(p)
                p.grad = v
                views.append(v)
            self.buckets[b] = flat
            self._members[b] = [p for _, p in items]
            self._views[b] = views
            self._offsets[b] = offs
            self._pending[b] = len(items)
        if self.overlap:
            for b, ps in self._members.items():
                for p in ps:
                    p.register_post_accumulate_grad_hook(self._make_hook(b))
        self.numel = sum(p.numel() for ps in self._members.values() for p in ps)      # without the alignment padding

    def _make_hook(self, b):
        def hook(_p):
            if self._pending[b] <= 0:          # a new step began without zero(): re-arm
                self._pending[b] = len(self._members[b])
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)
        return hook

    def _attach(self, b, copy: bool):
        """Make every ``p.grad`` of bucket b a view of the flat buffer again.  ``copy``: gradients found in
        foreign storage are copied into their slot first (they were produced by this step's backward);
        a gradient that is ``None`` gets a zeroed slot."""
        for p, v in zip(self._members[b], self._views[b]):
            g = p.grad
            if g is None:
                if copy:
                    v.zero_()
            elif g.data_ptr() != v.data_ptr() or g.shape != v.shape or not g.is_contiguous():
                if copy:
                    v.copy_(g)
            else:
                continue
            p.grad = v

    def _launch(self, b):
        self._attach(b, copy=True)
        flat = self.buckets[b]
        if self.average:
            flat.mul_(1.0 / self.world)
        self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def zero(self):
        """Start of a step: clear the flat buffers and (re-)attach the ``.grad`` views."""
        for b, flat in self.buckets.items():
            flat.zero_()
            self._attach(b, copy=False)
            self._pending[b] = len(self._members[b])
        self._works = []

    def finish(self):
        """After ``loss.backward()``: make sure every bucket is averaged across ranks (and, also at
        world size 1, that every gradient lives in its bucket again)."""
        if self.world == 1:
            for b in self.buckets:
                self._attach(b, copy=True)
            return
        if not self.overlap:
            for b in self.buckets:
                self._launch(b)
        else:
            for b, n in self._pending.items():
                if n > 0:          # a parameter got no gradient this step; reduce what there is
                    self._launch(b)
                    self._pending[b] = 0
        for w in self._works:
            w.wait()
        self._works = []
