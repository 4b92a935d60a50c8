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


class GradBuckets:
    """Flat per-block gradient storage + overlapped averaging all-reduce."""

    def __init__(self, net, process_group=None, overlap: bool = True):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.overlap = overlap and self.world > 1
        self.buckets: Dict[str, torch.Tensor] = {}
        self._members: Dict[str, List[torch.nn.Parameter]] = {}
        self._pending: Dict[str, int] = {}
        self._works = []
        groups: Dict[str, List] = {}
        for name, p in live_parameters(net).items():
            groups.setdefault(bucket_of(name), []).append((name, p))
        for b, items in groups.items():
            n = sum(p.numel() for _, p in items)
            flat = torch.zeros(n, device=items[0][1].device, dtype=items[0][1].dtype)
            off = 0
            for _, p in items:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self.buckets[b] = flat
            self._members[b] = [p for _, p in items]
        if self.overlap:
            for b, ps in self._members.items():
                for p in ps:
                    p.register_post_accumulate_grad_hook(self._make_hook(b))
        self.numel = sum(f.numel() for f in self.buckets.values())

    def _make_hook(self, b):
        def hook(_p):
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        flat = self.buckets[b]
        flat.mul_(1.0 / self.world)
        self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def zero(self):
        """Start of a step: clear the flat buffers (``.grad`` views stay attached)."""
        for b, flat in self.buckets.items():
            flat.zero_()
            self._pending[b] = len(self._members[b])
        self._works = []

    def finish(self):
        """After ``loss.backward()``: make sure every bucket is averaged across ranks."""
        if self.world == 1:
            return
        if not self.overlap:
            for b in self.buckets:
                self._launch(b)
        else:
            for b, n in self._pending.items():
                if n > 0:          # a parameter got no gradient this step; reduce what there is
                    self._launch(b)
        for w in self._works:
            w.wait()
        self._works = []
