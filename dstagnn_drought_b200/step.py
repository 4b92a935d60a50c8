"""The training step either side of the block stack, B200 side:

  * ``smooth_l1_loss``  -- ``nn.SmoothL1Loss()`` of train_DSTAGNN_my.py:132,156 as one streaming kernel each
                           way (reproducible reduction, gradient scaled by a device scalar);
  * ``FlatAdam``        -- ``optim.Adam`` of train_DSTAGNN_my.py:133,158 as ONE kernel per gradient bucket: the
                           live parameters, their gradients and the two moments live in flat fp32 buffers
                           (``p.data`` / ``p.grad`` are views), so the update streams four arrays once;
  * ``StepCounters``    -- device-resident step counters (dropout key, Adam step) so that nothing in the step
                           depends on a host integer;
  * ``TrainStep``       -- zero-grad -> forward -> loss -> backward -> (gradient all-reduce) -> Adam as one
                           callable, optionally captured in a CUDA graph: ~650 launches are replayed with a
                           single ``cudaGraphLaunch``, which is what the N <= 358 graphs -- whose step is host
                           bound otherwise -- need.

Dead parameters (EmbedT / residual_conv of blocks >= 1) are left out of every flat buffer: they
never receive a gradient in the reference either, so Adam never touches them there.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import _lib, ops
from .contraction import _check, arm_zero_arena, disarm_zero_arena
from .parallel import GradBuckets


def _st(t):
    return _lib.stream_ptr(t)


def shutdown_process_group(graphs_captured: bool):
    """End of a multi-rank run.  A process that captured NCCL collectives in CUDA graphs must not go through
    ``destroy_process_group`` / interpreter teardown (observed on B200, torch 2.11 + NCCL 2.28: the communicator
    abort waits forever on work that only exists inside the graphs): synchronise, barrier, flush and leave."""
    import sys

    import torch.distributed as dist
    if not dist.is_initialized():
        return
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    if graphs_captured:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    dist.destroy_process_group()


def multi_gpu_graph_env():
    """Call BEFORE ``init_process_group`` when multi-rank training steps are to be captured in CUDA graphs."""
    if os.environ.get("DSTAG_GRAPH_MULTI", "1") != "0":
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")


# ----------------------------------------------------------------------------------------------------
# loss
# ----------------------------------------------------------------------------------------------------
class _SmoothL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, workspace):
        lib = _lib.load()
        _check(pred, "pred"), _check(target, "target")
        if pred.shape != target.shape:
            raise RuntimeError(f"smooth_l1_loss: shapes differ: {tuple(pred.shape)} vs {tuple(target.shape)}")
        pred, target = pred.contiguous(), target.contiguous()
        loss = torch.empty((), device=pred.device, dtype=torch.float32)
        _lib.check(lib.dstag_smooth_l1_fwd(pred.data_ptr(), target.data_ptr(), loss.data_ptr(), workspace.data_ptr(),
                                           pred.numel(), _st(pred)), "dstag_smooth_l1_fwd")
        ctx.save_for_backward(pred, target)
        return loss

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        pred, target = ctx.saved_tensors
        dpred = torch.empty_like(pred)
        gout = gout.contiguous()
        _lib.check(lib.dstag_smooth_l1_bwd(pred.data_ptr(), target.data_ptr(), gout.data_ptr(), dpred.data_ptr(),
                                           pred.numel(), _st(pred)), "dstag_smooth_l1_bwd")
        return dpred, None, None


_loss_ws: Dict[int, torch.Tensor] = {}


def smooth_l1_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """mean SmoothL1 (beta = 1) of ``pred - target``: ``nn.SmoothL1Loss()(pred, target)``."""
    ws = _loss_ws.get(pred.device.index)
    if ws is None:
        ws = _loss_ws[pred.device.index] = torch.zeros(1025, device=pred.device, dtype=torch.float32)
    return _SmoothL1.apply(pred, target, ws)


# ----------------------------------------------------------------------------------------------------
# counters
# ----------------------------------------------------------------------------------------------------
class StepCounters:
    """int64[2] on the device: [0] = steps finished (folded into every dropout key), [1] = Adam step count."""

    def __init__(self, device):
        self.t = torch.zeros(2, device=device, dtype=torch.int64)
        self.t[1] = 1                                        # Adam's t is 1-based
        ops.set_seed_counter(device, self.t)

    @property
    def dropout_ptr(self):
        return self.t.data_ptr()

    @property
    def adam_ptr(self):
        return self.t.data_ptr() + 8

    def tick(self):
        _lib.check(_lib.load().dstag_tick(self.t.data_ptr(), 2, _st(self.t)), "dstag_tick")

    def close(self):
        ops.set_seed_counter(self.t.device, None)


# ----------------------------------------------------------------------------------------------------
# optimizer
# ----------------------------------------------------------------------------------------------------
class FlatAdam:
    """Adam (betas 0.9 / 0.999, eps 1e-8: the torch defaults the reference uses) over the gradient buckets of
    ``GradBuckets``: per bucket one flat parameter buffer (``p.data`` become views of it -- ``state_dict`` and
    checkpoints are unaffected), two moment buffers, one kernel per step."""

    def __init__(self, net, buckets: GradBuckets, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
                 counters: Optional[StepCounters] = None):
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        self.buckets = buckets
        self.counters = counters
        self.step_host = 0
        self.flat_p: Dict[str, torch.Tensor] = {}
        self.m: Dict[str, torch.Tensor] = {}
        self.v: Dict[str, torch.Tensor] = {}
        for b, members in buckets._members.items():
            flat = torch.zeros_like(buckets.buckets[b])     # same (128-byte aligned) member offsets as the gradients
            with torch.no_grad():
                for p, off in zip(members, buckets._offsets[b]):
                    view = flat[off:off + p.numel()].view_as(p)
                    view.copy_(p.data)
                    p.data = view                          # the parameter now lives inside the flat buffer
            self.flat_p[b] = flat
            self.m[b] = torch.zeros_like(flat)
            self.v[b] = torch.zeros_like(flat)

    def step(self, grad_scale: float = 1.0):
        lib = _lib.load()
        self.step_host += 1
        dev_ptr = self.counters.adam_ptr if self.counters is not None else None
        for b, flat in self.flat_p.items():
            g = self.buckets.buckets[b]
            _lib.check(lib.dstag_adam_step(flat.data_ptr(), g.data_ptr(), self.m[b].data_ptr(), self.v[b].data_ptr(),
                                           flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps,
                                           float(grad_scale), dev_ptr, self.step_host, _st(flat)), "dstag_adam_step")

    def state_dict(self):
        return {"step": self.step_host, "m": {b: t.clone() for b, t in self.m.items()},
                "v": {b: t.clone() for b, t in self.v.items()}}


# ----------------------------------------------------------------------------------------------------
# the whole step
# ----------------------------------------------------------------------------------------------------
class TrainStep:
    """``step(x, y) -> loss`` (a 0-d device tensor): zero-grad, forward, SmoothL1, backward, gradient
    averaging across ranks, Adam -- train_DSTAGNN_my.py:154-158 -- with every kernel of this library.

    ``graph=True`` captures the step in a CUDA graph after ``warmup`` eager steps.  With more than one rank the
    bucketed NCCL all-reduces (launched from the autograd hooks, on NCCL's own stream) are captured with it --
    NCCL >= 2.9.6 supports stream capture; the process group must be created with
    ``TORCH_NCCL_ASYNC_ERROR_HANDLING=0`` (the watchdog's event queries are illegal during a capture), which
    ``multi_gpu_graph_env()`` sets; ``DSTAG_GRAPH_MULTI=0`` keeps multi-rank steps eager.  A captured step reads its
    batch from static buffers: ``step`` copies the caller's tensors into them (device-to-device, or
    host-to-device from pinned memory) before the replay."""

    def __init__(self, net, lr: float, graph: bool = False, warmup: int = 3, process_group=None):
        import torch.distributed as dist
        self.net = net
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        # the buckets deliver the SUM over ranks; Adam folds the 1 / world into its gradient read
        self.buckets = GradBuckets(net, process_group=process_group, overlap=True, average=False)
        dev = next(iter(self.buckets.buckets.values())).device
        self.counters = StepCounters(dev)
        self.opt = FlatAdam(net, self.buckets, lr, counters=self.counters)
        # zero-initialised accumulators of the backward (bias / LayerNorm / table / mask-bias gradients) come out
        # of one arena that is cleared once per step
        self.arena = arm_zero_arena(dev, int(1.25 * self.buckets.numel) + (1 << 20))
        self.want_graph = graph and (self.world == 1 or os.environ.get("DSTAG_GRAPH_MULTI", "1") != "0")
        self.warmup = warmup
        self._graph = None
        self._seen = 0
        self._static = None
        self.kernels_per_replay = 0      # kernels of this library inside the captured graph
        self.replays = 0

    def close(self):
        self.counters.close()
        disarm_zero_arena(self.counters.t.device)

    def _eager(self, x, y):
        self.arena.reset()
        self.buckets.zero()
        loss = smooth_l1_loss(self.net(x), y)
        loss.backward()
        self.buckets.finish()
        self.opt.step(grad_scale=1.0 / self.world)
        self.counters.tick()
        return loss.detach()

    def _capture(self, x, y):
        self._static = (torch.empty_like(x), torch.empty_like(y))
        self._static[0].copy_(x), self._static[1].copy_(y)
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        with torch.cuda.graph(g):
            self._loss = self._eager(*self._static)
        self.kernels_per_replay = _lib.launch_count() - n0
        self._graph = g

    def step(self, x, y):
        if self._graph is not None:
            if x.shape != self._static[0].shape:             # ragged last batch of an epoch: run it eagerly
                return self._eager(x, y)
            self._static[0].copy_(x, non_blocking=True)
            self._static[1].copy_(y, non_blocking=True)
            self._graph.replay()
            self.replays += 1
            return self._loss
        self._seen += 1
        if self.want_graph and self._seen > self.warmup:
            self._capture(x, y)
            self._graph.replay()
            self.replays += 1
            return self._loss
        return self._eager(x, y)

    __call__ = step
