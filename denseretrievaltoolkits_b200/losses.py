"""Mirror of the contrastive losses in `DRT/trainer/losses.py` on fused B200 kernels.

`SimpleContrastiveLoss.forward(x, y, target=None, reduction='mean')` (losses.py:11-17) and
`DistributedContrastiveLoss` (losses.py:20-40) keep their names, arguments and values.  At the
reference's per-GPU shapes (B <= 128) the score matrix, log-sum-exp, NLL and the reduction run as
ONE CUDA launch on the tensor cores (`drt_inbatch_ce_fwd`; two launches with more row tiles) and
the backward as two (`drt_inbatch_ce_bwd`), fp32-accurate like the reference's `torch.matmul`
(exact 3-way bf16 split).  The backward never writes a saved tensor.
`inbatch_scores_and_loss` serves the loss block of `DRModel.forward`
(DRT/model/biencoder.py:107-119), which also returns the score matrix (biencoder.py:122).

CUDA tensors only: there is no CPU fallback (a CPU tensor raises).
"""
from __future__ import annotations

import torch
from torch import Tensor, nn
from torch import distributed as dist

from . import _lib
from ._nvtx import rng as _nvtx

_REDUCTIONS = ("mean", "sum", "none")
_KEEP_LOGITS_BYTES = 1 << 30
_NEEDS_WORK = {}


def _needs_work(lib, B: int, P: int, d: int, have_logits: bool) -> bool:
    """Does the backward of this shape want the [B,P] scratch (cached per shape)?"""
    key = (B, P, d, have_logits)
    v = _NEEDS_WORK.get(key)
    if v is None:
        v = _NEEDS_WORK[key] = bool(lib.drt_inbatch_ce_bwd_needs_work(B, P, d, 1 if have_logits else 0))
    return v


def _f32c(t: Tensor) -> Tensor:
    # called inside autograd.Function.forward/backward (grad mode is off there): no detach needed
    if t.dtype is not torch.float32:
        t = t.to(torch.float32)
    return t if t.is_contiguous() else t.contiguous()


class _InBatchCE(torch.autograd.Function):
    """(x [B,d], y [P,d], target|None, reduction, want_logits) -> (loss, logits|None).
    `loss` is a 0-dim tensor for 'mean'/'sum' and [B] for 'none'."""

    @staticmethod
    def forward(ctx, x: Tensor, y: Tensor, target, reduction: str, want_logits: bool):
        if not (x.is_cuda and y.is_cuda):
            raise RuntimeError("denseretrievaltoolkits_b200 losses need CUDA tensors: there is no CPU fallback")
        if x.dim() != 2 or y.dim() != 2 or x.shape[1] != y.shape[1]:
            raise RuntimeError(f"expected x [B,d] and y [P,d], got {tuple(x.shape)} and {tuple(y.shape)}")
        if reduction not in _REDUCTIONS:
            raise ValueError(f"{reduction} is not a valid value for reduction")
        lib = _lib.load()
        xc, yc = _f32c(x), _f32c(y)
        B, d = xc.shape
        P = yc.shape[0]
        dev = xc.device
        tgt = None
        if target is not None:
            tgt = target.detach()
            if tgt.dtype is not torch.int64 or tgt.device != dev or not tgt.is_contiguous():
                tgt = tgt.to(device=dev, dtype=torch.int64).contiguous()
        # the logits are kept for the backward (one elementwise pass instead of a second GEMM)
        # unless the matrix is huge; they are only RETURNED when the caller asked for them
        keep_logits = want_logits or (B * P * 4 <= _KEEP_LOGITS_BYTES and any(ctx.needs_input_grad[:2]))
        logits = torch.empty((B, P), dtype=torch.float32, device=dev) if keep_logits else None
        out = torch.empty((2 * B + 1,), dtype=torch.float32, device=dev)   # per-row loss | total | lse (one allocation)
        lse = out[B + 1:]                                                   # saved for backward
        scale = 1.0 / B if reduction == "mean" else 1.0
        stream = _lib.current_stream_ptr(dev.index)
        base = out.data_ptr()
        with _nvtx("drt.inbatch_ce_fwd"):
            rc = lib.drt_inbatch_ce_fwd(
                xc.data_ptr(), yc.data_ptr(), B, P, d, tgt.data_ptr() if tgt is not None else None, scale,
                logits.data_ptr() if logits is not None else None, base + 4 * (B + 1), base, base + 4 * B,
                dev.index, stream)
        if rc != 0:
            _lib.check(rc, "inbatch_ce_fwd")
        ctx.save_for_backward(xc, yc, lse, tgt if tgt is not None else lse, logits if logits is not None else lse)
        ctx.has_target = tgt is not None
        ctx.has_logits = logits is not None
        ctx.reduction = reduction
        ctx.scale = scale
        ctx.in_dtypes = (x.dtype, y.dtype)
        loss = out[:B] if reduction == "none" else out[B]
        if want_logits:
            ctx.mark_non_differentiable(logits)
            return loss, logits
        return loss, None

    @staticmethod
    def backward(ctx, grad_loss: Tensor, _grad_logits):
        xc, yc, lse, tgt, logits = ctx.saved_tensors
        lib = _lib.load()
        B, d = xc.shape
        P = yc.shape[0]
        dev = xc.device
        g = _f32c(grad_loss)
        per_row = ctx.reduction == "none"
        # the saved logits are never written: a second backward over the same graph
        # (retain_graph=True, two losses sharing the node) sees them unchanged
        work = torch.empty((B, P), dtype=torch.float32, device=dev) if _needs_work(lib, B, P, d, ctx.has_logits) else None
        need_x, need_y = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx = torch.empty_like(xc) if need_x else None
        dy = torch.empty_like(yc) if need_y else None
        _lib.check(lib.drt_inbatch_ce_bwd(
            xc.data_ptr(), yc.data_ptr(), B, P, d, tgt.data_ptr() if ctx.has_target else None,
            lse.data_ptr(), logits.data_ptr() if ctx.has_logits else None, g.data_ptr(), 1 if per_row else 0, ctx.scale,
            work.data_ptr() if work is not None else None,
            dx.data_ptr() if dx is not None else None, dy.data_ptr() if dy is not None else None,
            dev.index, _lib.current_stream_ptr(dev.index)), "inbatch_ce_bwd")
        if dx is not None and ctx.in_dtypes[0] is not torch.float32:
            dx = dx.to(ctx.in_dtypes[0])
        if dy is not None and ctx.in_dtypes[1] is not torch.float32:
            dy = dy.to(ctx.in_dtypes[1])
        return dx, dy, None, None, None


def inbatch_scores_and_loss(q_reps: Tensor, p_reps: Tensor, n_passages: int, return_scores: bool = True):
    """Loss block of DRModel.forward (biencoder.py:107-116): scores = q·pᵀ,
    target = arange(B) * train_n_passages, mean cross entropy.  Returns (loss, scores|None)."""
    B = q_reps.shape[0]
    target = torch.arange(0, B * int(n_passages), int(n_passages), device=q_reps.device, dtype=torch.long)
    loss, logits = _InBatchCE.apply(q_reps, p_reps, target, "mean", bool(return_scores))
    return loss, (logits if return_scores else None)


class SimpleContrastiveLoss(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, x: Tensor, y: Tensor, target: Tensor = None, reduction: str = "mean"):
        return _InBatchCE.apply(x, y, target, reduction, False)[0]


class _ShardedInBatchCE(torch.autograd.Function):
    """Cross-device in-batch loss without the W-fold redundant work of the reference.

    The reference all-gathers queries AND passages and every rank evaluates the whole
    [B·W, P·W] loss (losses.py:28-34, biencoder.py:103-119).  Here only the passages are
    gathered; rank r evaluates its own B rows against all P·W passages (targets shifted by the
    rank's row offset), the per-rank loss sums are all-reduced, and in the backward the passage
    gradient contributions dlogitsᵀ·x_local of all ranks are reduce-scattered.  Values and
    gradients equal the reference's (same sums, 1/W of the FLOPs per rank)."""

    @staticmethod
    def forward(ctx, x: Tensor, y: Tensor, group, rank: int, world: int, scale_loss: bool):
        lib = _lib.load()
        xc, yc = _f32c(x), _f32c(y)
        B, d = xc.shape
        P = yc.shape[0]
        dev = xc.device
        y_all = torch.empty((world * P, d), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(y_all, yc, group=group)
        tpq = (world * P) // (world * B)                                   # losses.py:13 on the gathered shapes
        target = (torch.arange(B, device=dev, dtype=torch.int64) + rank * B) * tpq
        lse = torch.empty((B,), dtype=torch.float32, device=dev)
        out = torch.empty((B + 1,), dtype=torch.float32, device=dev)
        base = out.data_ptr()
        # kept for the backward (dlogits is then formed from them instead of a second contraction)
        keep = B * world * P * 4 <= _KEEP_LOGITS_BYTES and any(ctx.needs_input_grad[:2])
        logits = torch.empty((B, world * P), dtype=torch.float32, device=dev) if keep else None
        _lib.check(lib.drt_inbatch_ce_fwd(xc.data_ptr(), y_all.data_ptr(), B, world * P, d, target.data_ptr(), 1.0,
                                          logits.data_ptr() if keep else None,
                                          lse.data_ptr(), base, base + 4 * B, dev.index,
                                          _lib.current_stream_ptr(dev.index)), "inbatch_ce_fwd")
        total = out[B:].clone()
        dist.all_reduce(total, group=group)
        coef = (float(world) if scale_loss else 1.0) / float(world * B)    # mean over all rows (x world_size)
        ctx.save_for_backward(xc, y_all, lse, target, logits if keep else lse)
        ctx.has_logits = keep
        ctx.meta = (group, rank, world, P, coef, x.dtype, y.dtype)
        return (total * coef).reshape(())

    @staticmethod
    def backward(ctx, grad_loss: Tensor):
        xc, y_all, lse, target, logits = ctx.saved_tensors
        group, rank, world, P, coef, xdt, ydt = ctx.meta
        lib = _lib.load()
        B, d = xc.shape
        dev = xc.device
        g = _f32c(grad_loss).reshape(1)
        work = (torch.empty((B, world * P), dtype=torch.float32, device=dev)
                if _needs_work(lib, B, world * P, d, ctx.has_logits) else None)
        dx = torch.empty_like(xc)
        dy_all = torch.empty_like(y_all)
        _lib.check(lib.drt_inbatch_ce_bwd(xc.data_ptr(), y_all.data_ptr(), B, world * P, d, target.data_ptr(), lse.data_ptr(),
                                          logits.data_ptr() if ctx.has_logits else None, g.data_ptr(), 0, coef,
                                          work.data_ptr() if work is not None else None, dx.data_ptr(), dy_all.data_ptr(),
                                          dev.index, _lib.current_stream_ptr(dev.index)), "inbatch_ce_bwd")
        dy = torch.empty((P, d), dtype=torch.float32, device=dev)
        dist.reduce_scatter_tensor(dy, dy_all, group=group)
        return dx.to(xdt), dy.to(ydt), None, None, None, None


class DistributedContrastiveLoss(SimpleContrastiveLoss):
    def __init__(self, n_target: int = 0, scale_loss: bool = True):
        assert dist.is_initialized(), "Distributed training has not been properly initialized."
        super().__init__()
        self.word_size = dist.get_world_size()
        self.rank = dist.get_rank()
        self.scale_loss = scale_loss

    def forward(self, x: Tensor, y: Tensor, **kwargs):
        if not kwargs and x.is_cuda and dist.get_backend() == "nccl":
            # default target / mean reduction: the sharded evaluation (1/W of the work per rank)
            return _ShardedInBatchCE.apply(x, y, None, self.rank, self.word_size, self.scale_loss)
        dist_x = self.gather_tensor(x)
        dist_y = self.gather_tensor(y)
        loss = super().forward(dist_x, dist_y, **kwargs)
        if self.scale_loss:
            loss = loss * self.word_size
        return loss

    def gather_tensor(self, t: Tensor) -> Tensor:
        # rank-major concatenation; the local slot keeps the autograd-carrying tensor
        # (losses.py:36-40, biencoder.py:243-254)
        return gather_rank_major(t, self.rank, self.word_size)


def gather_rank_major(t: Tensor, rank: int, world_size: int, group=None) -> Tensor:
    t = t.contiguous()
    parts = [torch.empty_like(t) for _ in range(world_size)]
    dist.all_gather(parts, t.detach(), group=group)
    parts[rank] = t
    return torch.cat(parts, dim=0)


def get_loss_function(training_args):
    if training_args.loss_fn == "SimpleContrastiveLoss":
        return DistributedContrastiveLoss()
