"""Mirror of the contrastive losses in `DRT/trainer/losses.py` on fused B200 kernels.

`SimpleContrastiveLoss.forward(x, y, target=None, reduction='mean')` (losses.py:11-17) and
`DistributedContrastiveLoss` (losses.py:20-40) keep their names, arguments and values; the
score matrix, log-sum-exp and NLL run as one CUDA launch (`drt_inbatch_ce_fwd`) and the
backward as `drt_inbatch_ce_bwd`, both fp32 like the reference's `torch.matmul`.
`inbatch_scores_and_loss` serves the loss block of `DRModel.forward`
(DRT/model/biencoder.py:107-119), which also returns the score matrix (biencoder.py:122).

CUDA tensors only: there is no CPU fallback (a CPU tensor raises).
"""
from __future__ import annotations

import torch
from torch import Tensor, nn
from torch import distributed as dist

from . import _lib


def _check_inputs(x: Tensor, y: Tensor) -> None:
    if not (x.is_cuda and y.is_cuda):
        raise RuntimeError("denseretrievaltoolkits_b200 losses need CUDA tensors: there is no CPU fallback")
    if x.dim() != 2 or y.dim() != 2 or x.shape[1] != y.shape[1]:
        raise RuntimeError(f"expected x [B,d] and y [P,d], got {tuple(x.shape)} and {tuple(y.shape)}")


class _InBatchCE(torch.autograd.Function):
    """Returns (per-row loss [B], logits [B,P] or empty)."""

    @staticmethod
    def forward(ctx, x: Tensor, y: Tensor, target, want_logits: bool):
        _check_inputs(x, y)
        lib = _lib.load()
        xc = x.detach().to(torch.float32).contiguous()
        yc = y.detach().to(torch.float32).contiguous()
        B, d = xc.shape
        P = yc.shape[0]
        dev = xc.device
        tgt = None
        if target is not None:
            tgt = target.detach().to(device=dev, dtype=torch.int64).contiguous()
        logits = torch.empty((B, P), dtype=torch.float32, device=dev) if want_logits else None
        lse = torch.empty((B,), dtype=torch.float32, device=dev)
        rows = torch.empty((B,), dtype=torch.float32, device=dev)
        total = torch.empty((1,), dtype=torch.float32, device=dev)
        _lib.check(lib.drt_inbatch_ce_fwd(
            xc.data_ptr(), yc.data_ptr(), B, P, d, tgt.data_ptr() if tgt is not None else None, 1.0,
            logits.data_ptr() if logits is not None else None, lse.data_ptr(), rows.data_ptr(),
            total.data_ptr(), dev.index, _lib.current_stream_ptr(dev.index)), "inbatch_ce_fwd")
        ctx.save_for_backward(xc, yc, lse, tgt if tgt is not None else torch.empty(0, device=dev))
        ctx.has_target = tgt is not None
        ctx.in_dtypes = (x.dtype, y.dtype)
        if logits is None:
            logits = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(logits)
        return rows, logits

    @staticmethod
    def backward(ctx, grad_rows: Tensor, _grad_logits):
        xc, yc, lse, tgt = ctx.saved_tensors
        lib = _lib.load()
        B, d = xc.shape
        P = yc.shape[0]
        dev = xc.device
        g = grad_rows.detach().to(torch.float32).contiguous()
        work = torch.empty((B, P), dtype=torch.float32, device=dev)
        need_x, need_y = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx = torch.empty_like(xc) if need_x else None
        dy = torch.empty_like(yc) if need_y else None
        _lib.check(lib.drt_inbatch_ce_bwd(
            xc.data_ptr(), yc.data_ptr(), B, P, d, tgt.data_ptr() if ctx.has_target else None,
            lse.data_ptr(), g.data_ptr(), work.data_ptr(),
            dx.data_ptr() if dx is not None else None, dy.data_ptr() if dy is not None else None,
            dev.index, _lib.current_stream_ptr(dev.index)), "inbatch_ce_bwd")
        if dx is not None:
            dx = dx.to(ctx.in_dtypes[0])
        if dy is not None:
            dy = dy.to(ctx.in_dtypes[1])
        return dx, dy, None, None


def _reduce(rows: Tensor, reduction: str) -> Tensor:
    if reduction == "mean":
        return rows.mean()
    if reduction == "sum":
        return rows.sum()
    if reduction == "none":
        return rows
    raise ValueError(f"{reduction} is not a valid value for reduction")


def inbatch_scores_and_loss(q_reps: Tensor, p_reps: Tensor, n_passages: int, return_scores: bool = True):
    """Loss block of DRModel.forward (biencoder.py:107-116): scores = q·pᵀ,
    target = arange(B) * train_n_passages, mean cross entropy.  Returns (loss, scores|None)."""
    B = q_reps.shape[0]
    target = torch.arange(B, device=q_reps.device, dtype=torch.long) * int(n_passages)
    rows, logits = _InBatchCE.apply(q_reps, p_reps, target, bool(return_scores))
    return rows.mean(), (logits if return_scores else None)


class SimpleContrastiveLoss(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, x: Tensor, y: Tensor, target: Tensor = None, reduction: str = "mean"):
        rows, _ = _InBatchCE.apply(x, y, target, False)
        return _reduce(rows, reduction)


class DistributedContrastiveLoss(SimpleContrastiveLoss):
    def __init__(self, n_target: int = 0, scale_loss: bool = True):
        assert dist.is_initialized(), "Distributed training has not been properly initialized."
        super().__init__()
        self.word_size = dist.get_world_size()
        self.rank = dist.get_rank()
        self.scale_loss = scale_loss

    def forward(self, x: Tensor, y: Tensor, **kwargs):
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
