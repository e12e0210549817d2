"""CPU oracle for the in-batch-negative score matrix + cross entropy.  TEST INFRASTRUCTURE ONLY
(imported by tests/, __graft_entry__.smoke() and bench.py's baseline legs; never by the product).

Restates, in float64-accumulated numpy so it can arbitrate between implementations:
  * `SimpleContrastiveLoss.forward(x, y, target=None, reduction='mean')`
    (`DRT/trainer/losses.py:11-17`): target_per_qry = y.size(0)//x.size(0);
    target = arange(0, B*tpq, tpq); logits = x @ y.T; F.cross_entropy(logits, target, reduction).
  * the loss block of `DRModel.forward` (`DRT/model/biencoder.py:107-119`):
    scores = q @ p.T; target = arange(B) * train_n_passages; mean CE; `* world_size` when
    training with negatives_x_device.
  * `DistributedContrastiveLoss.forward` (`losses.py:28-34`) = gather rank-major (`:36-40`),
    SimpleContrastiveLoss on the gathered tensors, `* world_size` when scale_loss.

PINNED: the reference's own `DRT.trainer.losses.SimpleContrastiveLoss` is importable in the
authoring container; `tools/make_golden.py` ran it (forward + autograd backward) on seeded
inputs and committed the results under tests/golden/loss_*.npz, and tests/test_oracle.py checks
this restatement against those vectors.
"""
from __future__ import annotations

import numpy as np


def default_target(B: int, P: int) -> np.ndarray:
    tpq = P // B  # losses.py:13
    return np.arange(0, B * tpq, tpq, dtype=np.int64)  # losses.py:14-15


def logits_fp32(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    return (x.astype(np.float32) @ y.astype(np.float32).T).astype(np.float32)  # losses.py:16


def contrastive_loss(x, y, target=None, reduction: str = "mean"):
    """Returns (loss, lse[B], logits[B,P]) with float64 softmax arithmetic over fp32 logits."""
    x = np.asarray(x, np.float32)
    y = np.asarray(y, np.float32)
    B, P = x.shape[0], y.shape[0]
    if target is None:
        target = default_target(B, P)
    logits = x.astype(np.float64) @ y.astype(np.float64).T
    m = logits.max(axis=1, keepdims=True)
    lse = (m + np.log(np.exp(logits - m).sum(axis=1, keepdims=True)))[:, 0]
    rows = lse - logits[np.arange(B), target]
    if reduction == "mean":
        loss = rows.mean()
    elif reduction == "sum":
        loss = rows.sum()
    elif reduction == "none":
        loss = rows
    else:
        raise ValueError(reduction)
    return loss, lse, logits


def contrastive_loss_grads(x, y, target=None, reduction: str = "mean", grad_out=1.0):
    """Analytic gradients (dx, dy) of contrastive_loss wrt x and y."""
    x = np.asarray(x, np.float32).astype(np.float64)
    y = np.asarray(y, np.float32).astype(np.float64)
    B, P = x.shape[0], y.shape[0]
    if target is None:
        target = default_target(B, P)
    logits = x @ y.T
    m = logits.max(axis=1, keepdims=True)
    p = np.exp(logits - m)
    p /= p.sum(axis=1, keepdims=True)
    p[np.arange(B), target] -= 1.0
    if reduction == "mean":
        g = np.full((B, 1), float(grad_out) / B)
    elif reduction == "sum":
        g = np.full((B, 1), float(grad_out))
    else:
        g = np.asarray(grad_out, np.float64).reshape(B, 1)
    dl = p * g
    return dl @ y, dl.T @ x


def gather_rank_major(per_rank: list[np.ndarray]) -> np.ndarray:
    """`dist_gather_tensor` / `gather_tensor` (biencoder.py:243-254, losses.py:36-40):
    concatenation of every rank's tensor along dim 0 in rank order."""
    return np.concatenate(per_rank, axis=0)


def distributed_contrastive_loss(x_ranks, y_ranks, scale_loss: bool = True):
    """`DistributedContrastiveLoss.forward` (losses.py:28-34); identical on every rank."""
    W = len(x_ranks)
    loss, _, _ = contrastive_loss(gather_rank_major(x_ranks), gather_rank_major(y_ranks))
    return loss * W if scale_loss else loss
