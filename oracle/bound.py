"""TEST INFRASTRUCTURE — numpy restatement of the exactness certificate's error bound.

Not product code: only tests/ may import this.  It restates, in float64, the quantities the CUDA
path keeps per row / per query (`ingest_rows_kernel`, `prep_queries_kernel` in
denseretrievaltoolkits_b200/csrc/select_kernels.cuh) and the inequality the certificate rests on,

    s(q, j)  <=  ub(q, j) = s~(q, j) + A_q r_j + B_q Dx_j + C_q Dt_j,

where s is the exact inner product (what faiss.IndexFlatIP computes in fp32,
DRT/evaluator/index.py:32), s~ the inner product of the operands' 16-bit images (bf16 by
default, IEEE fp16 with DRT_B200_FIRST_PASS=f16), r_j = |d_j - image(d_j)|, Dx / Dt the norms of the head / tail dims of d_j, A_q = |q~|, B_q / C_q =
|q - q~| over the head / tail dims, q~ the query's 16-bit image (the CUDA path adds its accumulation-rounding terms on
top).  The reference has no such notion (faiss scores every row in fp32); this file exists so
the bound's algebra is checked on the CPU independently of the kernels.
"""
from __future__ import annotations

import numpy as np


def bf16_round(a: np.ndarray) -> np.ndarray:
    """fp32 -> nearest-even bf16 -> fp32 (finite inputs)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return u.view(np.float32)


def f16_round(a: np.ndarray) -> np.ndarray:
    """fp32 -> nearest-even IEEE fp16 -> fp32: the query image of the fp16 first pass
    (`prep_queries_kernel`, f16 = 1; the power-of-two query scaling changes nothing here)."""
    with np.errstate(over="ignore"):
        return np.ascontiguousarray(a, dtype=np.float32).astype(np.float16).astype(np.float32)


def f16_sat_round(a: np.ndarray) -> np.ndarray:
    """Row image in fp16 mode: saturate at +-65504, then round (`ingest_rows_kernel`, sat16)."""
    return f16_round(np.clip(np.asarray(a, dtype=np.float32), -65504.0, 65504.0))


QUERY_ROUND = {"f16": f16_round, "bf16": bf16_round}
ROW_ROUND = {"f16": f16_sat_round, "bf16": bf16_round}


def row_bounds(x: np.ndarray, split: int | None = None, fmt: str = "bf16") -> np.ndarray:
    """[n,3] float64: (r, Dx, Dt) per row."""
    x64 = np.asarray(x, dtype=np.float32).astype(np.float64)
    split = x64.shape[1] if split is None else split
    r = np.linalg.norm(x64 - ROW_ROUND[fmt](x).astype(np.float64), axis=1)
    return np.stack([r, np.linalg.norm(x64[:, :split], axis=1), np.linalg.norm(x64[:, split:], axis=1)], axis=1)


def query_bounds(q: np.ndarray, split: int | None = None, qfmt: str = "bf16") -> np.ndarray:
    """[nq,3] float64: (A, B, C) per query, exact-arithmetic part only."""
    q64 = np.asarray(q, dtype=np.float32).astype(np.float64)
    split = q64.shape[1] if split is None else split
    qt = QUERY_ROUND[qfmt](q).astype(np.float64)
    e = q64 - qt
    return np.stack([np.linalg.norm(qt, axis=1), np.linalg.norm(e[:, :split], axis=1), np.linalg.norm(e[:, split:], axis=1)], axis=1)


def first_pass_scores(x: np.ndarray, q: np.ndarray, qfmt: str = "bf16") -> np.ndarray:
    """Inner products of the 16-bit images of the operands (both bf16 — the default — or both
    fp16: tcgen05 kind::f16 wants one format for A and B), accumulated exactly (float64)."""
    return QUERY_ROUND[qfmt](q).astype(np.float64) @ ROW_ROUND[qfmt](x).astype(np.float64).T


def upper_bounds(x: np.ndarray, q: np.ndarray, split: int | None = None, qfmt: str = "bf16") -> np.ndarray:
    """[nq,n] float64 ub(q,j)."""
    return first_pass_scores(x, q, qfmt) + query_bounds(q, split, qfmt) @ row_bounds(x, split, qfmt).T


def certified_topk(x: np.ndarray, q: np.ndarray, k: int, kprime: int, split: int | None = None, qfmt: str = "bf16"):
    """The selection rule of the CUDA path in float64: keep the kprime rows with the largest ub,
    rescore them exactly, and certify a query when its exact k-th score lies strictly above
    thr = the kprime-th largest ub.  Returns (ids [nq,k], certified [nq] bool)."""
    ub = upper_bounds(x, q, split, qfmt)
    exact = np.asarray(q, np.float64) @ np.asarray(x, np.float64).T
    n = x.shape[0]
    kp = min(kprime, n)
    ids = np.empty((q.shape[0], min(k, n)), dtype=np.int64)
    ok = np.empty((q.shape[0],), dtype=bool)
    for i in range(q.shape[0]):
        cand = np.argsort(-ub[i], kind="stable")[:kp]
        order = cand[np.lexsort((cand, -exact[i, cand]))]
        ids[i] = order[:ids.shape[1]]
        thr = ub[i, cand[-1]] if kp < n else -np.inf
        ok[i] = exact[i, ids[i, -1]] > thr
    return ids, ok


def adverse_row(q0: np.ndarray, mag: float = 256.0, fmt: str = "bf16") -> np.ndarray:
    """A row whose 16-bit image scores about -extra/2 against q0 while its exact score is about
    +extra/2: every component is a representable +-mag plus just under half an ulp in the
    direction of q0, so rounding removes extra = sum_i |q0_i| * 0.49 ulp from the first-pass score.
    `mag` must be a power of two; ulp = mag/128 for bf16 (8-bit significand), mag/1024 for fp16 (11)."""
    ulp = mag / (128.0 if fmt == "bf16" else 1024.0)   # spacing of the format in [mag, 2 mag)
    target = -0.5 * float(np.abs(q0).sum()) * 0.49 * ulp
    m = np.empty_like(q0)
    acc = 0.0
    for i in np.argsort(-np.abs(q0)):               # greedy signs, large steps first: q0 . m -> target
        v = float(q0[i])
        s = -1.0 if (acc > target) == (v > 0) else 1.0
        m[i] = s * mag
        acc += v * m[i]
    return (m + np.sign(q0) * np.float32(0.49 * ulp)).astype(np.float32)
