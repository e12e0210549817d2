"""CPU oracle for the exact-MIPS search path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this module; the product package (`denseretrievaltoolkits_b200`) never does and fails
loudly when its CUDA library is missing.

What it restates
----------------
The reference delegates all search arithmetic to `faiss.IndexFlatIP`
(`DRT/evaluator/index.py:16-33`): `index.add(x)` appends fp32 rows with ids = insertion order
(`index.py:28`), `index.search(q, k)` returns `(D float32[Q,k], I int64[Q,k])`
(`index.py:32`), and the wrapper re-orders each id row by `np.argsort(-scores)` and returns ids
only (`index.py:33`).

faiss (PyPI `faiss-cpu`, **version unpinned** by the reference: it has no requirements file) is
not installed and not installable here, so this file restates faiss' published `IndexFlatIP`
algorithm: blocked fp32 `sgemm` of the query block against corpus blocks followed by a per-query
min-heap / reservoir k-selection whose admission test is `threshold < score` with the threshold
initialised to `-FLT_MAX` and ids to `-1` (so NaN, -inf and -FLT_MAX scores never enter, and
`k > ntotal` leaves `(-FLT_MAX, -1)` padding).  Tie order inside faiss is implementation
defined; the canonical order used here and by the CUDA path is (score desc, id asc).

PARITY PIN STATUS: the reference ships no tests, fixtures or golden vectors for this path
(SURVEY.md §4/§8c), and faiss itself cannot run in this environment, so against the faiss
*binary* this oracle is **parity unpinned**.  It is pinned (tests/test_oracle.py) to
 (a) the mathematical definition, via a float64 brute-force cross-check,
 (b) an independent plain-C restatement (oracle/flat_ip_c.c, scalar fp32 + heap), and
 (c) the reference's own wrapper code `DRT/evaluator/index.py:16-44` executed over this oracle
     through a faiss-shaped stub (tools/make_golden.py generated tests/golden/search_*.npz), and
 (d) third-party exact search: torch blocked mm + topk (inner product) and scikit-learn's
     brute-force NearestNeighbors (squared L2, for the IndexFlatL2 restatement below).
"""
from __future__ import annotations

import numpy as np

FLT_LOWEST = np.float32(-3.4028234663852886e38)


def _select_block(scores: np.ndarray, ids: np.ndarray, k: int):
    """Per-row top-k of `scores` [Q, M] (ids [M] or [Q, M]) in canonical order
    (score desc, id asc), admitting only scores > -FLT_MAX (faiss: `threshold < score`).
    Returns (D [Q,k], I [Q,k]) padded with (-FLT_MAX, -1)."""
    Q, M = scores.shape
    if ids.ndim == 1:
        ids = np.broadcast_to(ids[None, :], (Q, M))
    D = np.full((Q, k), FLT_LOWEST, dtype=np.float32)
    I = np.full((Q, k), -1, dtype=np.int64)
    if M == 0:
        return D, I
    s = np.where(scores > FLT_LOWEST, scores, -np.inf).astype(np.float32)  # NaN -> -inf too
    kk = min(k, M)
    if kk < M:
        # k-th largest value per row; everything strictly above it is in, ties at it are
        # resolved by ascending id below.
        kth = np.partition(s, M - kk, axis=1)[:, M - kk]
    else:
        kth = np.full((Q,), -np.inf, dtype=np.float32)
    for r in range(Q):
        row = s[r]
        cand = np.nonzero((row >= kth[r]) & (row > -np.inf))[0]
        if cand.size == 0:
            continue
        order = np.lexsort((ids[r, cand], -row[cand].astype(np.float64)))
        keep = cand[order[:kk]]
        D[r, : keep.size] = scores[r, keep]
        I[r, : keep.size] = ids[r, keep]
    return D, I


class IndexFlatIP:
    """faiss.IndexFlatIP restatement (the subset the reference touches: index.py:19-32,
    trainer.py:235,245,257)."""

    def __init__(self, d: int, block_rows: int = 65536):
        self.d = int(d)
        self.is_trained = True
        self.verbose = False
        self._blocks: list[np.ndarray] = []
        self.ntotal = 0
        self._block_rows = block_rows

    def add(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise RuntimeError(f"add: expected [n,{self.d}] float32, got {x.shape}")
        self._blocks.append(x.copy())
        self.ntotal += x.shape[0]

    def reset(self) -> None:
        self._blocks, self.ntotal = [], 0

    def reconstruct_n(self, i0: int, n: int) -> np.ndarray:
        allx = np.concatenate(self._blocks) if self._blocks else np.zeros((0, self.d), np.float32)
        return allx[i0 : i0 + n].copy()

    def search(self, q: np.ndarray, k: int):
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.d:
            raise RuntimeError(f"search: expected [nq,{self.d}] float32, got {q.shape}")
        Q = q.shape[0]
        D = np.full((Q, k), FLT_LOWEST, dtype=np.float32)
        I = np.full((Q, k), -1, dtype=np.int64)
        base = 0
        for xb in self._blocks:
            for r0 in range(0, xb.shape[0], self._block_rows):
                blk = xb[r0 : r0 + self._block_rows]
                s = q @ blk.T  # fp32 sgemm, fp32 accumulate (faiss: cblas_sgemm per block)
                ids = np.arange(base + r0, base + r0 + blk.shape[0], dtype=np.int64)
                bd, bi = _select_block(s, ids, k)
                # merge the running result with this block's (both canonical, disjoint ids)
                md = np.concatenate([D, bd], axis=1)
                mi = np.concatenate([I, bi], axis=1)
                # padding ids -1 must sort last among equal (-FLT_MAX) scores
                mkey = np.where(mi < 0, np.iinfo(np.int64).max, mi)
                D, I = _merge_rows(md, mi, mkey, k)
            base += xb.shape[0]
        return D, I


def _merge_rows(md, mi, mkey, k):
    Q = md.shape[0]
    D = np.empty((Q, k), np.float32)
    I = np.empty((Q, k), np.int64)
    for r in range(Q):
        order = np.lexsort((mkey[r], -md[r].astype(np.float64)))[:k]
        D[r] = md[r, order]
        I[r] = mi[r, order]
    return D, I


def flat_ip_search(corpus: np.ndarray, q: np.ndarray, k: int):
    """One-shot helper: exact top-k inner product, canonical order."""
    idx = IndexFlatIP(corpus.shape[1])
    idx.add(corpus)
    return idx.search(q, k)


def flat_ip_search_stream(blocks, q: np.ndarray, k: int):
    """`IndexFlatIP.search` over a corpus that is handed over block by block —
    `blocks` yields fp32 arrays [m_i, d] in id order — so a full-size corpus (8.8M x 768 = 27 GB)
    never has to sit in host memory at once.  Same arithmetic and order as `flat_ip_search`."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    Q = q.shape[0]
    D = np.full((Q, k), FLT_LOWEST, dtype=np.float32)
    I = np.full((Q, k), -1, dtype=np.int64)
    base = 0
    for xb in blocks:
        xb = np.ascontiguousarray(xb, dtype=np.float32)
        for r0 in range(0, xb.shape[0], 65536):
            blk = xb[r0:r0 + 65536]
            ids = np.arange(base + r0, base + r0 + blk.shape[0], dtype=np.int64)
            bd, bi = _select_block(q @ blk.T, ids, k)
            md, mi = np.concatenate([D, bd], axis=1), np.concatenate([I, bi], axis=1)
            D, I = _merge_rows(md, mi, np.where(mi < 0, np.iinfo(np.int64).max, mi), k)
        base += xb.shape[0]
    return D, I


def flat_ip_search_f64(corpus: np.ndarray, q: np.ndarray, k: int):
    """float64 brute force used to pin the fp32 oracle to the mathematical definition."""
    s = q.astype(np.float64) @ corpus.astype(np.float64).T
    Q, N = s.shape
    kk = min(k, N)
    D = np.full((Q, k), float(FLT_LOWEST), dtype=np.float64)
    I = np.full((Q, k), -1, dtype=np.int64)
    ids = np.arange(N)
    for r in range(Q):
        order = np.lexsort((ids, -s[r]))[:kk]
        D[r, :kk] = s[r, order]
        I[r, :kk] = order
    return D, I


def wrapper_search_ids(index: IndexFlatIP, q_reps: np.ndarray, k: int = 1000) -> np.ndarray:
    """`BaseFaissIPRetriever.search` (DRT/evaluator/index.py:31-33): ids only, each row
    re-ordered by argsort(-scores) (identity on an already-sorted row; a stable sort is used so
    ties keep ascending-id order)."""
    scores, indices = index.search(q_reps, k)
    return np.array([ind[x] for ind, x in zip(indices, np.argsort(-scores, kind="stable"))])


def torch_flat_ip_search(corpus_t, q_t, k: int, block_rows: int = 262144):
    """The CPU baseline the north star names when faiss is absent: blocked fp32 `torch.mm`
    (MKL sgemm, all host threads) + `torch.topk`, merged across blocks.  Same arithmetic as
    IndexFlatIP.search; used by bench.py's cpu_baseline / --impl reference legs because it is
    the fastest faithful CPU implementation available here.  Tie order follows torch.topk
    (unspecified), so it is a *timing* path; parity tests use IndexFlatIP above."""
    import torch

    Q = q_t.shape[0]
    N = corpus_t.shape[0]
    best_d = torch.full((Q, 0), 0.0)
    best_i = torch.zeros((Q, 0), dtype=torch.int64)
    for r0 in range(0, N, block_rows):
        blk = corpus_t[r0 : r0 + block_rows]
        s = torch.mm(q_t, blk.t())
        kk = min(k, blk.shape[0])
        d, i = torch.topk(s, kk, dim=1)
        best_d = torch.cat([best_d, d], dim=1)
        best_i = torch.cat([best_i, i + r0], dim=1)
        if best_d.shape[1] > k:
            d2, sel = torch.topk(best_d, k, dim=1)
            best_i = torch.gather(best_i, 1, sel)
            best_d = d2
    return best_d, best_i


# ---- faiss.index_factory(d, "Flat") -------------------------------------------------------------
# `FaissRetriever.__init__` (DRT/evaluator/index.py:49-54) calls `faiss.index_factory(d,
# factory_str)` WITHOUT a metric, and faiss' published signature is
# `index_factory(d, description, metric=METRIC_L2)`: the string "Flat" therefore yields an
# IndexFlatL2 (squared euclidean distances, ascending).  Restated here for the one exact string.
METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1
FLT_MAX = np.float32(3.4028234663852886e38)


class IndexFlatL2:
    """faiss.IndexFlatL2 restatement: `search` -> (D float32 [Q,k] squared L2 distances in
    ascending order, I int64 [Q,k]); canonical tie order (distance asc, id asc); `k > ntotal`
    pads with (+FLT_MAX, -1).  Distances are accumulated in float64 from the fp32 inputs."""

    metric_type = METRIC_L2

    def __init__(self, d: int):
        self.d = int(d)
        self.is_trained = True
        self.verbose = False
        self._blocks: list[np.ndarray] = []
        self.ntotal = 0

    def add(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise RuntimeError(f"add: expected [n,{self.d}] float32, got {x.shape}")
        self._blocks.append(x.copy())
        self.ntotal += x.shape[0]

    def train(self, x) -> None:
        return None

    def reset(self) -> None:
        self._blocks, self.ntotal = [], 0

    def reconstruct_n(self, i0: int, n: int) -> np.ndarray:
        allx = np.concatenate(self._blocks) if self._blocks else np.zeros((0, self.d), np.float32)
        return allx[i0 : i0 + n].copy()

    def search(self, q: np.ndarray, k: int):
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.d:
            raise RuntimeError(f"search: expected [nq,{self.d}] float32, got {q.shape}")
        x = (np.concatenate(self._blocks) if self._blocks else np.zeros((0, self.d), np.float32)).astype(np.float64)
        Q, N = q.shape[0], x.shape[0]
        D = np.full((Q, k), FLT_MAX, dtype=np.float32)
        I = np.full((Q, k), -1, dtype=np.int64)
        kk = min(k, N)
        ids = np.arange(N)
        xn = np.einsum("ij,ij->i", x, x)
        for r in range(Q):
            qr = q[r].astype(np.float64)
            dist = np.maximum(xn - 2.0 * (x @ qr) + qr @ qr, 0.0)
            if N <= 4096:                      # small cases: the literal definition sum((q-x)^2)
                dist = ((x - qr[None, :]) ** 2).sum(1)
            order = np.lexsort((ids, dist))[:kk]
            D[r, :kk] = dist[order].astype(np.float32)
            I[r, :kk] = order
        return D, I


def index_factory(d: int, description: str, metric: int = METRIC_L2):
    """faiss.index_factory for the exact string only (index.py:50)."""
    if description.strip() != "Flat":
        raise RuntimeError(f"oracle index_factory: only 'Flat' is restated, got {description!r}")
    return IndexFlatL2(d) if metric == METRIC_L2 else IndexFlatIP(d)


def flat_l2_search(corpus: np.ndarray, q: np.ndarray, k: int):
    idx = IndexFlatL2(corpus.shape[1])
    idx.add(corpus)
    return idx.search(q, k)
