"""CPU oracle for the cross-shard candidate merge and the mining filter.  TEST INFRASTRUCTURE
ONLY (tests/, smoke(), bench baseline legs).

Restates
  * `merge_retrieval_results_by_score(results, topk)` (`DRT/model/utils.py:215-229`): per query,
    union of the partitions' {doc_id: score} dicts where the FIRST partition that mentions a
    doc id wins, then sort by score descending and keep `topk`.  Python's sort is stable, so
    equal scores keep first-seen order; the canonical order used here and by the CUDA merge is
    (score desc, id asc), which coincides whenever partitions hold disjoint ascending id
    ranges (the row-sharded store).
  * `process_sample` (`DRT/trainer/sampler.py:69-80`): walk retrieved ids in rank order, skip
    ids inside the query's own positive range [b, e), keep the first `num_negative`.

PINNED: `merge_retrieval_results_by_score` is importable here; tools/make_golden.py ran it and
committed tests/golden/merge_*.json; `process_sample` is a closure inside
`BM25Negatives.load_passages` (needs faiss at import) so it is pinned by restating its loop
verbatim in tools/make_golden.py against the same cases.
"""
from __future__ import annotations

import numpy as np

FLT_LOWEST = np.float32(-3.4028234663852886e38)


def merge_topk(scores: np.ndarray, ids: np.ndarray, k_out: int):
    """scores/ids: [G, Q, k_in]; ids < 0 are padding.  Returns (D [Q,k_out], I [Q,k_out])."""
    G, Q, k_in = scores.shape
    D = np.full((Q, k_out), FLT_LOWEST, np.float32)
    I = np.full((Q, k_out), -1, np.int64)
    for q in range(Q):
        seen: dict[int, float] = {}
        for g in range(G):  # first partition mentioning an id wins (utils.py:224-226)
            for j in range(k_in):
                i = int(ids[g, q, j])
                if i >= 0 and i not in seen:
                    seen[i] = scores[g, q, j]
        if not seen:
            continue
        sid = np.fromiter(seen.keys(), np.int64, len(seen))
        ssc = np.fromiter(seen.values(), np.float32, len(seen))
        order = np.lexsort((sid, -ssc.astype(np.float64)))[:k_out]
        D[q, : order.size] = ssc[order]
        I[q, : order.size] = sid[order]
    return D, I


def filter_negatives(ids: np.ndarray, pos_begin: np.ndarray, pos_end: np.ndarray,
                     num_negative: int) -> np.ndarray:
    """ids [Q,k] in rank order -> [Q,num_negative] (sampler.py:73-78); unfilled = -1."""
    Q, k = ids.shape
    out = np.full((Q, num_negative), -1, np.int64)
    for q in range(Q):
        n = 0
        for j in range(k):
            doc = int(ids[q, j])
            if doc < 0:
                continue
            if pos_begin[q] <= doc < pos_end[q]:  # sampler.py:74-75
                continue
            out[q, n] = doc
            n += 1
            if n == num_negative:  # sampler.py:77-78
                break
    return out
