"""Batching queries across loader steps.

`Trainer.evaluate` (DRT/trainer/trainer.py:287-311) encodes one loader batch — `eval_batch_size`
16 (run.sh:30) to 128 (DRT/arguments.py:189) queries — and searches it at once, so every call
streams the whole corpus shard from HBM for a handful of queries: 16 queries per 2.1 ms pass over
8.8M x 768 bf16 rows is 6k queries/s, against 100k+ when thousands of queries share a pass (the
search is HBM-bound below ~250 queries per pass).

`DeferredSearch` keeps the encoder outputs of successive steps on the device, searches them in
ONE pass once `max_queries` have accumulated (or on `flush()`), and hands the results back per
step, in order — bit-identical to per-step calls (searches are batch-size independent), so the
metric code after the search (trainer.py:298-345) is unchanged; it just runs a few steps later.

    ds = DeferredSearch(self.store, k=retrieve_num, max_queries=4096)     # store or index or retriever
    for batch in query_loader:                                            # trainer.py:287
        q_reps = self.model(query=batch[1]).q_reps.detach()               # stays on the GPU (no .cpu(), :295)
        for tag, (D, I) in ds.add(q_reps, tag=batch):                     # usually empty; full batches when flushed
            self._score_batch(tag, I)                                     # trainer.py:298-321, unchanged
    for tag, (D, I) in ds.flush():
        self._score_batch(tag, I)
"""
from __future__ import annotations

from typing import Any, Iterator, List, Tuple

import numpy as np


class DeferredSearch:
    """Accumulate query batches, search once, return per-step slices in submission order.

    `target` is anything with the search contract of this package: a `ShardedCorpusStore`
    (its `search_local_queries` is used under an initialised process group: every rank adds its
    OWN batches, equal sizes per step, and `add` / `flush` are collective), a `faiss_compat`
    index (`search(x, k) -> (D, I)`), or a `BaseFaissIPRetriever` (`search_with_scores`)."""

    def __init__(self, target, k: int, max_queries: int = 4096):
        if int(k) <= 0 or int(max_queries) <= 0:
            raise ValueError("k and max_queries must be positive")
        self.k = int(k)
        self.max_queries = int(max_queries)
        if hasattr(target, "search_local_queries"):
            self._search = lambda q: target.search_local_queries(q, self.k)
        elif hasattr(target, "search_with_scores"):
            self._search = lambda q: target.search_with_scores(q, self.k)
        else:
            self._search = lambda q: target.search(q, self.k)
        self._pending: List[Tuple[Any, Any]] = []      # (tag, queries)
        self._count = 0
        self.searches = 0                               # corpus passes issued so far

    def __len__(self) -> int:
        return self._count

    def add(self, q_reps, tag: Any = None) -> List[Tuple[Any, Tuple[Any, Any]]]:
        """Queue one step's queries ([n,d] CUDA tensor or numpy array; all steps of one flush must
        be of one kind).  Returns the results that became available: [] until `max_queries` are
        queued, then one `(tag, (D, I))` per queued step, in order."""
        if getattr(q_reps, "ndim", 0) != 2:
            raise RuntimeError(f"DeferredSearch.add: expected [n,d] queries, got shape {getattr(q_reps, 'shape', None)}")
        self._pending.append((tag, q_reps))
        self._count += int(q_reps.shape[0])
        return self.flush() if self._count >= self.max_queries else []

    def flush(self) -> List[Tuple[Any, Tuple[Any, Any]]]:
        """Search everything queued in one pass and return `(tag, (D, I))` per step, in order."""
        if not self._pending:
            return []
        pend, self._pending, self._count = self._pending, [], 0
        qs = [q for _, q in pend]
        if isinstance(qs[0], np.ndarray):
            allq = qs[0] if len(qs) == 1 else np.concatenate(qs, axis=0)
        else:
            import torch

            allq = qs[0] if len(qs) == 1 else torch.cat(qs, dim=0)
        D, I = self._search(allq)
        self.searches += 1
        out, r0 = [], 0
        for tag, q in pend:
            n = int(q.shape[0])
            out.append((tag, (D[r0:r0 + n], I[r0:r0 + n])))
            r0 += n
        return out

    def results(self, batches) -> Iterator[Tuple[Any, Tuple[Any, Any]]]:
        """Convenience: feed an iterable of `(tag, q_reps)` and iterate over `(tag, (D, I))`."""
        for tag, q in batches:
            yield from self.add(q, tag)
        yield from self.flush()
