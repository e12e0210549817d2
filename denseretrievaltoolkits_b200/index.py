"""Mirror of `DRT/evaluator/index.py`'s dense retrievers on the B200 store.

Same class names, constructor arguments and return conventions as the reference
(`BaseFaissIPRetriever` index.py:16-44, `FaissRetriever` index.py:47-54):
  * `BaseFaissIPRetriever(init_reps)`: ndarray -> d = init_reps.shape[1]; None -> no index;
    otherwise the argument is taken as the integer dimension.  `init_reps` is NOT added.
  * `.add(p_reps)` appends rows; ids are insertion order.
  * `.search(q_reps, k=1000)` returns ids only, int64 [Q,k], each row ordered by descending
    score (the reference re-orders with np.argsort(-scores), an identity on sorted rows; for
    the L2 index of `FaissRetriever(reps, "Flat")` that re-ordering is applied as upstream).
  * `.batch_search(q_reps, k, batch_size, quiet=False)` returns the concatenated ids.  (As
    written upstream it raises ValueError — index.py:40 unpacks `search`'s single array into
    two names; the intended behaviour, chunked search + concatenation, is implemented.)
`search_with_scores` / `batch_search_with_scores` additionally return the fp32 scores, which is
what `DRT/evaluator/retrieval.py` needs for its ranking files.
"""
from __future__ import annotations

import numpy as np

from . import faiss_compat as faiss


def _is_ndarray(x) -> bool:
    return isinstance(x, np.ndarray)


class BaseFaissIPRetriever:
    def __init__(self, init_reps):
        if _is_ndarray(init_reps) or (hasattr(init_reps, "shape") and len(getattr(init_reps, "shape")) == 2):
            index = faiss.IndexFlatIP(int(init_reps.shape[1]))
        elif init_reps is None:
            index = None
        else:
            index = faiss.IndexFlatIP(int(init_reps))
        self.index = index
        self.docid = []

    def add(self, p_reps):
        self.index.add(p_reps)

    def search_with_scores(self, q_reps, k: int = 1000):
        return self.index.search(q_reps, k)

    def search(self, q_reps, k: int = 1000):
        scores, indices = self.index.search(q_reps, k)
        if getattr(self.index, "metric_type", faiss.METRIC_INNER_PRODUCT) != faiss.METRIC_INNER_PRODUCT:
            # index.py:32-33 re-orders every row by argsort(-scores).  Rows of an inner-product
            # index are already in that order; the ascending distances of the L2 index that
            # `FaissRetriever(reps, "Flat")` holds come out farthest-first, as upstream.
            if _is_ndarray(scores):
                order = np.argsort(-scores, axis=1, kind="stable")
                return np.take_along_axis(indices, order, axis=1)
            import torch

            order = torch.argsort(-scores, dim=1, stable=True)
            return torch.gather(indices, 1, order)
        return indices

    def batch_search_with_scores(self, q_reps, k: int, batch_size: int, quiet: bool = False):
        from tqdm import tqdm

        num_query = q_reps.shape[0]
        all_scores, all_indices = [], []
        for start_idx in tqdm(range(0, num_query, batch_size), disable=quiet):
            s, i = self.index.search(q_reps[start_idx:start_idx + batch_size], k)
            all_scores.append(s)
            all_indices.append(i)
        if _is_ndarray(all_indices[0]):
            return np.concatenate(all_scores, axis=0), np.concatenate(all_indices, axis=0)
        import torch

        return torch.cat(all_scores, dim=0), torch.cat(all_indices, dim=0)

    def batch_search(self, q_reps, k: int, batch_size: int, quiet: bool = False):
        return self.batch_search_with_scores(q_reps, k, batch_size, quiet)[1]


class FaissRetriever(BaseFaissIPRetriever):
    def __init__(self, init_reps, factory_str: str):
        index = faiss.index_factory(int(init_reps.shape[1]), factory_str)
        self.index = index
        self.docid = []
        self.index.verbose = True
        if not self.index.is_trained:
            self.index.train(init_reps)
