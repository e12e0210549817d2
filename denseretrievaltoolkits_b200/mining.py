"""Dense hard-negative mining on the B200 store.

The reference mines negatives lexically (`BM25Negatives`, DRT/trainer/sampler.py:49-99):
retrieve `num_negative + len(positives)` documents per training query, drop the ones inside
the query's own positive id range [b, e) (sampler.py:73-76), keep the first `num_negative`
(sampler.py:77-78), and emit {'query', 'positives', 'negatives'} samples (sampler.py:79,95-99).
This module keeps that contract but retrieves with exact dense MIPS (`index.search`, the
a1-a3 path of SURVEY.md §8) and applies the exclusion rule on the device
(`drt_filter_negatives`).
"""
from __future__ import annotations

import json
from typing import Iterable, Optional

import torch

from . import _lib


def filter_negatives(ids: torch.Tensor, pos_begin: torch.Tensor, pos_end: torch.Tensor,
                     num_negative: int) -> torch.Tensor:
    """ids [Q,k] int64 CUDA, rank order -> [Q,num_negative] (unfilled = -1)."""
    if not ids.is_cuda:
        raise RuntimeError("filter_negatives needs CUDA tensors: there is no CPU fallback")
    lib = _lib.load()
    ids = ids.contiguous().to(torch.int64)
    dev = ids.device
    pb = pos_begin.to(device=dev, dtype=torch.int64).contiguous()
    pe = pos_end.to(device=dev, dtype=torch.int64).contiguous()
    Q, k = ids.shape
    out = torch.empty((Q, num_negative), dtype=torch.int64, device=dev)
    _lib.check(lib.drt_filter_negatives(ids.data_ptr(), Q, k, pb.data_ptr(), pe.data_ptr(), int(num_negative),
                                        out.data_ptr(), dev.index, _lib.current_stream_ptr(dev.index)),
               "filter_negatives")
    return out


def mine_hard_negatives(index, q_reps: torch.Tensor, pos_begin: torch.Tensor, pos_end: torch.Tensor,
                        num_negative: int, depth: Optional[int] = None, batch_size: int = 8192) -> torch.Tensor:
    """For every query: top-`depth` dense retrieval (default num_negative + the largest positive
    range, sampler.py:72), positive exclusion, first `num_negative` survivors.  `index` is a
    `faiss_compat.IndexFlatIP` or a `ShardedCorpusStore`; queries are streamed in batches."""
    if depth is None:
        depth = int(num_negative + (pos_end - pos_begin).max().item())
    outs = []
    for s in range(0, q_reps.shape[0], batch_size):
        _, ids = index.search(q_reps[s:s + batch_size], depth)
        if not torch.is_tensor(ids):
            ids = torch.from_numpy(ids).cuda()
        outs.append(filter_negatives(ids, pos_begin[s:s + batch_size], pos_end[s:s + batch_size], num_negative))
    return torch.cat(outs, dim=0)


def write_negatives_jsonl(path: str, samples: Iterable[dict], neg_ids, passages) -> None:
    """Same record layout `BM25Negatives.save` writes (sampler.py:95-99): one JSON object per
    line with 'query', 'positives' and the mined 'negatives' (passage payloads looked up by id)."""
    with open(path, "w", encoding="utf-8") as f:
        for sample, row in zip(samples, neg_ids.tolist() if hasattr(neg_ids, "tolist") else neg_ids):
            rec = dict(sample)
            rec["negatives"] = [passages[i] for i in row if i >= 0]
            json.dump(rec, f, ensure_ascii=False)
            f.write("\n")
