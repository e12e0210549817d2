"""Post-search evaluation step of `Trainer.evaluate`: answer matching over
the retrieved passages and Recall / MRR / NDCG, batched and parallelised on the CPU, plus the
cross-rank reduction the reference never does.

Mirrors, value for value:
  * `has_answers(text, answers, tokenizer, regex=False)` and `SimpleTokenizer`
    (`DRT/evaluator/nq_eval.py:145-218`): NFD-normalise, tokenise with
    `([\\p{L}\\p{N}\\p{M}]+)|([^\\p{Z}\\p{C}])` (ignore-case, unicode, multiline), lower-case, and
    test whether an answer's token sequence occurs contiguously in the passage's token sequence;
    with `regex=True` the answers are patterns searched in the normalised text.
  * `get_metrics(indices, topk)` (`DRT/evaluator/metrics.py:4-59`): Recall@k and MRR@k are SUMS
    over the batch of first-hit indicators / reciprocal ranks, NDCG@k is (sum of DCG) / (sum of
    ideal DCG) over the batch with natural-log discounts — exactly as the reference computes them
    (the trainer divides by the number of queries later, `trainer.py:338-340`).

This is CPU string work (not a GPU target); what changes against the reference's loop
(`trainer.py:302-311`: Q·k sequential `has_answers` calls that re-tokenise every answer for every
passage) is: answers are tokenised once per query, passages once per doc id (the same passage is
retrieved by many queries), and queries are spread over a process pool.
"""
from __future__ import annotations

import atexit
import math
import os
import unicodedata
from concurrent.futures import ProcessPoolExecutor
from typing import Dict, Hashable, List, Optional, Sequence

import numpy as np

try:
    import regex as _regex
except ImportError as e:  # pragma: no cover
    raise ImportError("denseretrievaltoolkits_b200.evaluation needs the `regex` package (as DRT.evaluator.nq_eval does)") from e
import re as _re

# SimpleTokenizer's pattern (nq_eval.py:147-160) without its two capture groups — the matches
# are the same, and a group-free pattern lets `findall` return the tokens from C
_TOKEN = _regex.compile(r"[\p{L}\p{N}\p{M}]+|[^\p{Z}\p{C}]",
                        flags=_regex.IGNORECASE + _regex.UNICODE + _regex.MULTILINE)


def tokenize_uncased(text: str) -> List[str]:
    """SimpleTokenizer.tokenize(text).words(uncased=True) (nq_eval.py:145-185)."""
    return [w.lower() for w in _TOKEN.findall(text)]


def _normalize(text: str) -> str:
    return unicodedata.normalize("NFD", text)


def _contains(haystack: List[str], needle: List[str]) -> bool:
    n, m = len(haystack), len(needle)
    if m == 0:
        return n >= 0        # the reference's range(0, n - 0 + 1) loop matches the empty answer
    first = needle[0]
    for i in range(0, n - m + 1):
        if haystack[i] == first and haystack[i:i + m] == needle:
            return True
    return False


def _regex_match(text: str, pattern: str) -> bool:
    try:
        compiled = _re.compile(pattern, flags=_re.IGNORECASE + _re.UNICODE + _re.MULTILINE)
    except BaseException:
        return False
    return compiled.search(text) is not None


def has_answers(text: str, answers: Sequence[str], tokenizer=None, regex: bool = False) -> bool:
    """Drop-in for nq_eval.has_answers (the `tokenizer` argument is accepted and ignored: the
    tokenisation rule is fixed to SimpleTokenizer's)."""
    text = _normalize(text)
    if regex:
        return any(_regex_match(text, _normalize(a)) for a in answers)
    words = tokenize_uncased(text)
    return any(_contains(words, tokenize_uncased(_normalize(a))) for a in answers)


def _hits_for_queries(args):
    docs_rows, doc_ids_rows, answers_rows, use_regex = args
    cache: Dict[Hashable, object] = {}
    out = np.zeros((len(docs_rows), max((len(r) for r in docs_rows), default=0)), dtype=np.int8)
    for qi, (docs, ids, answers) in enumerate(zip(docs_rows, doc_ids_rows, answers_rows)):
        if use_regex:
            pats = [_normalize(a) for a in answers]
        else:
            needles = [tokenize_uncased(_normalize(a)) for a in answers]
        for j, doc in enumerate(docs):
            key = ids[j] if ids is not None else None
            if use_regex:
                text = cache.get(key) if key is not None else None
                if text is None:
                    text = _normalize(doc)
                    if key is not None:
                        cache[key] = text
                hit = any(_regex_match(text, p) for p in pats)
            else:
                words = cache.get(key) if key is not None else None
                if words is None:
                    words = tokenize_uncased(_normalize(doc))
                    if key is not None:
                        cache[key] = words
                hit = any(_contains(words, nd) for nd in needles)
            out[qi, j] = 1 if hit else 0
    return out


# Passages per call below which the pool's pickling overhead outweighs the parallel speed-up
# (measured: 12.8k passages take 0.4 s on one core).  The pool is created once and kept:
# `Trainer.evaluate` calls this per query batch and worker start-up costs ~0.5 s.
_POOL_MIN_DOCS = 20000
_POOL_STATE = {"pool": None, "workers": 0}


def _pool(workers: int) -> ProcessPoolExecutor:
    st = _POOL_STATE
    if st["pool"] is None or st["workers"] != workers:
        if st["pool"] is not None:
            st["pool"].shutdown(wait=False, cancel_futures=True)
        # forkserver, not fork: the caller is a trainer process that holds a CUDA context and
        # NCCL / watchdog threads, and forking a multi-threaded CUDA process can deadlock the
        # children on inherited locks.  The workers only need this module.
        import multiprocessing

        st["pool"] = ProcessPoolExecutor(max_workers=workers, mp_context=multiprocessing.get_context("forkserver"))
        st["workers"] = workers
    return st["pool"]


def shutdown_pool() -> None:
    """Stop the worker processes `hits_matrix` keeps between calls."""
    if _POOL_STATE["pool"] is not None:
        _POOL_STATE["pool"].shutdown(wait=False, cancel_futures=True)      # never block interpreter exit
        _POOL_STATE["pool"], _POOL_STATE["workers"] = None, 0


atexit.register(shutdown_pool)


def hits_matrix(docs: Sequence[Sequence[str]], answers: Sequence[Sequence[str]],
                doc_ids: Optional[Sequence[Sequence[Hashable]]] = None, regex: bool = False,
                workers: Optional[int] = None) -> np.ndarray:
    """The `pos_index` matrix of trainer.py:298-311: out[q, j] = 1 iff retrieved passage j of
    query q contains one of the query's answers.  `doc_ids` (same shape as `docs`) enables the
    per-passage tokenisation cache."""
    Q = len(docs)
    if Q == 0:
        return np.zeros((0, 0), dtype=np.int8)
    ids = doc_ids if doc_ids is not None else [None] * Q
    if workers is None:
        workers = min(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1), 16)
    n_docs = sum(len(r) for r in docs)
    if workers <= 1 or Q < 2 * workers or n_docs < _POOL_MIN_DOCS:
        return _hits_for_queries((docs, ids, answers, regex))
    step = -(-Q // workers)
    jobs = [(docs[s:s + step], ids[s:s + step], answers[s:s + step], regex) for s in range(0, Q, step)]
    parts = list(_pool(workers).map(_hits_for_queries, jobs))
    width = max(p.shape[1] for p in parts)
    parts = [np.pad(p, ((0, 0), (0, width - p.shape[1]))) for p in parts]
    return np.concatenate(parts, axis=0)


def get_metrics(indices, topk: Sequence[int]) -> Dict[str, float]:
    """Same keys and values as DRT.evaluator.metrics.get_metrics (metrics.py:50-59)."""
    ind = np.asarray(indices) != 0
    if ind.ndim != 2:
        ind = ind.reshape(len(indices), -1)
    Q, K = ind.shape
    any_hit = ind.any(axis=1)
    first = np.where(any_hit, ind.argmax(axis=1), K)                 # position of the first hit
    disc = 1.0 / np.log(np.arange(K) + 2.0)                           # natural log, as metrics.py:38
    cnt = ind.sum(axis=1)
    result: Dict[str, float] = {}
    rec, mrr_, ndcg_ = [], [], []
    for k in topk:
        hit_k = any_hit & (first < k)
        rec.append(int(hit_k.sum()))
        mrr_.append(float((1.0 / (first[hit_k] + 1.0)).sum()) if hit_k.any() else 0)
        kk = min(k, K)
        dcg = float((ind[:, :kk] * disc[:kk]).sum())
        ideal_len = np.minimum(np.maximum(cnt, 1), k)                # metrics.py:39-42
        cum = np.concatenate([[0.0], np.cumsum(1.0 / np.log(np.arange(max(int(ideal_len.max()), 1)) + 2.0))])
        idcg = float(cum[ideal_len].sum())
        ndcg_.append(dcg / idcg)
    for name, data in zip(["Recall@", "MRR@", "NDCG@"], [rec, mrr_, ndcg_]):
        for k, v in zip(topk, data):
            result[name + str(k)] = v
    return result


def reduce_metrics(m_all: Dict[str, float], eval_num: int, group=None) -> Dict[str, float]:
    """Cross-rank reduction (the reference writes one metrics file per rank and never reduces,
    trainer.py:338-345): sums the per-rank accumulated metric sums and query counts and returns
    the global averages (same division as trainer.py:338-339)."""
    import torch
    import torch.distributed as dist

    keys = sorted(k for k in m_all if k != "query_num")
    vec = torch.tensor([float(m_all[k]) for k in keys] + [float(eval_num)], dtype=torch.float64)
    if dist.is_available() and dist.is_initialized():
        if dist.get_backend(group) == "nccl":
            vec = vec.cuda()
        dist.all_reduce(vec, group=group)
        vec = vec.cpu()
    total = float(vec[-1])
    out = {k: float(v) / total if total > 0 else 0.0 for k, v in zip(keys, vec[:-1])}
    out["query_num"] = int(total)
    return out
