"""Generates tests/golden/* in the AUTHORING container, where /root/reference is mounted.

The reference (yhao-wang/DenseRetrievalToolkits) ships no tests or golden vectors, so the pins
are produced by running the reference's own importable code on seeded inputs:
  * loss_*.npz    DRT.trainer.losses.SimpleContrastiveLoss forward + autograd backward
  * merge_*.json  DRT.model.utils.merge_retrieval_results_by_score
  * metrics.json  DRT.evaluator.metrics.get_metrics (consumer of the id lists)
  * search_*.npz  DRT/evaluator/index.py's BaseFaissIPRetriever executed, unmodified, over a
                  faiss-shaped stub whose IndexFlatIP is oracle/flat_ip.py (faiss itself is not
                  installable here), plus a float64 brute force of the same inputs
  * search_factory_flat.npz  FaissRetriever (index.py:47-54) unmodified over the same stub
  * bm25negatives.jsonl + mining_samples.json  the mined-negatives file as the reference's own
                  BM25Negatives.save writes it and its load_passages cache branch reads it
  * mining.json   the loop of process_sample (DRT/trainer/sampler.py:73-78) restated verbatim
                  (the closure lives in a module that needs faiss at import time)
Run:  python tools/make_golden.py        (needs /root/reference; tests never do)
"""
from __future__ import annotations

import json
import os
import sys
import types
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

from oracle import flat_ip  # noqa: E402


def loss_goldens():
    from DRT.trainer.losses import SimpleContrastiveLoss

    fn = SimpleContrastiveLoss()
    cases = [("b5n3", 5, 3, 64, None, "mean"), ("b16n2", 16, 2, 64, None, "mean"),
             ("b16n2_sum", 16, 2, 64, None, "sum"), ("b16n2_none", 16, 2, 64, None, "none"),
             ("b8n4_target", 8, 4, 64, "rand", "mean"), ("b128n8", 128, 8, 768, None, "mean")]
    for name, B, n, d, tgt, red in cases:
        seed = zlib.crc32(name.encode()) % 100000
        rng = np.random.default_rng(seed)
        scale = 1.0 if d == 768 else 0.5
        x = (rng.standard_normal((B, d)) * scale).astype(np.float32)
        y = (rng.standard_normal((B * n, d)) * scale).astype(np.float32)
        target = None
        if tgt == "rand":
            target = rng.integers(0, B * n, size=B).astype(np.int64)
        xt = torch.from_numpy(x).requires_grad_(True)
        yt = torch.from_numpy(y).requires_grad_(True)
        loss = fn(xt, yt, target=None if target is None else torch.from_numpy(target), reduction=red)
        (loss.sum() if red == "none" else loss).backward()
        payload = dict(loss=loss.detach().numpy(), reduction=np.array(red), B=B, n=n, d=d)
        if target is not None:
            payload["target"] = target
        if d <= 64:
            payload.update(x=x, y=y, dx=xt.grad.numpy(), dy=yt.grad.numpy())
        else:   # big case: inputs are regenerated from the seed (heads guard against RNG drift);
                # gradients are kept as a row subset + column sums to keep the fixture small
            payload.update(seed=np.array(seed), x_head=x[:2], y_head=y[:2], dx_rows=xt.grad.numpy()[:8],
                           dy_rows=yt.grad.numpy()[:16], dx_colsum=xt.grad.numpy().sum(0),
                           dy_colsum=yt.grad.numpy().sum(0))
        np.savez_compressed(os.path.join(OUT, f"loss_{name}.npz"), **payload)
        print("loss", name, float(np.asarray(loss.detach()).sum()))


def merge_goldens():
    from DRT.model.utils import merge_retrieval_results_by_score

    rng = np.random.default_rng(11)
    cases = []
    for ci, (G, Q, k, topk, overlap) in enumerate([(3, 4, 5, 6, False), (2, 3, 8, 8, True), (4, 2, 3, 20, False)]):
        results = []
        for g in range(G):
            res = {}
            for q in range(Q):
                if overlap:
                    ids = rng.choice(12, size=k, replace=False)
                else:
                    ids = g * 100 + rng.choice(50, size=k, replace=False)
                res[f"q{q}"] = {str(int(i)): float(np.float32(rng.standard_normal())) for i in ids}
            results.append(res)
        merged = merge_retrieval_results_by_score(results, topk=topk)
        cases.append(dict(results=results, topk=topk,
                          merged={q: list(v.items()) for q, v in merged.items()}))
    json.dump(cases, open(os.path.join(OUT, "merge_cases.json"), "w"))
    print("merge cases", len(cases))


def metrics_goldens():
    from DRT.evaluator.metrics import get_metrics

    rng = np.random.default_rng(5)
    hits = (rng.random((6, 20)) < 0.15).astype(np.int8)
    m = get_metrics(hits, [1, 5, 10, 20])
    json.dump(dict(hits=hits.tolist(), topk=[1, 5, 10, 20], metrics={k: float(v) for k, v in m.items()}),
              open(os.path.join(OUT, "metrics.json"), "w"))
    print("metrics", m)


def search_goldens():
    # run the reference wrapper (index.py:16-44) unmodified over a faiss-shaped stub
    stub = types.ModuleType("faiss")
    stub.IndexFlatIP = flat_ip.IndexFlatIP
    sys.modules["faiss"] = stub
    from DRT.evaluator.index import BaseFaissIPRetriever

    rng = np.random.default_rng(21)
    d = 64
    cases = {}

    def run(name, x, q, k):
        r = BaseFaissIPRetriever(x)          # ndarray ctor: creates an EMPTY index (index.py:18-19)
        assert r.index.ntotal == 0
        r.add(x)
        ids = r.search(q, k)                 # ids only (index.py:31-33)
        D, I = r.index.search(q, k)
        D64, I64 = flat_ip.flat_ip_search_f64(x, q, k)
        cases[name] = dict(x=x, q=q, k=np.array(k), wrapper_ids=ids, D=D, I=I, D64=D64.astype(np.float64), I64=I64)
        try:
            r.batch_search(q, k, 2, quiet=True)
            cases[name]["batch_search_raises"] = np.array(0)
        except ValueError:
            cases[name]["batch_search_raises"] = np.array(1)   # upstream bug (index.py:40)

    x = rng.standard_normal((1000, d)).astype(np.float32)
    q = rng.standard_normal((7, d)).astype(np.float32)
    run("basic", x, q, 10)
    run("k_gt_n", x[:6], q[:3], 9)
    xd = x[:300].copy(); xd[200:300] = xd[:100]              # duplicated rows -> exact ties
    run("ties", xd, q, 20)
    xz = x[:200].copy(); xz[10] = 0; xz[11] = 0               # zero vectors
    run("zeros", xz, q[:4], 15)
    xn = x[:200].copy(); xn[5, 0] = np.nan; xn[6, 1] = np.inf; xn[7, 2] = -np.inf
    qn = np.abs(q[:4]) + 0.1                                  # positive queries: inf rows -> +/-inf scores
    run("nonfinite", xn, qn, 12)
    for name, c in cases.items():
        np.savez_compressed(os.path.join(OUT, f"search_{name}.npz"), **c)
        print("search", name, c["wrapper_ids"].shape, "batch_search_raises", int(c["batch_search_raises"]))

    # FaissRetriever (index.py:47-54), unmodified, over the stub: index_factory(d, "Flat") without a
    # metric is faiss' default METRIC_L2, and the inherited search() re-orders the ascending
    # distances by argsort(-scores)
    stub.index_factory = flat_ip.index_factory
    from DRT.evaluator.index import FaissRetriever

    r = FaissRetriever(x, "Flat")
    assert r.index.ntotal == 0 and r.index.verbose is True
    r.add(x)
    ids = r.search(q, 10)
    D, I = r.index.search(q, 10)
    np.savez_compressed(os.path.join(OUT, "search_factory_flat.npz"), x=x, q=q, k=np.array(10), wrapper_ids=ids, D=D, I=I,
                        metric=np.array(r.index.metric_type))
    print("search factory_flat", ids.shape, "metric", r.index.metric_type)


def mining_goldens():
    rng = np.random.default_rng(31)
    cases = []
    for _ in range(4):
        k, num_negative = 30, 8
        neg_docs = [int(v) for v in rng.choice(200, size=k, replace=False)]
        b = int(rng.integers(0, 190)); e = b + int(rng.integers(1, 40))
        document = []
        for doc in neg_docs:                     # sampler.py:73-78, verbatim
            if doc >= b and doc < e:
                continue
            document.append(doc)
            if len(document) == num_negative:
                break
        cases.append(dict(ids=neg_docs, b=b, e=e, num_negative=num_negative, kept=document))
    json.dump(cases, open(os.path.join(OUT, "mining.json"), "w"))
    print("mining cases", len(cases))


def mining_jsonl_golden():
    """The on-disk record layout of mined negatives: written by the reference's own
    `BM25Negatives.save` (sampler.py:89-99) and read back by the cache branch of its
    `load_passages` (sampler.py:57-66), both imported and run unmodified (faiss stubbed)."""
    import shutil
    import tempfile

    stub = types.ModuleType("faiss")
    stub.IndexFlatIP = flat_ip.IndexFlatIP
    sys.modules.setdefault("faiss", stub)
    import DRT.trainer.sampler as sampler_mod
    from DRT.trainer.sampler import BM25Negatives

    # upstream bug: load_passages returns `ListDataset(data)` (sampler.py:99), a name that is not
    # defined anywhere (the class in that file is BM25Dataset, sampler.py:12-20) -> NameError after
    # the file was read.  Bind the name so the reference's own read loop can be run.
    sampler_mod.ListDataset = sampler_mod.BM25Dataset

    rng = np.random.default_rng(41)
    passages = [[int(t) for t in rng.integers(1000, 30000, size=int(rng.integers(5, 12)))] for _ in range(60)]
    passages[7] = [101, 2054, 2003, 1996, 3007, 1997, 2605, 102]
    samples, neg_ids = [], []
    for i in range(9):
        b = 5 * i
        samples.append({"query": [int(t) for t in rng.integers(1000, 30000, size=6)],
                        "positives": passages[b:b + 1 + i % 3]})
        row = [int(v) for v in rng.choice(60, size=4, replace=False)]
        if i % 4 == 0:
            row[-1] = -1                        # fewer survivors than num_negative: padded with -1
        neg_ids.append(row)
    samples[3]["query"] = "naïve café — 東京"      # a text query: ensure_ascii=False on disk
    expect = []
    for smp, row in zip(samples, neg_ids):
        rec = dict(smp)
        rec["negatives"] = [passages[j] for j in row if j >= 0]
        expect.append(rec)
    tmp = tempfile.mkdtemp()
    try:
        obj = object.__new__(BM25Negatives)
        obj.cache_dir = tmp
        BM25Negatives.save(obj, expect, os.path.join(tmp, "BM25data"), "bm25negatives")
        raw = open(os.path.join(tmp, "BM25data", "bm25negatives"), "rb").read()
        loaded = BM25Negatives.load_passages(obj, corpus=None)            # cache branch: reads the file back
        assert [loaded[i] for i in range(len(loaded))] == expect
    finally:
        shutil.rmtree(tmp)
    open(os.path.join(OUT, "bm25negatives.jsonl"), "wb").write(raw)
    json.dump(dict(samples=samples, neg_ids=neg_ids, passages=passages), open(os.path.join(OUT, "mining_samples.json"), "w"),
              ensure_ascii=False)
    print("mining jsonl golden:", len(expect), "records,", len(raw), "bytes")


def eval_goldens():
    """has_answers (nq_eval.py:203-218) and get_metrics (metrics.py:50-59) run by the reference."""
    from DRT.evaluator.metrics import get_metrics
    from DRT.evaluator.nq_eval import has_answers

    texts = [
        "The Eiffel Tower was completed in 1889, in Paris (France).",
        "Émile Zola wrote J'accuse…! in 1898; café culture thrived.",
        "He scored 3-2 in the U.S. Open; it's a record-breaking 100m dash!",
        "東京 is the capital of Japan. Tokyo-to has 14 million people.",
        "",
        "new york city, New York, NEW YORK",
        "The answer is forty-two (42).",
    ]
    answers = [["Paris"], ["paris", "London"], ["1889"], ["Zola"], ["emile zola"], ["Émile Zola"], ["J'accuse"],
               ["U.S. Open"], ["us open"], ["3-2"], ["100m"], ["東京"], ["tokyo-to"], [""], ["New York City"],
               ["york new"], ["forty two"], ["(42)"], ["France)."], ["completed in 1889 ,"], ["café"], ["cafe"]]
    cases = []
    for t in texts:
        for a in answers:
            cases.append(dict(text=t, answers=a, regex=False, hit=bool(has_answers(t, a))))
    for t, a in [(texts[0], [r"18\\d\\d"]), (texts[0], [r"paris|london"]), (texts[2], [r"\\d+-\\d+"]), (texts[2], ["(unclosed"]),
                 (texts[1], [r"caf."])]:
        cases.append(dict(text=t, answers=a, regex=True, hit=bool(has_answers(t, a, regex=True))))
    rng = np.random.default_rng(77)
    mcases = []
    for Q, K, p, topk in [(6, 20, 0.15, [1, 5, 10, 20]), (16, 100, 0.03, [5, 10, 20, 50, 100]), (4, 10, 0.0, [1, 5, 10]),
                          (5, 8, 0.9, [1, 3, 8, 20])]:
        hits = (rng.random((Q, K)) < p).astype(np.int8)
        m = get_metrics(hits, topk)
        mcases.append(dict(hits=hits.tolist(), topk=topk, metrics={k: float(v) for k, v in m.items()}))
    json.dump(dict(has_answers=cases, metrics=mcases), open(os.path.join(OUT, "evaluation.json"), "w"), ensure_ascii=False)
    print("evaluation goldens:", len(cases), "has_answers cases,", sum(c["hit"] for c in cases), "hits;", len(mcases), "metric cases")


if __name__ == "__main__":
    eval_goldens()
    loss_goldens()
    merge_goldens()
    metrics_goldens()
    search_goldens()
    mining_goldens()
    mining_jsonl_golden()
