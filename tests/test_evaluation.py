"""CPU: the post-search evaluation step (has_answers + metrics) against goldens produced by the
reference's own nq_eval.has_answers / metrics.get_metrics (tools/make_golden.py)."""
import json
import os

import numpy as np

from denseretrievaltoolkits_b200 import evaluation as ev

GOLD = os.path.join(os.path.dirname(__file__), "golden", "evaluation.json")


def test_has_answers_matches_reference():
    cases = json.load(open(GOLD))["has_answers"]
    assert sum(c["hit"] for c in cases) > 10
    for c in cases:
        assert ev.has_answers(c["text"], c["answers"], regex=c["regex"]) == c["hit"], c


def test_get_metrics_matches_reference():
    for c in json.load(open(GOLD))["metrics"]:
        got = ev.get_metrics(np.array(c["hits"], dtype=np.int8), c["topk"])
        assert set(got) == set(c["metrics"])
        for k, v in c["metrics"].items():
            assert abs(got[k] - v) <= 1e-12 * max(1.0, abs(v)), (k, got[k], v)


def test_hits_matrix_equals_per_pair_calls_and_uses_pool():
    cases = json.load(open(GOLD))["has_answers"]
    texts = sorted({c["text"] for c in cases})
    answers = [c["answers"] for c in cases if not c["regex"]][:22]
    docs = [texts for _ in answers]
    ids = [list(range(len(texts))) for _ in answers]
    want = np.array([[ev.has_answers(t, a) for t in texts] for a in answers], dtype=np.int8)
    np.testing.assert_array_equal(ev.hits_matrix(docs, answers, doc_ids=ids, workers=1), want)
    ev._POOL_MIN_DOCS = 0                                   # force the (persistent) process pool on this small case
    try:
        np.testing.assert_array_equal(ev.hits_matrix(docs * 4, answers * 4, doc_ids=ids * 4, workers=2), np.tile(want, (4, 1)))
        first = ev._POOL_STATE["pool"]
        np.testing.assert_array_equal(ev.hits_matrix(docs * 4, answers * 4, doc_ids=ids * 4, workers=2), np.tile(want, (4, 1)))
        assert first is not None and ev._POOL_STATE["pool"] is first        # the pool is reused between calls
    finally:
        ev._POOL_MIN_DOCS = 20000
        ev.shutdown_pool()
    np.testing.assert_array_equal(ev.hits_matrix(docs, answers, workers=1), want)     # no doc-id cache


def test_reduce_metrics_single_process():
    out = ev.reduce_metrics({"Recall@5": 3.0, "MRR@5": 1.5, "query_num": 0}, 4)
    assert out == {"MRR@5": 0.375, "Recall@5": 0.75, "query_num": 4}
