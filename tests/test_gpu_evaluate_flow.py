"""GPU: the whole Trainer.evaluate flow (DRT/trainer/trainer.py:269-346) on the new pieces:
encoder outputs stay on the device -> sharded store -> per-batch search -> doc-id mapping ->
answer matching -> Recall/MRR/NDCG, checked against the same flow run through the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_evaluate_flow_matches_oracle_flow():
    from denseretrievaltoolkits_b200.evaluation import get_metrics, hits_matrix, reduce_metrics
    from denseretrievaltoolkits_b200.store import ShardedCorpusStore
    from oracle import flat_ip

    rng = np.random.default_rng(0)
    n_docs, d, n_q, k, topk = 20000, 768, 96, 100, [5, 10, 20, 50, 100]
    doc_ids = rng.permutation(n_docs) + 1000                          # external doc ids (self.idx)
    texts = {int(i): f"Passage {int(i)} mentions the entity ent{int(i) % 977} in passing." for i in doc_ids}
    emb = rng.standard_normal((n_docs, d), dtype=np.float32)
    gold = rng.integers(0, n_docs, size=n_q)
    q = (emb[gold] + 1.5 * rng.standard_normal((n_q, d), dtype=np.float32)).astype(np.float32)
    answers = [[f"ent{int(doc_ids[g]) % 977}"] for g in gold]         # several passages share an entity

    # "_encoding_corpus": batches of encoder outputs are added straight from the device
    store = ShardedCorpusStore(d, num_virtual_shards=3, device=0, seg_rows=4096)
    for part in np.array_split(emb, 7):
        store.add(torch.from_numpy(part).cuda(), shard=0)             # rank-0-style concatenation order
    store.finalize()
    m_all = {f"{m}@{t}": 0.0 for m in ("Recall", "MRR", "NDCG") for t in topk}
    o_all = dict(m_all)
    oracle_index = flat_ip.IndexFlatIP(d)
    oracle_index.add(emb)
    eval_num = 0
    for s in range(0, n_q, 32):                                       # eval_batch_size = 32
        qb = torch.from_numpy(q[s:s + 32]).cuda()
        _, indices = store.search_local_queries(qb, k)
        indices = indices.cpu().numpy()
        ids = [[int(doc_ids[i]) for i in row] for row in indices]
        docs = [[texts[i] for i in row] for row in ids]
        pos = hits_matrix(docs, answers[s:s + 32], doc_ids=ids, workers=1)
        for key, v in get_metrics(pos, topk).items():
            m_all[key] += v
        _, oi = oracle_index.search(q[s:s + 32], k)
        opos = hits_matrix([[texts[int(doc_ids[i])] for i in row] for row in oi], answers[s:s + 32], workers=1)
        for key, v in get_metrics(opos, topk).items():
            o_all[key] += v
        np.testing.assert_array_equal(indices, oi)                    # same ranked lists as the CPU path
        eval_num += len(indices)
    got = reduce_metrics(m_all, eval_num)
    want = reduce_metrics(o_all, eval_num)
    assert got == want and got["query_num"] == n_q
    assert got["Recall@5"] > 0.9                                      # the planted passage is found


def test_deferred_search_equals_per_step_searches():
    """Trainer.evaluate with eval_batch_size 16 (run.sh:30): queueing the encoder outputs of many
    steps and searching once returns, per step, exactly what the per-step searches return."""
    from denseretrievaltoolkits_b200 import faiss_compat
    from denseretrievaltoolkits_b200.deferred import DeferredSearch
    from denseretrievaltoolkits_b200.index import BaseFaissIPRetriever
    from denseretrievaltoolkits_b200.store import ShardedCorpusStore

    rng = np.random.default_rng(1)
    n, d, k = 60000, 768, 100
    emb = rng.standard_normal((n, d), dtype=np.float32)
    steps = [torch.from_numpy(rng.standard_normal((16, d), dtype=np.float32)).cuda() for _ in range(40)]
    steps.append(torch.from_numpy(rng.standard_normal((5, d), dtype=np.float32)).cuda())      # ragged last batch
    index = faiss_compat.IndexFlatIP(d, device=0, seg_rows=1 << 14)
    index.add(emb)
    store = ShardedCorpusStore(d, num_virtual_shards=2, device=0, seg_rows=1 << 14)
    store.add_split(torch.from_numpy(emb).cuda())
    store.finalize()
    retr = BaseFaissIPRetriever(emb)
    retr.add(emb)
    per_step = [index.search(q, k) for q in steps]
    for target in (index, store, retr):
        ds = DeferredSearch(target, k, max_queries=256)
        got = list(ds.results(enumerate(steps)))
        assert [t for t, _ in got] == list(range(len(steps)))
        assert ds.searches == 3                                        # 645 queries: 256 + 256 + tail
        for (_, (D, I)), (Dr, Ir) in zip(got, per_step):
            assert torch.equal(I, Ir) and torch.equal(D, Dr)
    # numpy batches through the host API
    ds = DeferredSearch(index, k, max_queries=100)
    got = list(ds.results((i, q.cpu().numpy()) for i, q in enumerate(steps[:10])))
    for (_, (D, I)), (Dr, Ir) in zip(got, per_step):
        np.testing.assert_array_equal(I, Ir.cpu().numpy())
        np.testing.assert_array_equal(D, Dr.cpu().numpy())
