"""GPU (B200) parity tests of the search path, all through the C ABI (ctypes -> libdrt_b200.so).

Bar (BASELINE.json north_star): scores within 1e-4 relative of the reference arithmetic, top-k
id lists identical except where scores tie within that tolerance, recall@k >= 0.999.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import flat_ip  # noqa: E402

RTOL = 1e-4          # north_star: "scores within 1e-4 relative"


def _mk(device=0, d=768, seg_rows=0):
    from denseretrievaltoolkits_b200 import faiss_compat

    return faiss_compat.IndexFlatIP(d, device=device, seg_rows=seg_rows)


def _check_parity(D, I, Dr, Ir, k, n, scale):
    """ids identical except among scores that tie within tolerance; scores within RTOL."""
    kk = min(k, n)
    recall = np.mean([len(set(a[:kk]) & set(b[:kk])) / kk for a, b in zip(I, Ir)])
    assert recall >= 0.999, recall
    valid = Ir >= 0
    np.testing.assert_array_equal(I >= 0, valid)
    # asserted at what the kernels deliver, not at the 1e-4 bar: 1e-5 relative plus the fp32
    # rounding of two summation orders, 1e-6 |q||d| (scale = typical |q| = |d|)
    rtol, atol = 1e-5, 1e-6 * scale * scale
    np.testing.assert_allclose(D[valid], Dr[valid], rtol=rtol, atol=atol)
    diff = (I != Ir) & valid
    # a differing id must sit in a near-tie: the oracle's score at that rank equals ours within tol
    assert np.all(np.abs(D[diff] - Dr[diff]) <= rtol * np.abs(Dr[diff]) + atol)
    assert (np.diff(D, axis=1) <= 0).all()
    assert ((I[:, kk:] == -1).all() and (D[:, kk:] == np.float32(-3.4028234663852886e38)).all())
    return recall


@pytest.mark.parametrize("ctas", [1, 2])
@pytest.mark.parametrize("nq,n,k,seg_rows", [(7, 1000, 10, 256), (128, 100000, 100, 1 << 15), (33, 5000, 1000, 2048),
                                             (3, 5, 8, 256), (65, 777, 50, 256), (300, 40001, 200, 8192)])
def test_search_matches_oracle(ctas, nq, n, k, seg_rows):
    from denseretrievaltoolkits_b200 import _lib

    rng = np.random.default_rng(nq * 1000 + n)
    x = rng.standard_normal((n, 768), dtype=np.float32)
    q = rng.standard_normal((nq, 768), dtype=np.float32)
    index = _mk(seg_rows=seg_rows)
    for part in np.array_split(x, 3):          # several add() calls, ids = insertion order
        index.add(part)
    assert index.ntotal == n
    flags = _lib.SEARCH_FORCE_2CTA if ctas == 2 else _lib.SEARCH_FORCE_1CTA
    D, I = index.search(q, k, flags=flags)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (nq, k)
    Dr, Ir = flat_ip.flat_ip_search(x, q, k)
    _check_parity(D, I, Dr, Ir, k, n, scale=np.sqrt(768.0))
    st = index.search_stats()
    assert st["ctas_per_tile"] == ctas and st["filter_launches"] >= 1 and st["overflow_retries"] == 0


@pytest.mark.parametrize("name", ["basic", "k_gt_n", "ties", "zeros", "nonfinite"])
def test_search_matches_reference_wrapper_goldens(golden_dir, name):
    """The mirror of BaseFaissIPRetriever returns the ids the reference wrapper returned."""
    from denseretrievaltoolkits_b200.index import BaseFaissIPRetriever

    g = np.load(os.path.join(golden_dir, f"search_{name}.npz"))
    x, q, k = g["x"], g["q"], int(g["k"])
    r = BaseFaissIPRetriever(x)
    assert r.index.ntotal == 0                   # ctor does not add (index.py:18-19)
    r.add(x)
    ids = r.search(q, k)
    D, I = r.search_with_scores(q, k)
    assert ids.dtype == np.int64 and ids.shape == (q.shape[0], k)
    np.testing.assert_array_equal(ids, I)
    if name == "nonfinite":
        # +inf scores first, NaN / -inf rows never returned; bf16 first pass keeps the finite order
        assert (I[:, 0] == 6).all() and 5 not in I and 7 not in I
        fin = np.isfinite(g["D"])
        same = (I == g["I"]) & fin
        assert same[fin].all()
        np.testing.assert_allclose(D[same], g["D"][same], rtol=1e-5, atol=1e-5)
        return
    _check_parity(D, I, g["D"], g["I"], k, x.shape[0], scale=np.sqrt(64.0))
    if name in ("ties", "k_gt_n", "zeros"):
        np.testing.assert_array_equal(ids, g["wrapper_ids"])     # exact ties: (score desc, id asc)
    assert r.batch_search(q, k, 2, quiet=True).shape == ids.shape


def test_duplicate_rows_tie_break_by_id():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((30000, 768), dtype=np.float32)
    x[-300:] = x[:300]                            # DistributedSampler-style duplicated rows
    q = rng.standard_normal((64, 768), dtype=np.float32)
    index = _mk(seg_rows=4096)
    index.add(x)
    D, I = index.search(q, 100)
    Dr, Ir = flat_ip.flat_ip_search(x, q, 100)
    _check_parity(D, I, Dr, Ir, 100, 30000, scale=np.sqrt(768.0))
    for r in range(64):                           # equal scores => ascending ids
        eq = D[r, 1:] == D[r, :-1]
        assert (I[r, 1:][eq] > I[r, :-1][eq]).all()


def test_device_tensor_api_equals_host_api():
    rng = np.random.default_rng(6)
    x = rng.standard_normal((20000, 768), dtype=np.float32)
    q = rng.standard_normal((50, 768), dtype=np.float32)
    a, b = _mk(seg_rows=4096), _mk(seg_rows=8192)
    a.add(x)
    b.add(torch.from_numpy(x).cuda())            # zero-copy ingest of device rows
    Da, Ia = a.search(q, 64)
    Db, Ib = b.search(torch.from_numpy(q).cuda(), 64)
    assert Db.is_cuda and Ib.dtype == torch.int64
    np.testing.assert_array_equal(Ia, Ib.cpu().numpy())
    np.testing.assert_array_equal(Da, Db.cpu().numpy())
    np.testing.assert_array_equal(a.reconstruct_n(100, 50), x[100:150])
    # idempotence: same call, same bits
    Da2, Ia2 = a.search(q, 64)
    np.testing.assert_array_equal(Ia, Ia2)
    np.testing.assert_array_equal(Da, Da2)


def test_adversarial_row_order_triggers_retry_and_stays_exact():
    """Scores increase with the row id, so every row beats the running threshold: the candidate
    buffer overflows, the search is redone with safer chunking, results stay exact."""
    n = 60000
    rng = np.random.default_rng(7)
    u = rng.standard_normal(768).astype(np.float32)
    u /= np.linalg.norm(u)
    x = (np.linspace(0.0, 40.0, n, dtype=np.float32)[:, None] * u[None, :]
         + 0.01 * rng.standard_normal((n, 768), dtype=np.float32)).astype(np.float32)
    q = (u[None, :] * np.linspace(1.0, 2.0, 9, dtype=np.float32)[:, None]).astype(np.float32)
    index = _mk(seg_rows=1 << 15)
    index.add(x)
    D, I = index.search(q, 100)
    assert index.search_stats()["overflow_retries"] >= 1
    Dr, Ir = flat_ip.flat_ip_search(x, q, 100)
    _check_parity(D, I, Dr, Ir, 100, n, scale=40.0)


def test_empty_index_and_errors():
    index = _mk()
    D, I = index.search(np.zeros((3, 768), np.float32), 5)
    assert (I == -1).all() and (D == np.float32(-3.4028234663852886e38)).all()
    with pytest.raises(RuntimeError):
        index.search(np.zeros((3, 100), np.float32), 5)
    with pytest.raises(RuntimeError):
        index.search(np.zeros((3, 768), np.float32), 5000)       # k > DRT_MAX_K
    from denseretrievaltoolkits_b200 import faiss_compat

    with pytest.raises(RuntimeError):
        faiss_compat.index_factory(768, "IVF100,PQ8")             # approximate: refused, no fallback
    # faiss' signature: index_factory(d, description, metric=METRIC_L2)
    assert isinstance(faiss_compat.index_factory(768, "Flat"), faiss_compat.IndexFlatL2)
    assert isinstance(faiss_compat.index_factory(768, "Flat", faiss_compat.METRIC_INNER_PRODUCT), faiss_compat.IndexFlatIP)


def _check_l2(D, I, Dr, Ir, x, q):
    """ids identical except among distances that tie within tolerance; distances within RTOL of
    the float64 oracle (tolerance scaled by the cancelling terms |q|^2 + |x|^2)."""
    valid = Ir >= 0
    np.testing.assert_array_equal(I >= 0, valid)
    atol = RTOL * (float((q.astype(np.float64) ** 2).sum(1).max()) + float((x.astype(np.float64) ** 2).sum(1).max()))
    np.testing.assert_allclose(D[valid], Dr[valid], rtol=RTOL, atol=atol)
    diff = (I != Ir) & valid
    assert np.all(np.abs(D[diff] - Dr[diff]) <= RTOL * np.abs(Dr[diff]) + atol)
    assert (np.diff(D, axis=1) >= 0).all()                       # ascending distances
    assert (D[~valid] == np.float32(3.4028234663852886e38)).all()
    k = I.shape[1]
    kk = min(k, x.shape[0])
    assert np.mean([len(set(a[:kk]) & set(b[:kk])) / kk for a, b in zip(I, Ir)]) >= 0.999


@pytest.mark.parametrize("n,d,k", [(20000, 768, 100), (3000, 100, 10), (5, 64, 8)])
def test_index_flat_l2_matches_oracle(n, d, k, tmp_path):
    """`faiss.index_factory(d, "Flat")` = IndexFlatL2 (what FaissRetriever holds, index.py:50):
    exact squared-L2 search on the inner-product store through the norm-augmented rows."""
    from denseretrievaltoolkits_b200 import faiss_compat

    rng = np.random.default_rng(n + d)
    x = (rng.standard_normal((n, d)) * rng.uniform(0.5, 1.5, size=(n, 1))).astype(np.float32)   # varied norms
    q = rng.standard_normal((60, d), dtype=np.float32)
    index = faiss_compat.index_factory(d, "Flat")
    index.add(x[: n // 2])
    index.add(torch.from_numpy(x[n // 2:]).cuda())
    assert index.ntotal == n and index.d == d and index.metric_type == faiss_compat.METRIC_L2
    D, I = index.search(q, k)
    Dr, Ir = flat_ip.flat_l2_search(x, q, k)
    _check_l2(D, I, Dr, Ir, x, q)
    Dd, Id = index.search(torch.from_numpy(q).cuda(), k)
    np.testing.assert_array_equal(Id.cpu().numpy(), I)
    np.testing.assert_allclose(Dd.cpu().numpy(), D, rtol=1e-6, atol=1e-3)
    np.testing.assert_array_equal(index.reconstruct_n(0, min(n, 7)), x[: min(n, 7)])
    path = str(tmp_path / "l2.faiss")
    faiss_compat.write_index(index, path)
    again = faiss_compat.read_index(path)
    assert isinstance(again, faiss_compat.IndexFlatL2)
    D2, I2 = again.search(q, k)
    np.testing.assert_array_equal(I2, I)


def test_faiss_retriever_flat_matches_reference_wrapper_golden(golden_dir):
    """The reference's FaissRetriever, run unmodified over the oracle stub, pinned the ids it
    returns for index_factory(d, "Flat"): L2 neighbours, re-ordered by argsort(-distance)."""
    from denseretrievaltoolkits_b200.index import FaissRetriever

    g = np.load(os.path.join(golden_dir, "search_factory_flat.npz"))
    r = FaissRetriever(g["x"], "Flat")
    assert r.index.ntotal == 0 and r.index.verbose is True
    r.add(g["x"])
    ids = r.search(g["q"], int(g["k"]))
    np.testing.assert_array_equal(ids, g["wrapper_ids"])
    D, I = r.index.search(g["q"], int(g["k"]))
    np.testing.assert_array_equal(I, g["I"])
    np.testing.assert_allclose(D, g["D"], rtol=RTOL, atol=1e-3)


def test_write_read_index_roundtrip(tmp_path):
    from denseretrievaltoolkits_b200 import faiss_compat

    rng = np.random.default_rng(8)
    x = rng.standard_normal((3000, 128), dtype=np.float32)
    q = rng.standard_normal((11, 128), dtype=np.float32)
    a = _mk(d=128, seg_rows=1024)
    a.add(x)
    faiss_compat.write_index(a, str(tmp_path / "idx"))
    b = faiss_compat.read_index(str(tmp_path / "idx"))
    assert b.ntotal == 3000 and b.d == 128
    Da, Ia = a.search(q, 30)
    Db, Ib = b.search(q, 30)
    np.testing.assert_array_equal(Ia, Ib)
    np.testing.assert_array_equal(Da, Db)


def test_reference_index_module_runs_unmodified_on_faiss_shim(golden_dir):
    """sys.modules['faiss'] = faiss_compat: code written against faiss (the reference's
    index.py does `faiss.IndexFlatIP(d)`, `.add`, `.search`) works as is."""
    import sys

    import denseretrievaltoolkits_b200 as pkg

    saved = sys.modules.get("faiss")
    try:
        pkg.install_as_faiss()
        import faiss

        g = np.load(os.path.join(golden_dir, "search_basic.npz"))
        idx = faiss.IndexFlatIP(g["x"].shape[1])
        idx.add(g["x"])
        scores, indices = idx.search(g["q"], int(g["k"]))
        ids = np.array([ind[o] for ind, o in zip(indices, np.argsort(-scores))])   # index.py:33
        np.testing.assert_array_equal(ids, g["wrapper_ids"])
    finally:
        if saved is not None:
            sys.modules["faiss"] = saved
        else:
            sys.modules.pop("faiss", None)


def test_retrieval_cli_on_gpu(tmp_path):
    import pickle

    from denseretrievaltoolkits_b200 import retrieval

    rng = np.random.default_rng(9)
    x = rng.standard_normal((5000, 768), dtype=np.float32)
    q = rng.standard_normal((20, 768), dtype=np.float32)
    lookups = []
    for i, part in enumerate(np.array_split(x, 4)):
        lk = [f"d{i}_{j}" for j in range(part.shape[0])]
        lookups += lk
        pickle.dump((part, lk), open(tmp_path / f"c{i}.pkl", "wb"))
    pickle.dump((q, list(range(20))), open(tmp_path / "q.pkl", "wb"))
    scores, psg = retrieval.main(["--query_reps", str(tmp_path / "q.pkl"), "--passage_reps", str(tmp_path / "c*.pkl"),
                                  "--depth", "100", "--batch_size", "8", "--save_ranking_to", str(tmp_path / "r.tsv"),
                                  "--save_text", "--quiet"])
    Dr, Ir = flat_ip.flat_ip_search(x, q, 100)
    want = np.array([[lookups[i] for i in row] for row in Ir], dtype=object)
    assert (psg == want).mean() > 0.999
    assert sum(1 for _ in open(tmp_path / "r.tsv")) == 2000


def test_virtual_shards_equal_single_store():
    """G-shard search + merge kernel == 1-shard search, ids bit for bit."""
    from denseretrievaltoolkits_b200.store import ShardedCorpusStore

    rng = np.random.default_rng(10)
    x = rng.standard_normal((50000, 768), dtype=np.float32)
    x[45000:45100] = x[100:200]                   # ties across shards
    q = rng.standard_normal((100, 768), dtype=np.float32)
    single = _mk(seg_rows=8192)
    single.add(x)
    Ds, Is = single.search(q, 100)
    for G in (2, 4, 8):
        st = ShardedCorpusStore(768, num_virtual_shards=G, device=0, seg_rows=4096)
        st.add_split(x)
        assert st.ntotal == 50000
        D, I = st.search(q, 100)
        np.testing.assert_array_equal(I, Is)
        np.testing.assert_array_equal(D, Ds)
        assert st.last_search["local_depth"] == {2: 88, 4: 56, 8: 40}[G]      # reduced per-shard depth


def test_reduced_shard_depth_requeries_when_rows_correlate_with_queries():
    """Shards are searched to mean + 6 sigma of their expected share of the top-k; when one shard
    holds (nearly) the whole top-k the truncation check fires, those queries are searched again
    at full depth, the result is still bit-identical to one store, and the store stops reducing."""
    from denseretrievaltoolkits_b200.store import ShardedCorpusStore

    rng = np.random.default_rng(12)
    x = rng.standard_normal((40000, 256), dtype=np.float32)
    x[:5000] *= 3.0                                  # shard 0 of 8 owns every query's top-k
    q = rng.standard_normal((64, 256), dtype=np.float32)
    single = _mk(d=256, seg_rows=4096)
    single.add(x)
    Ds, Is = single.search(q, 100)
    st = ShardedCorpusStore(256, num_virtual_shards=8, device=0, seg_rows=4096)
    st.add_split(x)
    D, I = st.search(q, 100)
    np.testing.assert_array_equal(I, Is)
    np.testing.assert_array_equal(D, Ds)
    assert st.last_search == {"local_depth": 40, "requeried": 64, "redone": 0}
    D, I = st.search(torch.from_numpy(q).cuda(), 100)          # now at full depth, device tensors
    assert st.last_search == {"local_depth": 100, "requeried": 0, "redone": 0}
    np.testing.assert_array_equal(I.cpu().numpy(), Is)
    # a few queries only: mixed result rows
    x2 = rng.standard_normal((40000, 256), dtype=np.float32)
    q2 = rng.standard_normal((64, 256), dtype=np.float32)
    x2[:60] = 6.0 * q2[5] / np.linalg.norm(q2[5]) + 0.01 * x2[:60]    # 60 of query 5's top-100 sit in shard 0
    single2 = _mk(d=256, seg_rows=4096)
    single2.add(x2)
    Ds2, Is2 = single2.search(q2, 100)
    st2 = ShardedCorpusStore(256, num_virtual_shards=8, device=0, seg_rows=4096)
    st2.add_split(x2)
    D2, I2 = st2.search(q2, 100)
    np.testing.assert_array_equal(I2, Is2)
    np.testing.assert_array_equal(D2, Ds2)
    assert 1 <= st2.last_search["requeried"] <= 8 and st2._reduce_depth


@pytest.mark.parametrize("k", [100, 1000])
def test_scale_properties_2m_rows(k):
    """Size-independent properties at a multi-segment scale (2.1M x 768): agreement with an
    independent GPU fp32 reference (torch.matmul, TF32 off, + topk) on a query subset,
    sortedness, score == exact dot of the returned row, idempotence."""
    import bench

    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda", 0)
    n = 2 * (1 << 20) + 12345
    index = _mk()
    bench.fill_rows(torch, index.add, 0, n, dev)
    nq = 512
    q = bench.make_queries(torch, nq, dev)
    D, I = index.search(q, k)
    assert index.search_stats()["overflow_retries"] == 0
    assert (D[:, 1:] <= D[:, :-1]).all() and (I >= 0).all() and (I < n).all()
    sub = slice(0, 64)
    best_d = torch.full((64, 0), 0.0, device=dev)
    best_i = torch.zeros((64, 0), dtype=torch.int64, device=dev)
    for c in range(0, n, bench.CHUNK):
        rows = bench.make_corpus_chunk(torch, c // bench.CHUNK, dev)[: min(bench.CHUNK, n - c)]
        s = q[sub] @ rows.t()
        d, i = torch.topk(s, k, dim=1)
        best_d = torch.cat([best_d, d], 1)
        best_i = torch.cat([best_i, i + c], 1)
        d2, sel = torch.topk(best_d, k, dim=1)
        best_d, best_i = d2, torch.gather(best_i, 1, sel)
        del rows, s
    ref_sets = [set(r.tolist()) for r in best_i]
    got = I[sub].cpu()
    recall = np.mean([len(ref_sets[r] & set(got[r].tolist())) / k for r in range(64)])
    assert recall >= 0.999, recall
    torch.testing.assert_close(D[sub], best_d, rtol=RTOL, atol=1e-3)
    D2, I2 = index.search(q, k)
    assert torch.equal(I, I2) and torch.equal(D, D2)


@pytest.mark.parametrize("d", [64, 128, 384, 1024, 4096, 8192])
def test_other_embedding_dims(d):
    """Multiples of 64 (one 128-byte bf16 swizzle span) are stored as they are: MiniLM 384, BERT-large 1024..."""
    rng = np.random.default_rng(d)
    n = 30000 if d <= 1024 else 6000
    x = rng.standard_normal((n, d), dtype=np.float32)
    q = rng.standard_normal((150, d), dtype=np.float32)
    index = _mk(d=d, seg_rows=8192 if d <= 1024 else 2048)
    index.add(x)
    D, I = index.search(q, 50)
    Dr, Ir = flat_ip.flat_ip_search(x, q, 50)
    _check_parity(D, I, Dr, Ir, 50, n, scale=np.sqrt(float(d)))


@pytest.mark.parametrize("d", [1, 50, 100, 300, 770])
def test_dims_that_are_not_multiples_of_64(d, tmp_path):
    """`projection_out_dim` (arguments.py:46) is free in the reference: such rows are stored
    zero-padded to the next multiple of 64, invisible through add / search / reconstruct / files."""
    from denseretrievaltoolkits_b200 import faiss_compat

    rng = np.random.default_rng(1000 + d)
    x = rng.standard_normal((9000, d), dtype=np.float32)
    q = rng.standard_normal((70, d), dtype=np.float32)
    index = _mk(d=d, seg_rows=4096)
    index.add(x[:5000])
    index.add(torch.from_numpy(x[5000:]).cuda())            # device rows, unpadded pitch
    assert index.d == d and index.ntotal == 9000
    D, I = index.search(q, 30)
    Dr, Ir = flat_ip.flat_ip_search(x, q, 30)
    _check_parity(D, I, Dr, Ir, 30, 9000, scale=np.sqrt(float(d)))
    Dd, Id = index.search(torch.from_numpy(q).cuda(), 30)   # device queries, unpadded pitch
    np.testing.assert_array_equal(Id.cpu().numpy(), I)
    np.testing.assert_array_equal(index.reconstruct_n(4090, 20), x[4090:4110])   # across a segment boundary
    path = str(tmp_path / "idx.faiss")
    faiss_compat.write_index(index, path)
    again = faiss_compat.read_index(path)
    assert again.d == d
    D2, I2 = again.search(q, 30)
    np.testing.assert_array_equal(I2, I)
    np.testing.assert_array_equal(D2, D)


def test_near_duplicate_corpus_falls_back_to_exact_fp32_pass():
    """Score gaps far below bf16 resolution: the margin check cannot certify the bf16 pass, so
    the flagged queries go through larger k' and finally the exact fp32 first pass.  The data are
    small integers, so fp32 (and float64) inner products are exact in any summation order and
    the result must equal the float64 oracle id for id, ties broken by ascending id."""
    n, d = 40000, 768
    rng = np.random.default_rng(11)
    c = rng.integers(256, 768, size=d)
    x = (c[None, :] + (rng.random((n, d)) < 0.05)).astype(np.float32)      # rows differ by sparse +1s
    q = rng.integers(0, 3, size=(6, d)).astype(np.float32)
    index = _mk(seg_rows=1 << 14)
    index.add(x)
    D, I = index.search(q, 50)
    st = index.search_stats()
    assert st["exact_queries"] == 6 and st["flagged_queries"] == 0, st
    Dr, Ir = flat_ip.flat_ip_search_f64(x, q, 50)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D.astype(np.float64), Dr)


def test_randomised_shapes_against_oracle():
    """Seeded fuzz over (nq, n, k, seg_rows, dim, #adds): tile / segment / chunk boundaries,
    k > n, single rows, CTA-variant switch at 128 queries."""
    rng = np.random.default_rng(2024)
    cases = [(1, 1, 1, 256, 64), (129, 257, 3, 256, 64), (128, 256, 256, 256, 64), (127, 65537, 100, 16384, 64)]
    for _ in range(26):
        d = int(rng.choice([64, 128]))
        cases.append((int(rng.integers(1, 600)), int(rng.integers(1, 70000)), int(rng.integers(1, 300)),
                      int(rng.choice([256, 1024, 4096, 16384])), d))
    for nq, n, k, seg_rows, d in cases:
        x = rng.standard_normal((n, d), dtype=np.float32)
        if n > 50 and rng.random() < 0.3:
            x[-(n // 10):] = x[: n // 10]             # exact duplicates -> ties
        q = rng.standard_normal((nq, d), dtype=np.float32)
        index = _mk(d=d, seg_rows=seg_rows)
        for part in np.array_split(x, int(rng.integers(1, 5))):
            if part.shape[0]:
                index.add(part)
        D, I = index.search(q, k)
        Dr, Ir = flat_ip.flat_ip_search(x, q, k)
        try:
            _check_parity(D, I, Dr, Ir, k, n, scale=np.sqrt(float(d)))
        except AssertionError as e:
            raise AssertionError(f"case nq={nq} n={n} k={k} seg_rows={seg_rows} d={d}: {e}") from e


def _aniso(rng, n, d=768):
    """SURVEY §8d 'aniso' (precision stress, BERT-like): x = mu + s * z, |mu| = 8, per-dimension
    scale log-uniform in [0.25, 4]."""
    mu = np.random.default_rng(7).standard_normal(d).astype(np.float32)
    mu *= 8.0 / np.linalg.norm(mu)
    s = np.exp(np.random.default_rng(8).uniform(np.log(0.25), np.log(4.0), size=d)).astype(np.float32)
    return (mu[None, :] + s[None, :] * rng.standard_normal((n, d), dtype=np.float32)).astype(np.float32)


@pytest.mark.parametrize("k", [10, 100, 500])
def test_anisotropic_embeddings_with_common_offset(k):
    """BERT-like geometry: a large common mean component and a 16x spread of per-dimension
    scales make the bf16 error large relative to the score gaps; results must stay exact."""
    rng = np.random.default_rng(31 + k)
    x = _aniso(rng, 120000)
    q = _aniso(rng, 256)
    index = _mk(seg_rows=1 << 15)
    index.add(x)
    D, I = index.search(q, k)
    st = index.search_stats()
    assert st["flagged_queries"] == 0
    Dr, Ir = flat_ip.flat_ip_search(x, q, k)
    scale = float(np.abs(Dr).max())
    _check_parity(D, I, Dr, Ir, k, 120000, scale=scale * 1e-2)


def test_unit_norm_small_magnitude_embeddings():
    """Cosine-style retrieval: rows normalised to unit length (elements ~0.036)."""
    rng = np.random.default_rng(77)
    x = rng.standard_normal((80000, 768), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = rng.standard_normal((200, 768), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    index = _mk(seg_rows=1 << 15)
    index.add(x)
    D, I = index.search(q, 100)
    Dr, Ir = flat_ip.flat_ip_search(x, q, 100)
    _check_parity(D, I, Dr, Ir, 100, 80000, scale=1e-2)


def test_query_count_above_internal_batch():
    """More than 16,384 queries in one call are processed in internal batches (bounded workspace);
    device and host paths must agree with each other and with the oracle."""
    rng = np.random.default_rng(55)
    x = rng.standard_normal((6000, 128), dtype=np.float32)
    q = rng.standard_normal((20001, 128), dtype=np.float32)
    index = _mk(d=128, seg_rows=2048)
    index.add(x)
    D, I = index.search(q, 20)
    Dd, Id = index.search(torch.from_numpy(q).cuda(), 20)
    np.testing.assert_array_equal(I, Id.cpu().numpy())
    np.testing.assert_array_equal(D, Dd.cpu().numpy())
    Dr, Ir = flat_ip.flat_ip_search(x, q[16000:16800], 20)       # rows straddling the batch boundary
    _check_parity(D[16000:16800], I[16000:16800], Dr, Ir, 20, 6000, scale=np.sqrt(128.0))


def test_host_threads_sharing_one_store():
    """Two host threads searching the same store at once (ctypes drops the GIL): calls are
    serialised inside the library, so both get the results a sequential run gives."""
    import threading

    rng = np.random.default_rng(91)
    x = rng.standard_normal((30000, 256), dtype=np.float32)
    qs = [rng.standard_normal((257, 256), dtype=np.float32) for _ in range(2)]
    index = _mk(d=256, seg_rows=4096)
    index.add(x)
    want = [index.search(q, 50) for q in qs]
    got = [None, None]
    errs = []

    def work(i):
        try:
            for _ in range(20):
                got[i] = index.search(qs[i], 50)
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for (D, I), (Dw, Iw) in zip(got, want):
        np.testing.assert_array_equal(I, Iw)
        np.testing.assert_array_equal(D, Dw)


def test_single_query_single_row_and_k1():
    """Degenerate shapes: one query, one row, k = 1 (the sampler's retriever.search(query, n) shape)."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((1, 768), dtype=np.float32)
    q = rng.standard_normal((1, 768), dtype=np.float32)
    index = _mk(seg_rows=256)
    index.add(x)
    D, I = index.search(q, 1)
    assert I.tolist() == [[0]] and abs(D[0, 0] - float(q[0] @ x[0])) <= 1e-4 * abs(D[0, 0]) + 1e-4
    D, I = index.search(q, 3)
    assert I.tolist() == [[0, -1, -1]]
    x2 = rng.standard_normal((1000, 768), dtype=np.float32)
    index.add(x2)
    D, I = index.search(q, 1)
    Dr, Ir = flat_ip.flat_ip_search(np.concatenate([x, x2]), q, 1)
    np.testing.assert_array_equal(I, Ir)


@pytest.mark.parametrize("nq", [5, 300])
def test_maximum_depth_k2048(nq):
    """k = DRT_MAX_K: k' = 2,344 candidates per query in the 16,384-entry buffer, both tile variants."""
    rng = np.random.default_rng(2048 + nq)
    x = rng.standard_normal((60000, 256), dtype=np.float32)
    q = rng.standard_normal((nq, 256), dtype=np.float32)
    index = _mk(d=256, seg_rows=1 << 14)
    index.add(x)
    D, I = index.search(q, 2048)
    st = index.search_stats()
    assert st["overflow_retries"] == 0 and st["flagged_queries"] == 0
    Dr, Ir = flat_ip.flat_ip_search(x, q[:40], 2048)
    _check_parity(D[:40], I[:40], Dr, Ir, 2048, 60000, scale=16.0)
    assert (np.diff(D, axis=1) <= 0).all() and all(len(set(r)) == 2048 for r in I)


def test_device_queries_with_4_byte_alignment():
    """A device query pointer that is only 4-byte aligned (a view at an odd element offset) is
    staged inside the library instead of being read with 16-byte vector loads."""
    rng = np.random.default_rng(8)
    x = rng.standard_normal((20000, 128), dtype=np.float32)
    q = rng.standard_normal((50, 128), dtype=np.float32)
    index = _mk(d=128, seg_rows=4096)
    index.add(x)
    flat = torch.zeros(50 * 128 + 1, device="cuda")
    flat[1:] = torch.from_numpy(q).cuda().reshape(-1)
    qv = flat[1:].view(50, 128)
    assert qv.data_ptr() % 16 == 4 and qv.is_contiguous()
    D, I = index.search(qv, 20)
    Dh, Ih = index.search(q, 20)
    np.testing.assert_array_equal(I.cpu().numpy(), Ih)
    np.testing.assert_array_equal(D.cpu().numpy(), Dh)


def test_sharded_store_save_and_load(tmp_path):
    from denseretrievaltoolkits_b200.store import ShardedCorpusStore

    rng = np.random.default_rng(14)
    x = rng.standard_normal((30000, 128), dtype=np.float32)
    q = rng.standard_normal((40, 128), dtype=np.float32)
    st = ShardedCorpusStore(128, num_virtual_shards=4, device=0, seg_rows=4096)
    st.add_split(x)
    D, I = st.search(q, 50)
    st.save(str(tmp_path / "store"))
    again = ShardedCorpusStore.load(str(tmp_path / "store"), num_virtual_shards=4, device=0, seg_rows=4096)
    assert again.ntotal == 30000 and again._offsets == st._offsets
    D2, I2 = again.search(q, 50)
    np.testing.assert_array_equal(I2, I)
    np.testing.assert_array_equal(D2, D)
    with pytest.raises(RuntimeError):
        ShardedCorpusStore.load(str(tmp_path / "store"), num_virtual_shards=2, device=0)
