"""CPU: pins the oracles (test infrastructure) to the golden vectors generated from the
reference's own importable code (tools/make_golden.py) and to each other."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import c_oracle, flat_ip, inbatch_loss, merge


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


@pytest.mark.parametrize("name", ["basic", "k_gt_n", "ties", "zeros", "nonfinite"])
def test_search_oracle_matches_reference_wrapper_goldens(golden_dir, name):
    g = _load(golden_dir, f"search_{name}.npz")
    x, q, k = g["x"], g["q"], int(g["k"])
    D, I = flat_ip.flat_ip_search(x, q, k)
    assert D.dtype == np.float32 and I.dtype == np.int64 and I.shape == (q.shape[0], k)
    np.testing.assert_array_equal(I, g["I"])
    np.testing.assert_array_equal(D, g["D"])
    idx = flat_ip.IndexFlatIP(x.shape[1])
    idx.add(x)
    # BaseFaissIPRetriever.search (index.py:31-33): ids only
    np.testing.assert_array_equal(flat_ip.wrapper_search_ids(idx, q, k), g["wrapper_ids"])


@pytest.mark.parametrize("name", ["basic", "k_gt_n", "ties", "zeros"])
def test_search_oracle_matches_float64_definition(golden_dir, name):
    g = _load(golden_dir, f"search_{name}.npz")
    D, I = flat_ip.flat_ip_search(g["x"], g["q"], int(g["k"]))
    valid = g["I64"] >= 0
    # fp32 vs float64 scores: within fp32 accumulation error
    np.testing.assert_allclose(D[valid], g["D64"][valid], rtol=1e-5, atol=1e-5)
    # ids agree except where float64 scores are closer than the fp32 rounding
    diff = (I != g["I64"]) & valid
    if diff.any():
        rows = np.nonzero(diff.any(axis=1))[0]
        for r in rows:
            cols = np.nonzero(diff[r])[0]
            gaps = np.abs(g["D64"][r, cols] - D[r, cols])
            assert (gaps < 1e-4).all()


def test_padding_and_nonfinite_semantics(golden_dir):
    g = _load(golden_dir, "search_k_gt_n.npz")
    D, I = flat_ip.flat_ip_search(g["x"], g["q"], int(g["k"]))
    n = g["x"].shape[0]
    assert (I[:, n:] == -1).all() and (D[:, n:] == np.float32(-3.4028234663852886e38)).all()
    assert (I[:, :n] >= 0).all()
    g = _load(golden_dir, "search_nonfinite.npz")
    D, I = flat_ip.flat_ip_search(g["x"], g["q"], int(g["k"]))
    assert 5 not in I and 7 not in I          # NaN / -inf scores never enter (faiss: thr < score)
    assert (I[:, 0] == 6).all() and np.isinf(D[:, 0]).all()   # +inf ranks first


@pytest.mark.parametrize("name", ["basic", "k_gt_n", "ties", "zeros", "nonfinite"])
def test_c_oracle_agrees_with_numpy_oracle(golden_dir, name):
    g = _load(golden_dir, f"search_{name}.npz")
    D, I = c_oracle.search(g["x"], g["q"], int(g["k"]))
    same = I == g["I"]
    # scalar accumulation order differs from sgemm: ids may swap only between near-equal scores
    assert same.mean() > 0.97
    fin = np.isfinite(g["D"]) & same
    np.testing.assert_allclose(D[fin], g["D"][fin], rtol=1e-4, atol=1e-5)
    if name in ("k_gt_n", "ties"):
        np.testing.assert_array_equal(I, g["I"])


def test_oracle_blocking_is_invariant():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3000, 32)).astype(np.float32)
    x[2000:2100] = x[:100]
    q = rng.standard_normal((9, 32)).astype(np.float32)
    a = flat_ip.IndexFlatIP(32, block_rows=257)
    b = flat_ip.IndexFlatIP(32, block_rows=100000)
    for part in np.array_split(x, 3):
        a.add(part)
    b.add(x)
    Da, Ia = a.search(q, 50)
    Db, Ib = b.search(q, 50)
    np.testing.assert_array_equal(Ia, Ib)
    np.testing.assert_allclose(Da, Db, rtol=1e-6)
    assert (np.diff(Da, axis=1) <= 0).all()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "loss_*.npz"))))
def test_loss_oracle_matches_reference_goldens(path):
    g = np.load(path, allow_pickle=False)
    B, n, d = int(g["B"]), int(g["n"]), int(g["d"])
    red = str(g["reduction"])
    if "x" in g:
        x, y = g["x"], g["y"]
    else:
        rng = np.random.default_rng(int(g["seed"]))
        x = (rng.standard_normal((B, d)) * 1.0).astype(np.float32)
        y = (rng.standard_normal((B * n, d)) * 1.0).astype(np.float32)
        if not (np.array_equal(x[:2], g["x_head"]) and np.array_equal(y[:2], g["y_head"])):
            pytest.skip("numpy RNG stream differs from the authoring container")
    target = g["target"] if "target" in g else None
    loss, lse, logits = inbatch_loss.contrastive_loss(x, y, target=target, reduction=red)
    np.testing.assert_allclose(np.asarray(loss, np.float64), g["loss"].astype(np.float64), rtol=2e-5, atol=1e-5)
    if red == "none":
        dx, dy = inbatch_loss.contrastive_loss_grads(x, y, target, "none", grad_out=np.ones(B))
    else:
        dx, dy = inbatch_loss.contrastive_loss_grads(x, y, target, red)
    if "dx" in g:
        np.testing.assert_allclose(dx, g["dx"], rtol=1e-3, atol=2e-5)
        np.testing.assert_allclose(dy, g["dy"], rtol=1e-3, atol=2e-5)
    else:
        np.testing.assert_allclose(dx[:8], g["dx_rows"], rtol=1e-3, atol=2e-5)
        np.testing.assert_allclose(dy[:16], g["dy_rows"], rtol=1e-3, atol=2e-5)
        np.testing.assert_allclose(dx.sum(0), g["dx_colsum"], rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(dy.sum(0), g["dy_colsum"], rtol=1e-3, atol=1e-4)


def test_default_target_rule():
    # losses.py:13-15 and biencoder.py:109-114: positives at columns 0, n, 2n, ...
    np.testing.assert_array_equal(inbatch_loss.default_target(4, 12), [0, 3, 6, 9])
    np.testing.assert_array_equal(inbatch_loss.default_target(3, 7), [0, 2, 4])


def test_distributed_loss_is_gathered_loss_times_world():
    rng = np.random.default_rng(3)
    xs = [rng.standard_normal((4, 16)).astype(np.float32) for _ in range(2)]
    ys = [rng.standard_normal((8, 16)).astype(np.float32) for _ in range(2)]
    full, _, _ = inbatch_loss.contrastive_loss(np.concatenate(xs), np.concatenate(ys))
    assert np.isclose(inbatch_loss.distributed_contrastive_loss(xs, ys), 2 * full)


def test_merge_oracle_matches_reference_goldens(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "merge_cases.json")))
    for c in cases:
        results, topk = c["results"], c["topk"]
        qids = list(results[0].keys())
        G, k_in = len(results), max(len(r[q]) for r in results for q in qids)
        scores = np.full((G, len(qids), k_in), np.float32(-3.4028234663852886e38), np.float32)
        ids = np.full((G, len(qids), k_in), -1, np.int64)
        for g, res in enumerate(results):
            for qi, q in enumerate(qids):
                for j, (doc, sc) in enumerate(res[q].items()):
                    ids[g, qi, j] = int(doc)
                    scores[g, qi, j] = sc
        D, I = merge.merge_topk(scores, ids, topk)
        for qi, q in enumerate(qids):
            want = c["merged"][q]
            got_ids = [int(i) for i in I[qi] if i >= 0]
            assert got_ids == [int(doc) for doc, _ in want]
            np.testing.assert_allclose(D[qi, : len(want)], [s for _, s in want], rtol=1e-6)


def test_mining_filter_matches_reference_loop(golden_dir):
    for c in json.load(open(os.path.join(golden_dir, "mining.json"))):
        out = merge.filter_negatives(np.array([c["ids"]], np.int64), np.array([c["b"]]), np.array([c["e"]]),
                                     c["num_negative"])
        kept = [int(v) for v in out[0] if v >= 0]
        assert kept == c["kept"]


def test_l2_oracle_and_factory_golden(golden_dir):
    """index_factory(d, "Flat") without a metric is faiss' IndexFlatL2; the golden holds what the
    reference's FaissRetriever (index.py:47-54) returned over this oracle."""
    g = np.load(os.path.join(golden_dir, "search_factory_flat.npz"))
    x, q, k = g["x"], g["q"], int(g["k"])
    assert int(g["metric"]) == flat_ip.METRIC_L2
    idx = flat_ip.index_factory(x.shape[1], "Flat")
    assert isinstance(idx, flat_ip.IndexFlatL2)
    idx.add(x)
    D, I = idx.search(q, k)
    np.testing.assert_array_equal(I, g["I"])
    np.testing.assert_allclose(D, g["D"], rtol=1e-6)
    # the literal definition, float64
    d2 = ((q[:, None, :].astype(np.float64) - x[None, :, :].astype(np.float64)) ** 2).sum(-1)
    np.testing.assert_array_equal(I, np.argsort(d2, axis=1, kind="stable")[:, :k])
    # the wrapper's argsort(-scores) re-order: farthest of the k first
    ids = np.array([ind[o] for ind, o in zip(I, np.argsort(-D, kind="stable"))])
    np.testing.assert_array_equal(ids, g["wrapper_ids"])
    # k > ntotal pads with (+FLT_MAX, -1); IP metric still available
    D2, I2 = flat_ip.flat_l2_search(x[:3], q[:2], 5)
    assert (I2[:, 3:] == -1).all() and (D2[:, 3:] == flat_ip.FLT_MAX).all()
    assert isinstance(flat_ip.index_factory(8, "Flat", flat_ip.METRIC_INNER_PRODUCT), flat_ip.IndexFlatIP)
    with pytest.raises(RuntimeError):
        flat_ip.index_factory(8, "IVF16,Flat")


def test_oracles_agree_with_independent_third_party_implementations():
    """faiss cannot be installed here, so the restatements are also checked against exact-search
    code that is neither ours nor the reference's: torch (blocked mm + topk, the CPU path the north
    star names when faiss is absent) for inner product, scikit-learn's brute-force
    NearestNeighbors for squared L2."""
    import torch
    from sklearn.neighbors import NearestNeighbors

    rng = np.random.default_rng(99)
    x = rng.standard_normal((5000, 96)).astype(np.float32)
    q = rng.standard_normal((40, 96)).astype(np.float32)
    k = 25
    D, I = flat_ip.flat_ip_search(x, q, k)
    Dt, It = flat_ip.torch_flat_ip_search(torch.from_numpy(x), torch.from_numpy(q), k, block_rows=1024)
    np.testing.assert_array_equal(I, It.numpy())              # continuous data: no ties
    np.testing.assert_allclose(D, Dt.numpy(), rtol=1e-5, atol=1e-5)
    D2, I2 = flat_ip.flat_l2_search(x, q, k)
    nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="sqeuclidean").fit(x.astype(np.float64))
    Ds, Is = nn.kneighbors(q.astype(np.float64))
    np.testing.assert_array_equal(I2, Is)
    np.testing.assert_allclose(D2, Ds, rtol=1e-5)


def test_oracle_matches_real_faiss_when_importable():
    """"if `import faiss` succeeds at run time, run real faiss as the primary oracle".
    faiss is not installable in the authoring container nor present on the GPU image (no network),
    so this pin is skipped there; wherever a real faiss exists it checks the restatement's scores,
    ids, tie order and k > ntotal padding against `faiss.IndexFlatIP` itself."""
    faiss = pytest.importorskip("faiss")
    if not hasattr(faiss, "omp_get_max_threads"):
        pytest.skip("`faiss` on sys.modules is this repo's compat module, not the real library")
    rng = np.random.default_rng(2)
    x = rng.standard_normal((3000, 64)).astype(np.float32)
    x[2900:2950] = x[100:150]                       # exact ties
    q = rng.standard_normal((17, 64)).astype(np.float32)
    for k in (10, 100, 4000):                       # 4000 > ntotal: (-FLT_MAX, -1) padding
        index = faiss.IndexFlatIP(64)
        index.add(x)
        Df, If = index.search(q, k)
        Do, Io = flat_ip.flat_ip_search(x, q, k)
        np.testing.assert_allclose(Do, Df, rtol=1e-5, atol=1e-5)
        assert (Io == If).mean() > 0.98             # tie order inside faiss is implementation defined
        assert ((Io < 0) == (If < 0)).all()
