"""GPU (B200): BASELINE.json's full corpus size (8.8M x 768, the cfg2 / headline shape) checked
through size-independent properties and an independent GPU fp32 reference on a query subset
(SURVEY.md §8d "Parity at scale")."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full_store():
    import bench
    from denseretrievaltoolkits_b200 import faiss_compat

    free, _ = torch.cuda.mem_get_info(0)
    if free < 70e9:
        pytest.skip("needs ~50 GB of free HBM")
    dev = torch.device("cuda", 0)
    index = faiss_compat.IndexFlatIP(bench.DIM, device=0)
    bench.fill_rows(torch, index.add, 0, bench.HEADLINE["n"], dev)
    torch.cuda.synchronize()
    return index


def _torch_reference(q, n, k):
    import bench

    torch.backends.cuda.matmul.allow_tf32 = False
    dev = q.device
    best_d = torch.full((q.shape[0], 0), 0.0, device=dev)
    best_i = torch.zeros((q.shape[0], 0), dtype=torch.int64, device=dev)
    for c in range(0, n, bench.CHUNK):
        rows = bench.make_corpus_chunk(torch, c // bench.CHUNK, dev)[: min(bench.CHUNK, n - c)]
        d, i = torch.topk(q @ rows.t(), k, dim=1)
        best_d, best_i = torch.cat([best_d, d], 1), torch.cat([best_i, i + c], 1)
        d2, sel = torch.topk(best_d, k, dim=1)
        best_d, best_i = d2, torch.gather(best_i, 1, sel)
        del rows
    return best_d, best_i


@pytest.mark.parametrize("k,nq", [(100, 1024), (1000, 512), (200, 700)])
def test_full_corpus_parity_and_properties(full_store, k, nq):
    import bench

    index = full_store
    n = bench.HEADLINE["n"]
    assert index.ntotal == n
    q = bench.make_queries(torch, nq, torch.device("cuda", 0))
    D, I = index.search(q, k)
    st = index.search_stats()
    assert st["overflow_retries"] == 0 and st["flagged_queries"] == 0
    assert (D[:, 1:] <= D[:, :-1]).all() and (I >= 0).all() and (I < n).all()
    # no duplicate ids within a row
    assert all(len(set(r.tolist())) == k for r in I[:16].cpu())
    sub = 48
    Dr, Ir = _torch_reference(q[:sub], n, k)
    got = I[:sub].cpu()
    recall = np.mean([len(set(Ir[r].tolist()) & set(got[r].tolist())) / k for r in range(sub)])
    assert recall >= 0.999, recall                       # north_star: recall@k vs reference >= 0.999
    torch.testing.assert_close(D[:sub], Dr, rtol=1e-4, atol=1e-3)   # scores within 1e-4 relative
    assert (I[:sub] == Ir).float().mean() > 0.995        # ids identical except near-ties
    # the 1-CTA and CTA-pair kernels agree bit for bit on ids and scores
    from denseretrievaltoolkits_b200 import _lib

    D1, I1 = index.search(q[:256], k, flags=_lib.SEARCH_FORCE_1CTA)
    D2, I2 = index.search(q[:256], k, flags=_lib.SEARCH_FORCE_2CTA)
    assert torch.equal(I1, I2) and torch.equal(D1, D2)
    assert torch.equal(I1, I[:256])                      # batch-size independence / idempotence


def test_small_query_batch_is_hbm_bound_shape(full_store):
    """Trainer-style call (eval batch 16..128): results equal the big-batch rows."""
    import bench

    index = full_store
    q = bench.make_queries(torch, 1024, torch.device("cuda", 0))
    D, I = index.search(q, 100)
    Ds, Is = index.search(q[:16], 100)
    assert torch.equal(Is, I[:16]) and torch.equal(Ds, D[:16])
    assert index.search_stats()["ctas_per_tile"] == 1
