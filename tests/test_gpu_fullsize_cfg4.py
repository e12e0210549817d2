"""GPU (B200): BASELINE cfg4's corpus (Wikipedia-DPR scale, 21M x 768) on ONE B200 — 64.5 GB
fp32 + 32 GB bf16 of the 180 GB — 3,600 queries, top-100, a fixed 512-query subset against the
independent GPU fp32 reference.  Skipped when less than 110 GB of HBM is free."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_cfg4_21m_rows_single_gpu():
    import bench
    from denseretrievaltoolkits_b200 import faiss_compat
    from test_gpu_fullsize import SUB_GPU, _assert_parity, _torch_reference

    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info(0)
    if free < 110e9:
        pytest.skip("needs ~100 GB of free HBM")
    n, nq, k = 21_000_000, 3600, 100
    dev = torch.device("cuda", 0)
    index = faiss_compat.IndexFlatIP(bench.DIM, device=0)
    bench.fill_rows(torch, index.add, 0, n, dev)
    assert index.ntotal == n
    q = bench.make_queries(torch, nq, dev)
    D, I = index.search(q, k)
    st = index.search_stats()
    assert st["overflow_retries"] == 0 and st["flagged_queries"] == 0 and st["exact_queries"] == 0
    assert (D[:, 1:] <= D[:, :-1]).all() and (I >= 0).all() and (I < n).all()
    Dr, Ir = _torch_reference(q[:SUB_GPU], n, k)
    _assert_parity(D[:SUB_GPU], I[:SUB_GPU], Dr, Ir, k)
