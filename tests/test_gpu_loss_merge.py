"""GPU (B200) parity tests: fused in-batch CE (fwd + bwd), cross-shard merge kernel, mining
filter — all through the C ABI."""
import glob
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import inbatch_loss  # noqa: E402
from oracle import merge as omerge  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "loss_*.npz"))))
def test_loss_matches_reference_goldens(path):
    """Values produced by the reference's own SimpleContrastiveLoss (tools/make_golden.py)."""
    from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss

    g = np.load(path)
    B, n, d = int(g["B"]), int(g["n"]), int(g["d"])
    red = str(g["reduction"])
    if "x" in g:
        x, y = g["x"], g["y"]
    else:
        rng = np.random.default_rng(int(g["seed"]))
        x = rng.standard_normal((B, d)).astype(np.float32)
        y = rng.standard_normal((B * n, d)).astype(np.float32)
        if not (np.array_equal(x[:2], g["x_head"]) and np.array_equal(y[:2], g["y_head"])):
            pytest.skip("numpy RNG stream differs from the authoring container")
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    yt = torch.from_numpy(y).cuda().requires_grad_(True)
    tgt = torch.from_numpy(g["target"]).cuda() if "target" in g else None
    loss = SimpleContrastiveLoss()(xt, yt, target=tgt, reduction=red)
    np.testing.assert_allclose(loss.detach().cpu().numpy(), g["loss"], rtol=1e-4, atol=1e-5)   # 1e-4 relative
    (loss.sum() if red == "none" else loss).backward()
    dx, dy = xt.grad.cpu().numpy(), yt.grad.cpu().numpy()
    if "dx" in g:
        np.testing.assert_allclose(dx, g["dx"], rtol=1e-3, atol=2e-5)
        np.testing.assert_allclose(dy, g["dy"], rtol=1e-3, atol=2e-5)
    else:
        np.testing.assert_allclose(dx[:8], g["dx_rows"], rtol=1e-3, atol=2e-5)
        np.testing.assert_allclose(dy[:16], g["dy_rows"], rtol=1e-3, atol=2e-5)
        np.testing.assert_allclose(dx.sum(0), g["dx_colsum"], rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(dy.sum(0), g["dy_colsum"], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("B,n,d", [(128, 8, 768), (16, 2, 768), (5, 3, 64), (33, 7, 96), (256, 8, 768), (512, 8, 768), (1024, 8, 768)])
def test_loss_matches_torch_fp32_and_oracle(B, n, d):
    from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss, inbatch_scores_and_loss

    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator(device="cuda").manual_seed(B * 7 + n)
    x = torch.randn((B, d), generator=gen, device="cuda")
    y = torch.randn((B * n, d), generator=gen, device="cuda")
    x1, y1 = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    x2, y2 = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    loss = SimpleContrastiveLoss()(x1, y1)
    target = torch.arange(0, B * n, n, device="cuda")
    ref = torch.nn.functional.cross_entropy(x2 @ y2.t(), target)
    torch.testing.assert_close(loss, ref, rtol=1e-4, atol=1e-5)
    loss.backward()
    ref.backward()
    torch.testing.assert_close(x1.grad, x2.grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(y1.grad, y2.grad, rtol=1e-3, atol=1e-5)
    o, _, _ = inbatch_loss.contrastive_loss(x.cpu().numpy(), y.cpu().numpy())
    assert abs(loss.item() - o) <= 1e-4 * abs(o)
    # DRModel.forward loss block: also returns the score matrix (biencoder.py:107-122)
    l2, scores = inbatch_scores_and_loss(x, y, n)
    torch.testing.assert_close(l2, ref.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(scores, x @ y.t(), rtol=1e-4, atol=1e-3)


def test_merge_kernel_matches_reference_goldens():
    from denseretrievaltoolkits_b200.store import _cuda_merge

    cases = json.load(open(os.path.join(GOLD, "merge_cases.json")))
    for c in cases:
        results, topk = c["results"], c["topk"]
        qids = list(results[0].keys())
        G, k_in = len(results), max(len(r[q]) for r in results for q in qids)
        scores = np.full((G, len(qids), k_in), np.float32(-3.4028234663852886e38), np.float32)
        ids = np.full((G, len(qids), k_in), -1, np.int64)
        for g, res in enumerate(results):
            for qi, q in enumerate(qids):
                for j, (doc, sc) in enumerate(res[q].items()):
                    ids[g, qi, j], scores[g, qi, j] = int(doc), sc
        k_out = min(topk, G * k_in)
        D, I = _cuda_merge(torch.from_numpy(scores).cuda(), torch.from_numpy(ids).cuda(), k_out)
        Do, Io = omerge.merge_topk(scores, ids, k_out)
        np.testing.assert_array_equal(I.cpu().numpy(), Io)
        np.testing.assert_array_equal(D.cpu().numpy(), Do)
        for qi, q in enumerate(qids):
            want = c["merged"][q][:k_out]
            assert [int(i) for i in I[qi].tolist() if i >= 0][: len(want)] == [int(doc) for doc, _ in want]


def test_merge_kernel_random_with_padding_and_duplicates():
    from denseretrievaltoolkits_b200.store import _cuda_merge

    rng = np.random.default_rng(3)
    for G, Q, k_in, k_out in [(8, 50, 100, 100), (2, 7, 1000, 1000), (4, 3, 5, 20), (3, 9, 2048, 500)]:
        ids = rng.integers(0, 5000, size=(G, Q, k_in)).astype(np.int64)
        scores = rng.standard_normal((G, Q, k_in)).astype(np.float32)
        pad = rng.random((G, Q, k_in)) < 0.1
        ids[pad] = -1
        scores[pad] = np.float32(-3.4028234663852886e38)
        D, I = _cuda_merge(torch.from_numpy(scores).cuda(), torch.from_numpy(ids).cuda(), k_out)
        Do, Io = omerge.merge_topk(scores, ids, k_out)
        if G * k_in <= 8192:
            np.testing.assert_array_equal(I.cpu().numpy(), Io)
            np.testing.assert_array_equal(D.cpu().numpy(), Do)
        else:   # hierarchical fold: same set and order when ids are unique per query is not
                # guaranteed with duplicates across halves; check scores are the top ones
            np.testing.assert_allclose(D.cpu().numpy()[:, :10], Do[:, :10])


def test_sorted_unique_merge_equals_generic_merge():
    """Rank-based merge (ordered, id-disjoint shard lists) == sort-based merge == oracle."""
    from denseretrievaltoolkits_b200.store import _cuda_merge

    rng = np.random.default_rng(13)
    for G, Q, k_in, k_out in [(8, 40, 100, 100), (2, 9, 1000, 1000), (4, 5, 7, 28), (3, 6, 200, 50), (8, 3, 1000, 1000)]:
        scores = np.sort(np.round(rng.standard_normal((G, Q, k_in)), 1).astype(np.float32), axis=2)[:, :, ::-1].copy()
        ids = np.empty((G, Q, k_in), np.int64)
        for g in range(G):                                  # disjoint ascending id ranges; ties -> id asc
            for q in range(Q):
                raw = np.sort(rng.choice(10000, size=k_in, replace=False)) + g * 10000
                order = np.lexsort((raw, -scores[g, q]))    # keep canonical order inside equal scores
                ids[g, q] = raw
                scores[g, q] = scores[g, q][order]
        npad = rng.integers(0, max(1, k_in // 3), size=(G, Q))
        for g in range(G):
            for q in range(Q):
                if npad[g, q]:
                    ids[g, q, -npad[g, q]:] = -1
                    scores[g, q, -npad[g, q]:] = np.float32(-3.4028234663852886e38)
        sc, idt = torch.from_numpy(scores).cuda(), torch.from_numpy(ids).cuda()
        Df, If = _cuda_merge(sc, idt, k_out, sorted_unique=True)
        Do, Io = omerge.merge_topk(scores, ids, k_out)
        np.testing.assert_array_equal(If.cpu().numpy(), Io)
        np.testing.assert_array_equal(Df.cpu().numpy(), Do)
        if G * k_in <= 8192:
            Dg, Ig = _cuda_merge(sc, idt, k_out)
            assert torch.equal(Ig, If) and torch.equal(Dg, Df)


def test_mining_filter_matches_reference_loop():
    from denseretrievaltoolkits_b200.mining import filter_negatives

    for c in json.load(open(os.path.join(GOLD, "mining.json"))):
        out = filter_negatives(torch.tensor([c["ids"]], dtype=torch.int64).cuda(), torch.tensor([c["b"]]),
                               torch.tensor([c["e"]]), c["num_negative"])
        assert [int(v) for v in out[0].tolist() if v >= 0] == c["kept"]
    rng = np.random.default_rng(4)
    ids = rng.integers(-1, 1000, size=(300, 230)).astype(np.int64)
    pb = rng.integers(0, 900, size=300).astype(np.int64)
    pe = pb + rng.integers(1, 100, size=300)
    out = filter_negatives(torch.from_numpy(ids).cuda(), torch.from_numpy(pb), torch.from_numpy(pe), 200)
    np.testing.assert_array_equal(out.cpu().numpy(), omerge.filter_negatives(ids, pb, pe, 200))


def test_dense_mining_end_to_end():
    from denseretrievaltoolkits_b200 import faiss_compat
    from denseretrievaltoolkits_b200.mining import mine_hard_negatives
    from oracle import flat_ip

    rng = np.random.default_rng(12)
    x = rng.standard_normal((20000, 768), dtype=np.float32)
    q = rng.standard_normal((40, 768), dtype=np.float32)
    pb = rng.integers(0, 19000, size=40).astype(np.int64)
    pe = pb + rng.integers(1, 4, size=40)
    q += x[pb]                                    # the positive really is retrieved near the top
    index = faiss_compat.IndexFlatIP(768, device=0, seg_rows=4096)
    index.add(x)
    neg = mine_hard_negatives(index, torch.from_numpy(q).cuda(), torch.from_numpy(pb), torch.from_numpy(pe), 50, depth=60)
    _, Ir = flat_ip.flat_ip_search(x, q, 60)
    ref = omerge.filter_negatives(Ir, pb, pe, 50)
    assert (neg.cpu().numpy() == ref).mean() > 0.999
    for r in range(40):
        assert not ((neg[r].cpu().numpy() >= pb[r]) & (neg[r].cpu().numpy() < pe[r])).any()


def test_cfg5_shaped_mining_top200_with_positive_exclusion_in_batches(tmp_path):
    """BASELINE cfg5's shape at test size: depth num_negative + |positives| = 200 + up to 3, the
    query stream cut into several internal batches, a row-sharded store underneath, positive
    exclusion on the device, and the JSONL the reference's sampler reads (sampler.py:57-66)."""
    import json

    from denseretrievaltoolkits_b200.mining import mine_hard_negatives, write_negatives_jsonl
    from denseretrievaltoolkits_b200.store import ShardedCorpusStore
    from oracle import flat_ip

    rng = np.random.default_rng(13)
    n, nq, num_negative = 50_000, 300, 200
    x = rng.standard_normal((n, 768), dtype=np.float32)
    q = rng.standard_normal((nq, 768), dtype=np.float32)
    pb = rng.integers(0, n - 4, size=nq).astype(np.int64)
    pe = pb + rng.integers(1, 4, size=nq)
    q += 2.0 * x[pb]                                 # positives sit at the top of the ranking
    store = ShardedCorpusStore(768, num_virtual_shards=4, device=0, seg_rows=1 << 13)
    store.add_split(torch.from_numpy(x).cuda())
    store.finalize()
    neg = mine_hard_negatives(store, torch.from_numpy(q).cuda(), torch.from_numpy(pb), torch.from_numpy(pe), num_negative,
                              batch_size=128)        # 3 internal batches (128 + 128 + 44)
    neg = neg.cpu().numpy()
    depth = num_negative + int((pe - pb).max())
    _, Ir = flat_ip.flat_ip_search(x, q, depth)
    ref = omerge.filter_negatives(Ir, pb, pe, num_negative)
    assert neg.shape == (nq, num_negative) and (neg >= 0).all()
    assert (neg == ref).mean() > 0.999               # differing entries: fp32 near-ties at depth 200
    assert not ((neg >= pb[:, None]) & (neg < pe[:, None])).any()
    passages = {i: [int(i), int(i) + 1] for i in np.unique(neg)}
    samples = [{"query": [int(r)], "positives": [[int(p)] for p in range(pb[r], pe[r])]} for r in range(nq)]
    path = tmp_path / "bm25negatives"
    write_negatives_jsonl(str(path), samples, neg, passages)
    recs = [json.loads(line) for line in open(path, encoding="utf-8")]
    assert len(recs) == nq and all(len(r["negatives"]) == num_negative for r in recs)
    assert recs[17]["negatives"][0] == passages[int(neg[17, 0])] and recs[17]["positives"] == samples[17]["positives"]


def test_tensor_core_loss_path_is_fp32_accurate(monkeypatch):
    """Large shapes run the logits / dx / dy contractions on tcgen05 through an exact 3-way bf16
    split (6 partial products, fp32 accumulation).  Scores must stay within the path's 1e-4
    relative tolerance — measured ~1e-6 against float64 — and agree with the fp32 SIMT path."""
    from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss, inbatch_scores_and_loss

    B, n, d = 1024, 8, 768
    gen = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn((B, d), generator=gen, device="cuda") * 3.0
    y = torch.randn((B * n, d), generator=gen, device="cuda") * 0.3 + 0.1
    loss_tc, scores = inbatch_scores_and_loss(x, y, n)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref64 = x.double() @ y.double().t()
    scale = ref64.abs().max().item()
    err_tc = (scores.double() - ref64).abs().max().item()
    err_fp32 = ((x @ y.t()).double() - ref64).abs().max().item()      # cuBLAS sgemm, the reference's arithmetic
    # fp32 accumulation inside the tensor core truncates (~0.5 ulp of the accumulator per MMA step),
    # so the six partial products are accumulated in ascending magnitude and only the last 48 steps
    # (hi*hi) run at full scale: ~1e-6 of the score scale, on par with cuBLAS' sgemm (printed
    # beside it) and 4000x better than a single bf16 pass (~4e-3 * scale)
    print(f"tensor-core logits: max abs err {err_tc:.3e} (scale {scale:.1f}), cuBLAS fp32 {err_fp32:.3e}")
    assert err_tc <= 3e-6 * scale and err_tc <= 2.0 * err_fp32, (err_tc, err_fp32, scale)
    x1, y1 = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    l1 = SimpleContrastiveLoss()(x1, y1)
    l1.backward()
    monkeypatch.setenv("DRT_B200_CE_SIMT", "1")              # same call on the fp32 CUDA-core path
    x2, y2 = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    l2 = SimpleContrastiveLoss()(x2, y2)
    l2.backward()
    monkeypatch.delenv("DRT_B200_CE_SIMT")
    torch.testing.assert_close(l1, l2, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(loss_tc, l2.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(x1.grad, x2.grad, rtol=1e-3, atol=1e-4 * x2.grad.abs().max().item())
    torch.testing.assert_close(y1.grad, y2.grad, rtol=1e-3, atol=1e-4 * y2.grad.abs().max().item())
    # ragged large shape (M, N not multiples of the 128 / 256 tiles; K multiple of 32)
    xr, yr = x[:1000].contiguous(), y[:7968].contiguous()
    lr_, sr = inbatch_scores_and_loss(xr, yr, 7)
    ref = xr.double() @ yr.double().t()
    assert (sr.double() - ref).abs().max().item() <= 3e-6 * ref.abs().max().item()
    tgt = torch.arange(0, 1000 * 7, 7, device="cuda")
    torch.testing.assert_close(lr_, torch.nn.functional.cross_entropy(ref.float(), tgt), rtol=1e-4, atol=1e-5)
