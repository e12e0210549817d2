"""GPU tests of the exactness certificate (rigorous per-row error bound).

faiss.IndexFlatIP.search is exact for every row (DRT/evaluator/index.py:32).  The first pass
here is bf16, so exactness has to be PROVEN per query: these cases are built so that an
estimate based on the errors observed among the candidates (round 1's check) returns wrong
lists without noticing — heavy-tailed row norms, and a high-norm row whose bf16 rounding
conspires against one query."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import bound  # noqa: E402


def _mk(d=768, seg_rows=0):
    from denseretrievaltoolkits_b200 import faiss_compat

    return faiss_compat.IndexFlatIP(d, device=0, seg_rows=seg_rows)


def _exact_f64(x, q, k):
    s = q.astype(np.float64) @ x.astype(np.float64).T
    I = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(s, I, axis=1), I


def _assert_lists_equal_up_to_fp32_ties(D, I, Dr, Ir, x, q):
    """ids equal the float64 brute force except where two exact scores differ by less than the
    fp32 resolution of the rescoring dot product; scores within that resolution.  The resolution
    is relative to |q||d| (a small score of a huge row is a cancellation), so atol carries the
    largest row norm: 3e-7 |q|max |d|max ~ a few ulp of the largest partial sums."""
    atol = 3e-7 * float(np.linalg.norm(q, axis=1).max()) * float(np.linalg.norm(x.astype(np.float64), axis=1).max())
    np.testing.assert_allclose(D, Dr, rtol=2e-6, atol=atol)
    diff = I != Ir
    if diff.any():
        assert np.all(np.abs(D[diff] - Dr[diff]) <= 2e-6 * np.abs(Dr[diff]) + atol)
        assert diff.mean() < 0.002


def test_planted_high_norm_rows_whose_first_pass_score_falls_below_the_candidates():
    """Two planted rows, each the exact top-1 of one query, each built so that 16-bit rounding —
    bf16 for row A / query 0, fp16 for row B / query 1 — removes ~600 from its first-pass score
    (see oracle/bound.py::adverse_row).  Whichever format the library's first pass uses, one of
    them has a NEGATIVE first-pass score, far below the ~k' best first-pass scores (> 80): an error
    estimate taken from the candidates (max ~0.2; round 1's check) would never look at it."""
    rng = np.random.default_rng(21)
    n, d, nq, k = 100_000, 768, 64, 10
    x = rng.standard_normal((n, d), dtype=np.float32)
    q = rng.standard_normal((nq, d), dtype=np.float32)
    ia, ib = 54_321, 12_345
    x[ia] = bound.adverse_row(q[0], 256.0, "bf16")
    x[ib] = bound.adverse_row(q[1], 2048.0, "f16")
    for qi, row, fmt in ((0, ia, "bf16"), (1, ib, "f16")):
        assert float(q[qi].astype(np.float64) @ x[row].astype(np.float64)) > 250.0
        assert float(bound.first_pass_scores(x[row:row + 1], q[qi:qi + 1], fmt)[0, 0]) < 0.0
    index = _mk(seg_rows=1 << 15)
    index.add(x)
    D, I = index.search(q, k)
    st = index.search_stats()
    Dr, Ir = _exact_f64(x, q, k)
    # (row A has norm ~7,000, so it can also outrank row B for query 1: B is rank 1 or 2 there)
    assert Ir[0, 0] == ia and ib in Ir[1, :2]
    assert I[0, 0] == ia and ib in I[1, :2], "the planted rows must be found: exactness is a proof, not an estimate"
    _assert_lists_equal_up_to_fp32_ties(D, I, Dr, Ir, x, q)
    assert st["flagged_queries"] == 0


def test_values_beyond_the_fp16_range_and_tiny_queries():
    """fp16 images saturate at +-65504 and flush below 6e-8: rows with huge components, queries
    with huge or tiny magnitudes (scaled by a power of two), all still exact."""
    rng = np.random.default_rng(26)
    n, d, nq, k = 60_000, 768, 48, 20
    x = rng.standard_normal((n, d), dtype=np.float32)
    x[100] *= 1e6                                    # every component beyond the fp16 range
    x[200, 5] = 3e5
    x[300, 7] = -2e5
    q = rng.standard_normal((nq, d), dtype=np.float32)
    q[3] *= 1e6
    q[4] *= 1e-9
    q[5] *= 40000.0
    index = _mk(seg_rows=1 << 14)
    index.add(x)
    D, I = index.search(q, k)
    st = index.search_stats()
    Dr, Ir = _exact_f64(x, q, k)
    np.testing.assert_allclose(D, Dr, rtol=3e-5, atol=0)
    diff = I != Ir
    assert diff.mean() < 0.005 and np.all(np.abs(D[diff] - Dr[diff]) <= 3e-5 * np.abs(Dr[diff]))
    assert st["flagged_queries"] == 0


@pytest.mark.parametrize("k", [10, 100])
def test_lognormal_row_norms_100x_spread(k):
    rng = np.random.default_rng(22)
    n, d, nq = 200_000, 768, 96
    x = rng.standard_normal((n, d), dtype=np.float32)
    x *= np.exp(rng.normal(0.0, np.log(100.0) / 4.0, size=(n, 1))).astype(np.float32)   # +-2 sigma = 100x
    q = rng.standard_normal((nq, d), dtype=np.float32)
    index = _mk(seg_rows=1 << 16)
    index.add(x)
    D, I = index.search(q, k)
    st = index.search_stats()
    Dr, Ir = _exact_f64(x, q, k)
    _assert_lists_equal_up_to_fp32_ties(D, I, Dr, Ir, x, q)
    assert st["flagged_queries"] == 0
    # also through the device API and a second add (tile maxima accumulate across adds)
    index.add(x[:1000] * 3.0)
    x2 = np.concatenate([x, x[:1000] * 3.0])
    D2, I2 = index.search(torch.from_numpy(q).cuda(), k)
    Dr2, Ir2 = _exact_f64(x2, q, k)
    _assert_lists_equal_up_to_fp32_ties(D2.cpu().numpy(), I2.cpu().numpy(), Dr2, Ir2, x2, q)


def test_a_few_huge_rows_do_not_poison_every_query():
    """One in 5,000 rows has a 1000x norm.  A single global error bound would be ~1000x too
    large for every other row and flag every query; per-row bounds keep the certificate tight."""
    rng = np.random.default_rng(23)
    n, d, nq, k = 150_000, 768, 128, 100
    x = rng.standard_normal((n, d), dtype=np.float32)
    big = rng.choice(n, size=n // 5000, replace=False)
    x[big] *= 1000.0
    q = rng.standard_normal((nq, d), dtype=np.float32)
    index = _mk(seg_rows=1 << 15)
    index.add(x)
    D, I = index.search(q, k)
    st = index.search_stats()
    Dr, Ir = _exact_f64(x, q, k)
    _assert_lists_equal_up_to_fp32_ties(D, I, Dr, Ir, x, q)
    assert st["flagged_queries"] == 0 and st["exact_queries"] == 0
    assert st["refined_queries"] <= nq // 4


def test_outlier_rows_do_not_flood_the_candidate_buffers():
    """An outlier row makes its 256-row tile's bound so large that the tile-level test would pass
    every row of the tile: 120 such tiles in one segment are 30,720 admissions per query, over the
    16,384-entry candidate buffer (4.4M rows of this family took 601 launches and 57 ms before the
    filter re-tested rows of mixed-norm tiles against their own bounds; 15 launches / 10.5 ms after)."""
    rng = np.random.default_rng(29)
    n, d, nq, k = 160_000, 768, 64, 100
    x = rng.standard_normal((n, d), dtype=np.float32)
    big = np.arange(120) * 1280 + 77          # distinct tiles, one segment
    x[big] *= 1000.0
    q = rng.standard_normal((nq, d), dtype=np.float32)
    index = _mk()
    index.add(x)
    D, I = index.search(q, k)
    st = index.search_stats()
    Dr, Ir = _exact_f64(x, q, k)
    _assert_lists_equal_up_to_fp32_ties(D, I, Dr, Ir, x, q)
    assert st["overflow_retries"] == 0, st
    assert st["flagged_queries"] == 0 and st["exact_queries"] == 0


def test_certificate_statistics_on_gaussian_data():
    """The default k' certifies (nearly) every query of the benchmark distribution in one pass."""
    rng = np.random.default_rng(24)
    n, d, nq = 400_000, 768, 512
    x = torch.randn((n, d), generator=torch.Generator().manual_seed(24)).numpy()
    q = rng.standard_normal((nq, d), dtype=np.float32)
    index = _mk()
    index.add(x)
    for k in (100, 1000):
        D, I = index.search(q, k)
        st = index.search_stats()
        assert st["flagged_queries"] == 0 and st["exact_queries"] == 0
        assert st["refined_queries"] <= nq // 50, st
        Dr, Ir = _exact_f64(x, q[:64], k)
        _assert_lists_equal_up_to_fp32_ties(D[:64], I[:64], Dr, Ir, x, q)


def test_small_index_allocates_a_small_segment():
    """a 10-row IndexFlatIP(768) used to pin a full 2^20-row segment (4.8 GB)."""
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info(0)
    idx = [_mk() for _ in range(8)]
    rng = np.random.default_rng(25)
    x = rng.standard_normal((10, 768), dtype=np.float32)
    for i in idx:
        i.add(x)
    free1, _ = torch.cuda.mem_get_info(0)
    assert free0 - free1 < 1 << 30
    q = rng.standard_normal((3, 768), dtype=np.float32)
    D, I = idx[0].search(q, 4)
    np.testing.assert_array_equal(I, np.argsort(-(q @ x.T), axis=1, kind="stable")[:, :4])
    # growth keeps rows and ids: 10 -> 5000 -> 70000 rows through several re-allocations
    big = rng.standard_normal((70_000, 768), dtype=np.float32)
    idx[0].add(big[:4990])
    idx[0].add(big[4990:])
    allx = np.concatenate([x, big])
    np.testing.assert_array_equal(idx[0].reconstruct_n(0, allx.shape[0]), allx)
    D, I = idx[0].search(q, 10)
    Dr, Ir = _exact_f64(allx, q, 10)
    np.testing.assert_array_equal(I, Ir)


def test_fp16_first_pass_mode_in_a_subprocess():
    """DRT_B200_FIRST_PASS=f16 (fp16 images on both sides: 8x tighter bound, k' 140 instead of 200)
    is read once per process, so the adversarial and range cases are re-run in a child process."""
    import os
    import subprocess
    import sys

    env = dict(os.environ, DRT_B200_FIRST_PASS="f16")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_certificate.py"), "-q", "-m", "gpu", "-x",
                        "-k", "planted or beyond_the_fp16 or huge_rows or statistics"], env=env, capture_output=True, text=True,
                       timeout=900, cwd=root)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert "4 passed" in p.stdout, p.stdout[-500:]
