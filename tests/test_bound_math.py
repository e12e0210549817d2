"""CPU checks of the exactness certificate's algebra (oracle/bound.py): the upper bound really
bounds the exact score — for random data, heavy-tailed norms and rows built to make bf16
rounding conspire against one query — and a certified query's list equals brute force."""
import numpy as np
import pytest

from oracle import bound, flat_ip


@pytest.mark.parametrize("qfmt", ["f16", "bf16"])
@pytest.mark.parametrize("case", ["gauss", "lognormal", "aniso", "tail"])
def test_upper_bound_holds(case, qfmt):
    rng = np.random.default_rng(11)
    d, n, nq = 192, 3000, 40
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    split = None
    if case == "lognormal":
        x *= np.exp(rng.normal(0.0, np.log(100.0) / 4.0, size=(n, 1))).astype(np.float32)
    elif case == "aniso":
        sc = np.exp(rng.uniform(np.log(0.25), np.log(4.0), size=d)).astype(np.float32)
        x = x * sc + 0.3
        q = q * sc
    elif case == "tail":                            # squared-L2 augmentation: exact tail dims
        nrm = -0.5 * (x.astype(np.float64) ** 2).sum(1)
        n1 = bound.bf16_round(nrm.astype(np.float32))
        n2 = bound.bf16_round((nrm - n1).astype(np.float32))
        n3 = bound.bf16_round((nrm - n1 - n2).astype(np.float32))
        x = np.concatenate([x, n1[:, None], n2[:, None], n3[:, None]], axis=1)
        q = np.concatenate([q, np.ones((nq, 3), np.float32)], axis=1)
        split = d
    exact = q.astype(np.float64) @ x.astype(np.float64).T
    ub = bound.upper_bounds(x, q, split, qfmt)
    assert (ub >= exact - 1e-9 * np.abs(exact)).all()
    # and it is not vacuous: within ~sqrt(d) x the typical first-pass error
    typical = np.abs(bound.first_pass_scores(x, q, qfmt) - exact).mean()
    assert (ub - exact).mean() < 40.0 * np.sqrt(d) * typical


def test_rows_beyond_the_fp16_range_saturate_and_stay_covered():
    """fp16 images saturate at +-65504; the residual norm carries the rest, so the bound holds."""
    rng = np.random.default_rng(13)
    x = rng.standard_normal((500, 64)).astype(np.float32)
    x[7] *= 1e6
    x[9, 3] = -3e5
    q = rng.standard_normal((8, 64)).astype(np.float32)
    q[2] *= 1e-7                                     # tiny query: fp16 subnormals, covered by e_q
    exact = q.astype(np.float64) @ x.astype(np.float64).T
    assert (bound.upper_bounds(x, q, qfmt="f16") >= exact - 1e-9 * np.abs(exact)).all()
    assert (bound.upper_bounds(x, q, qfmt="bf16") >= exact - 1e-9 * np.abs(exact)).all()


def test_adverse_row_is_covered_and_certified_lists_are_exact():
    rng = np.random.default_rng(12)
    d, n, nq, k = 256, 4000, 16, 10
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    x[1234] = bound.adverse_row(q[0])
    sp = bound.first_pass_scores(x, q, "bf16")       # the row is built against bf16 rounding
    exact = q.astype(np.float64) @ x.astype(np.float64).T
    # the planted row is the exact top-1 of query 0 but sits far below the first-pass top-k
    assert np.argmax(exact[0]) == 1234
    assert sp[0, 1234] < np.sort(sp[0])[-4 * k]
    assert bound.upper_bounds(x, q, qfmt="bf16")[0, 1234] >= exact[0, 1234]
    assert bound.upper_bounds(x, q, qfmt="f16")[0, 1234] >= exact[0, 1234]
    ids, ok = bound.certified_topk(x, q, k, kprime=64, qfmt="bf16")
    _, Ir = flat_ip.flat_ip_search(x, q, k)
    assert ids[0, 0] == 1234
    np.testing.assert_array_equal(ids[ok], Ir[ok])
    assert ok.mean() > 0.5                          # the bound is tight enough to certify at k' = 64
