"""Many-query cosine search on the tensor pipe (search_tc.cuh): a split-bf16 tcgen05 GEMM filters candidates, the canonical
fmaf chain re-scores them -- ids and scores must stay BIT-EXACT against the oracle (apply_r.lua:265-282), the
measured approximation error must sit well inside the bound the filter assumes, and every escape hatch (special rows,
special queries, candidate overflow) must land on the exact answer."""
import numpy as np
import pytest

from util import assert_bitexact

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


def _db(N, d, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(size=(N, d), dtype=np.float32) * np.float32(scale))


def _check(orc, ctx, db, q, k, expect_tc=True):
    want_ids, want_sc = orc.search_cosine(db, q, k)
    ctx.set_option("search_tc", 1)
    ctx.db_set(db)
    s0, f0 = ctx.tc_counters()
    ids, sc = ctx.search_cosine(q, k)
    s1, f1 = ctx.tc_counters()
    np.testing.assert_array_equal(ids, want_ids)
    assert_bitexact(sc, want_sc, "scores (tensor-core path)")
    if expect_tc is not None:
        assert (s1 - s0 == 1) == expect_tc and (f1 - f0 == 1) == (not expect_tc), (s1 - s0, f1 - f0)
    ctx.set_option("search_tc", 0)                      # the fmaf-chain kernels alone give the same bits
    ids0, sc0 = ctx.search_cosine(q, k)
    ctx.set_option("search_tc", 1)
    np.testing.assert_array_equal(ids0, want_ids)
    assert_bitexact(sc0, want_sc, "scores (fmaf-chain kernels)")


TC_CASES = [
    # N, d, Q, k
    (40000, 100, 200, 20),      # BASELINE configs[3] shape, ragged query tile, two sampling levels
    (20000, 100, 4096, 20),     # the full needle batch
    (30000, 256, 300, 100),     # BASELINE configs[4] shape (four 64-column slices)
    (10000, 32, 64, 5),         # one slice
    (9001, 30, 130, 128),       # d not a multiple of 4 / 16, max k, ragged row tile
    (12000, 1024, 50, 20),      # pixelwise measure d = C*H*W = 1024: 16 slices
    (66000, 64, 100, 1),        # k = 1, two levels
]


@pytest.mark.parametrize("case", TC_CASES, ids=lambda c: "N%d_d%d_Q%d_k%d" % c)
def test_tc_search_exact(orc, ctx, case):
    N, d, Q, k = case
    db = _db(N, d, 71)
    rows = np.random.default_rng(72).choice(N, size=Q // 2, replace=False)
    q = np.concatenate([db[rows], _db(Q - Q // 2, d, 73)])
    _check(orc, ctx, db, q, k)


def test_tc_search_scales_and_offsets(orc, ctx):
    """Rows and queries of wildly different norms, a common offset (clustered cosines near 1), tiny and huge magnitudes."""
    rng = np.random.default_rng(75)
    N, d, Q, k = 30000, 100, 128, 20
    db = _db(N, d, 76)
    db *= np.exp(rng.uniform(-20, 20, size=(N, 1))).astype(np.float32)          # norms from 1e-9 to 1e9
    db[:10000] += np.float32(5.0) * np.linalg.norm(db[:10000], axis=1, keepdims=True) / np.float32(10.0)   # a shared direction
    db[20000] = 0.0                                                                 # a zero row: every cosine 0
    db[20001] = 1e-25                                                               # |x|^2 far below the 1e-12 guard
    q = np.concatenate([db[rng.choice(N, size=64, replace=False)], _db(64, d, 77, 1e-6)])
    _check(orc, ctx, db, q, k, expect_tc=None)        # the clustered third may make the filter decline: exact either way
    _check(orc, ctx, db[10000:], q, k, expect_tc=True)


def test_tc_search_special_rows(orc, ctx):
    """NaN / inf rows are packed as zeros and re-scored exactly for every query at the final level."""
    N, d, Q, k = 20000, 64, 96, 20
    db = _db(N, d, 78)
    db[5, 3] = np.nan
    db[777] = np.inf
    db[778, 0] = -np.inf
    db[12345] = 3e38                          # |x|^2 overflows to inf: cosine dot*0 -> NaN or 0
    db[19999, 63] = np.nan
    q = np.concatenate([db[[1, 2, 3]], _db(Q - 3, d, 79)])
    _check(orc, ctx, db, q, k)
    # k larger than the number of finite-scored rows near the end: NaN rows fill the tail in id order
    _check(orc, ctx, db[:9000], q, 128)


def test_tc_search_fallbacks(orc, ctx):
    """Candidate overflow (thousands of exact duplicates), a NaN / zero query: flagged, answered by the fmaf-chain kernels."""
    rng = np.random.default_rng(80)
    N, d, Q, k = 16000, 32, 64, 20
    base = _db(50, d, 81)
    db = base[rng.integers(0, 50, size=N)]                          # 320 copies of each vector: ties everywhere, lists still fit
    q = np.concatenate([base[:32], _db(32, d, 82)])
    _check(orc, ctx, db, q, k, expect_tc=True)
    db = base[rng.integers(0, 5, size=N)]                           # 3200 copies of each: more candidates than a list holds
    _check(orc, ctx, db, q, k, expect_tc=False)
    db2 = _db(N, d, 83)
    q2 = _db(Q, d, 84)
    q2[7] = 0.0                                                     # zero query: every score 0 -> lowest ids
    _check(orc, ctx, db2, q2, k, expect_tc=False)
    q2[7] = np.nan
    _check(orc, ctx, db2, q2, k, expect_tc=False)


@pytest.mark.parametrize("d", [32, 100, 256, 1024])
def test_tc_error_bound_holds(orc, ctx, d):
    """The filter's only assumption: |approximate cosine - fmaf-chain cosine| <= eps(d) = 2^-13 + d*2^-20.  Measured on the
    device over every pair of a 4096-row database x 256 queries (incl. near-duplicates, where the score is ~1)."""
    N, Q = 4096, 256
    db = _db(N, d, 90 + d)
    db[100:200] = db[0] + 0.01 * _db(100, d, 91)
    db *= np.exp(np.random.default_rng(92).uniform(-8, 8, size=(N, 1))).astype(np.float32)
    q = np.concatenate([db[:128], _db(128, d, 93)])
    ctx.db_set(db)
    approx, eps = ctx.tc_scores(q)
    assert abs(eps - (2.0 ** -13 + d * 2.0 ** -20)) < 1e-9
    ids, exact = orc.search_cosine(db, q, 128)                       # exact chain scores of each query's 128 best rows
    worst = float(np.abs(np.take_along_axis(approx, ids, axis=1) - exact).max())
    # and of arbitrary pairs: float64 cosine differs from the chain by far less than eps
    ref = (q.astype(np.float64) @ db.astype(np.float64).T) / (np.linalg.norm(q.astype(np.float64), axis=1)[:, None] * np.linalg.norm(db.astype(np.float64), axis=1)[None, :])
    worst_all = float(np.abs(approx - ref).max())
    assert worst <= eps / 8 and worst_all <= eps / 8, (worst, worst_all, eps)


def test_tc_search_full_size(ctx):
    """BASELINE configs[3] size (1M x 100, 4096 needles, top-20): tensor-core path == fmaf-chain kernels, bit for bit."""
    rng = np.random.default_rng(31)
    N, d, Q, k = 1_000_000, 100, 4096, 20
    db = rng.standard_normal(size=(N, d), dtype=np.float32)
    rows = (np.arange(1, Q + 1, dtype=np.int64) * 244)
    ctx.db_set(db)
    ctx.set_option("search_tc", 1)
    s0, _ = ctx.tc_counters()
    ids, sc = ctx.search_rows(rows, k)
    assert ctx.tc_counters()[0] == s0 + 1
    assert (ids[:, 0] == rows).all()
    ctx.set_option("search_tc", 0)
    ids0, sc0 = ctx.search_rows(rows, k)
    ctx.set_option("search_tc", 1)
    np.testing.assert_array_equal(ids, ids0)
    assert_bitexact(sc, sc0)
