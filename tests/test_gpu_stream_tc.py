"""GPU parity of the TMA -> tf32 tcgen05 filter pipeline (csrc/stream_tc.cuh) that serves cosine top-k for <= 32 needles
(apply_r.lua:267-282), kmeans for k <= 32 (unsup.kmeans at apply_r.lua:198) and the cosine-min assignment
(apply_r.lua:206-218).  The tensor core only FILTERS: every returned id, score, label, centroid and cosine must still be
bit-identical to the oracle's sequential fmaf chains, and to the fmaf-chain kernels the pipeline replaces."""
import numpy as np
import pytest

from util import assert_bitexact

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.set_option("stream_tc", 1)
    c.set_option("dbg", 0)
    c.close()


def _db(N, d, seed=0, scale=1.0):
    return (np.random.default_rng(seed).normal(size=(N, d)) * scale).astype(np.float32)


def _init(k, d, seed=6):
    c = np.random.default_rng(seed).normal(size=(k, d)).astype(np.float32)
    return (c / np.linalg.norm(c, axis=1, keepdims=True)).astype(np.float32)


# N, d, Q, k: every NQP variant (8 / 16 / 32 columns), both list lengths (k <= 32, k <= 128), d below one box, ragged last
# box, many boxes (the slot ring wraps inside a tile), ragged last tile, k > N
SEARCH = [
    (10000, 32, 4, 20), (10000, 32, 5, 100), (20011, 100, 4, 20), (5000, 100, 9, 20), (5000, 100, 16, 33), (5000, 100, 17, 20),
    (4097, 100, 32, 128), (2000, 1024, 5, 100), (3000, 256, 8, 100), (3000, 4, 3, 7), (3000, 20, 3, 7), (129, 36, 2, 100),
    (128, 64, 1, 1), (50, 32, 3, 100), (1, 32, 2, 3), (60000, 68, 6, 20),
]


@pytest.mark.parametrize("case", SEARCH, ids=lambda c: "N%d_d%d_Q%d_k%d" % c)
def test_search_exact_and_equal_to_chain_kernels(orc, ctx, case):
    N, d, Q, k = case
    db = _db(N, d, 1)
    q = np.concatenate([db[: min(Q, N) // 2], _db(Q - min(Q, N) // 2, d, 2)])[:Q]   # some needles are db rows (cos = 1 ties with themselves)
    want_ids, want_sc = orc.search_cosine(db, q, k)
    ctx.db_set(db)
    for stream_tc in (1, 0):
        ctx.set_option("stream_tc", stream_tc)
        ids, sc = ctx.search_cosine(q, k)
        np.testing.assert_array_equal(ids, want_ids)
        assert_bitexact(sc, want_sc, "scores")
    ctx.set_option("stream_tc", 1)


def test_search_near_ties_duplicates_and_specials(orc, ctx):
    """Rows that differ from each other by less than the tf32 resolution, exact duplicates (lowest id wins), NaN / inf /
    zero rows, a zero needle and a NaN needle: the filter may only ever pass extra candidates."""
    rng = np.random.default_rng(7)
    N, d = 30000, 100
    base = rng.standard_normal(size=(1, d)).astype(np.float32)
    x = (base + np.float32(1e-5) * rng.standard_normal(size=(N, d))).astype(np.float32)    # cosines differ in the 1e-10s
    x[5000:10000] = rng.standard_normal(size=(5000, d)).astype(np.float32)
    x[123] = x[77]; x[29999] = x[77]                                                        # exact duplicates
    x[200] = 0.0
    x[201, 3] = np.nan
    x[202, 5] = np.inf
    x[203] *= np.float32(1e-30)
    x[204] *= np.float32(1e18)
    q = np.stack([x[77], base[0], x[6000], np.zeros(d, np.float32), x[204], x[203]]).astype(np.float32)
    qn = q.copy(); qn[2, 0] = np.nan
    ctx.db_set(x)
    for qq in (q, qn):
        for k in (20, 100):
            wi, ws = orc.search_cosine(x, qq, k)
            ids, sc = ctx.search_cosine(qq, k)
            np.testing.assert_array_equal(ids, wi)
            assert_bitexact(sc, ws)


LABEL = [
    # N, d, k
    (5000, 100, 20), (5000, 100, 1), (5000, 100, 2), (5000, 100, 8), (5000, 100, 9), (4097, 100, 16), (5000, 100, 17), (5000, 100, 32),
    (3000, 32, 20), (3000, 4, 20), (3000, 20, 5), (3000, 128, 20), (3001, 256, 20), (3001, 256, 32), (2000, 288, 8), (2000, 192, 24), (127, 32, 20), (128, 32, 20),
    (129, 32, 20), (1, 32, 20), (60000, 68, 20),
]


@pytest.mark.parametrize("case", LABEL, ids=lambda c: "N%d_d%d_k%d" % c)
def test_kmeans_and_cosine_min_exact(orc, ctx, case):
    N, d, k = case
    x = _db(N, d, 31)
    init = _init(k, d, seed=32)
    want_c, want_t, want_l = orc.kmeans(x, k, 3, init)
    cenq = _db(k, d, 33)
    if N > 9:
        x[7] = cenq[0]; x[8] = cenq[k - 1]                          # rows equal to the first / last centroid
        want_c, want_t, want_l = orc.kmeans(x, k, 3, init)
    want_cl, want_cv = orc.assign_cosine_min(x, cenq)
    ctx.db_set(x)
    for dbg in (0, 1 << 16):                                        # normal; every row through the every-chain list kernel
        ctx.set_option("dbg", dbg)
        cen, tot, lab = ctx.kmeans(k, 3, init)
        np.testing.assert_array_equal(lab, want_l)
        assert_bitexact(tot, want_t, "total counts")
        assert_bitexact(cen, want_c, "centroids")
        cl, cv = ctx.assign_cosine_min(cenq)
        np.testing.assert_array_equal(cl, want_cl)
        assert_bitexact(cv, want_cv, "cos")
    ctx.set_option("dbg", 0)


def test_labels_adversarial(orc, ctx):
    """Duplicate centroids (every row an exact tie), a tight cluster whose rows differ below the tf32 resolution, 13 decades
    of dynamic range, zero / NaN / inf rows, a zero centroid and a NaN centroid."""
    rng = np.random.default_rng(55)
    N, d, k = 40000, 100, 20
    x = rng.standard_normal(size=(N, d), dtype=np.float32)
    x[:5000] = x[0] + np.float32(1e-4) * rng.standard_normal(size=(5000, d), dtype=np.float32)
    x[5000:10000] *= np.exp(rng.uniform(-15, 15, size=(5000, 1))).astype(np.float32)
    x[12345] = 0.0
    init = _init(k, d, seed=56)
    init[7] = init[3]
    init[9] = 0.0
    x[100] = init[5]; x[101] = init[19]
    for special in (False, True):
        if special:
            x[300, 1] = np.nan
            x[301, 2] = np.inf
            init[17, 3] = np.nan
        want_cl, want_cv = orc.assign_cosine_min(x, init)
        ctx.db_set(x)
        cl, cv = ctx.assign_cosine_min(init)
        np.testing.assert_array_equal(cl, want_cl)
        assert_bitexact(cv, want_cv)
        if not special:                                            # (kmeans refuses non-finite databases: fixed-point sums)
            want_c, want_t, want_l = orc.kmeans(x, k, 2, init)
            cen, tot, lab = ctx.kmeans(k, 2, init)
            np.testing.assert_array_equal(lab, want_l)
            assert_bitexact(tot, want_t); assert_bitexact(cen, want_c)


def test_full_size_equals_chain_kernels(ctx):
    """BASELINE configs[3] database size (1M x 100): the pipeline and the fmaf-chain kernels agree bit for bit on the top-k,
    the kmeans centroids / counts / labels and the cosine-min assignment (a size the oracle cannot finish in seconds)."""
    N, d, k = 1_000_000, 100, 20
    ctx.db_synthetic(N, d, seed=5)
    rows = np.array([99, 199, 299, 399], np.int64)
    init = _init(k, d, seed=9)
    out = {}
    for stream_tc in (1, 0):
        ctx.set_option("stream_tc", stream_tc)
        out[stream_tc] = (ctx.search_rows(rows, 20), ctx.kmeans(k, 2, init), ctx.assign_cosine_min(init))
    ctx.set_option("stream_tc", 1)
    a, b = out[1], out[0]
    np.testing.assert_array_equal(a[0][0], b[0][0]); assert_bitexact(a[0][1], b[0][1])
    assert_bitexact(a[1][0], b[1][0]); assert_bitexact(a[1][1], b[1][1]); np.testing.assert_array_equal(a[1][2], b[1][2])
    np.testing.assert_array_equal(a[2][0], b[2][0]); assert_bitexact(a[2][1], b[2][1])
    assert (a[0][0][:, 0] == rows).all()                           # every needle finds itself first


def test_error_bound_holds_on_device(ctx):
    """The filter's only assumption: |tf32 tensor-core score - fmaf chain| <= tfs_eps(d) |x||c|.  Under dbg bit 18 the
    cosine-min epilogue measures the ratio for the winner of every row (N rows x several data shapes); it must stay below 1
    with room to spare, and the chains-on-top-of-the-filter counter must stay a small fraction of the rows."""
    rng = np.random.default_rng(77)
    worst = 0.0
    for N, d, k, kind in ((400_000, 100, 20, "normal"), (200_000, 256, 32, "normal"), (300_000, 32, 20, "uniform"), (200_000, 100, 20, "lognormal")):
        if kind == "normal":
            x = rng.standard_normal(size=(N, d), dtype=np.float32)
        elif kind == "uniform":
            x = rng.random(size=(N, d), dtype=np.float32)             # all-positive rows: no cancellation, the largest relative errors
        else:
            x = (rng.standard_normal(size=(N, d)) * np.exp(rng.uniform(-8, 8, size=(N, d)))).astype(np.float32)
        cen = x[rng.choice(N, k, replace=False)] + rng.standard_normal(size=(k, d)).astype(np.float32) * 0.01
        ctx.db_set(x)
        ctx.tfs_stats()
        ctx.set_option("dbg", 4 << 16)
        ctx.assign_cosine_min(cen)
        ctx.set_option("dbg", 0)
        st = ctx.tfs_stats()
        assert st["launches"] == 1
        assert 0.0 < st["max_error_over_bound"] < 0.75, st
        assert st["listed_rows"] == 0, st
        worst = max(worst, st["max_error_over_bound"])
    print("largest observed |approximate - exact| / bound:", worst)
