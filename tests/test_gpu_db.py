"""GPU parity of the database kernels: cosine top-k (apply_r.lua:265-282), kmeans (unsup.kmeans,
apply_r.lua:198), cosine-min assignment (:206-218), cluster members + mean image (:222-243).
Everything here is BIT-EXACT against the oracle: ids, labels, scores, centroids."""
import numpy as np
import pytest

from util import assert_bitexact

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


def _db(N, d, seed=0, scale=1.0):
    return (np.random.default_rng(seed).normal(size=(N, d)) * scale).astype(np.float32)


SEARCH_CASES = [
    # N, d, Q, k
    (10000, 32, 5, 100),     # apply_r.lua's own values (5 needles, top-100)
    (10000, 32, 4, 20),      # BASELINE config 1
    (2000, 1024, 5, 100),    # pixelwise measure, d = C*H*W
    (5000, 100, 70, 20),     # 64-wide query tiles, ragged last tile
    (3000, 30, 33, 128),     # d not a multiple of 4 or 32, max k
    (777, 7, 17, 5),
    (50, 32, 3, 100),        # k > N: ids -1 past the end
    (1, 32, 2, 3),
]


@pytest.mark.parametrize("case", SEARCH_CASES, ids=lambda c: "N%d_d%d_Q%d_k%d" % c)
def test_search_exact(orc, ctx, case):
    N, d, Q, k = case
    db = _db(N, d, 1)
    q = np.concatenate([db[: min(Q, N) // 2], _db(Q - min(Q, N) // 2, d, 2)])[:Q]   # some queries are db rows
    want_ids, want_sc = orc.search_cosine(db, q, k)
    ctx.db_set(db)
    ids, sc = ctx.search_cosine(q, k)
    np.testing.assert_array_equal(ids, want_ids)
    assert_bitexact(sc, want_sc, "scores")


CONFIG5_SEARCH = [
    # BASELINE configs[4] shapes (d = 256, top-100) at sizes the oracle finishes in seconds
    (50000, 256, 70, 100),     # many queries, d > 128
    (30000, 256, 4096, 100),   # the full query batch
    (20000, 256, 17, 100),     # 17 queries: first size past the streaming family
]


@pytest.mark.parametrize("case", CONFIG5_SEARCH, ids=lambda c: "N%d_d%d_Q%d_k%d" % c)
def test_search_config5_shapes_exact(orc, ctx, case):
    N, d, Q, k = case
    db = _db(N, d, 61)
    rows = np.random.default_rng(62).choice(N, size=Q // 2, replace=False)
    q = np.concatenate([db[rows], _db(Q - Q // 2, d, 63)])
    want_ids, want_sc = orc.search_cosine(db, q, k)
    ctx.db_set(db)
    ids, sc = ctx.search_cosine(q, k)
    np.testing.assert_array_equal(ids, want_ids)
    assert_bitexact(sc, want_sc, "scores")


def test_kmeans_config5_shape_exact(orc, ctx):
    """k = 1024 centroids over d = 256 rows (BASELINE configs[4]), 2 iterations."""
    N, d, k = 20000, 256, 1024
    x = _db(N, d, 64)
    init = _init(k, d, seed=65)
    want_c, want_t, want_l = orc.kmeans(x, k, 2, init)
    ctx.db_set(x)
    cen, tot, lab = ctx.kmeans(k, 2, init)
    np.testing.assert_array_equal(lab, want_l)
    assert_bitexact(tot, want_t, "total counts")
    assert_bitexact(cen, want_c, "centroids")


def test_kmeans_tc_large_k_adversarial(orc, ctx):
    """k > 32 runs on the tensor cores (kmeans_tc.cuh): duplicate centroids, a NaN centroid, NaN / zero / huge rows, a tight cluster
    of rows -- labels, counts and centroids stay bit-exact (near-ties and special values go to the exact chains), and the
    fmaf-chain kernels alone (kmeans_tc = 0) give the same bits."""
    rng = np.random.default_rng(66)
    N, d, k = 12000, 100, 150
    x = rng.standard_normal(size=(N, d), dtype=np.float32)
    x[:3000] = x[1] + np.float32(1e-3) * rng.standard_normal(size=(3000, d), dtype=np.float32)
    x[3000:5000] *= np.exp(rng.uniform(-12, 12, size=(2000, 1))).astype(np.float32)
    x[7000] = 0.0
    x[7001, 5] = np.nan
    x[7002] = 1e12
    init = _init(k, d, seed=67)
    init[77] = init[3]; init[149] = init[3]                       # exact ties: lowest index
    for nan_centroid in (False, True):
        ini = init.copy()
        if nan_centroid:
            ini[100, 7] = np.nan                                  # TH's max scan: the first NaN wins every row
        want_c, want_t, want_l = orc.kmeans(x, k, 2, ini)
        ctx.db_set(x)
        for tc in (1, 0):
            ctx.set_option("kmeans_tc", tc)
            cen, tot, lab = ctx.kmeans(k, 2, ini)
            np.testing.assert_array_equal(lab, want_l)
            assert_bitexact(tot, want_t, "total counts")
            assert_bitexact(cen, want_c, "centroids")
    ctx.set_option("kmeans_tc", 1)
    big = x.copy(); big[7002] = 1e30                              # |x| * N >= 2^62: the fixed-point sums cannot hold it -> refused, not wrapped
    ctx.db_set(big)
    with pytest.raises(Exception):
        ctx.kmeans(k, 1, init)


def test_db_aliases_resident_attrs(pkg, orc, ctx):
    """ganrev_db_set(NULL): the database IS the resident ATTRS0 buffer (no copy); overwriting ATTRS0 un-sets it."""
    C, H, W, nd, N = 1, 32, 32, 32, 300
    c = pkg.Context(0)
    try:
        c.load_G(C, H, W, nd, pkg.weights.init_G(C, H, W, nd, stress=True))
        c.load_R(0, C, H, W, nd, pkg.weights.init_R(C, H, W, nd, stress=True))
        img = c.forward_G(np.random.default_rng(1).normal(size=(N, nd)).astype(np.float32))
        att = c.forward_R(0, img)
        c.db_set(None, N=N, d=nd)
        ids, sc = c.search_rows(np.array([5, 17], np.int64), 10)
        want_ids, want_sc = orc.search_cosine(att, att[[5, 17]], 10)
        np.testing.assert_array_equal(ids, want_ids)
        assert_bitexact(sc, want_sc)
        c.forward_R(0, img[:10])                                  # ATTRS0 overwritten
        with pytest.raises(pkg.GanrevError):
            c.search_rows(np.array([5], np.int64), 10)
        c.db_set(att)                                             # an own copy survives
        c.forward_R(0, img[:10])
        ids2, _ = c.search_rows(np.array([5, 17], np.int64), 10)
        np.testing.assert_array_equal(ids2, want_ids)
    finally:
        c.close()


def test_search_rows_equals_search_by_vector(orc, ctx):
    """apply_r.lua:268: needles are rows i*100 of the searched tensor."""
    db = _db(10000, 32, 5)
    rows = np.array([99, 199, 299, 399, 499], np.int64)
    ctx.db_set(db)
    ids_r, sc_r = ctx.search_rows(rows, 100)
    ids_v, sc_v = ctx.search_cosine(db[rows], 100)
    np.testing.assert_array_equal(ids_r, ids_v)
    assert_bitexact(sc_r, sc_v)
    want_ids, want_sc = orc.search_cosine(db, db[rows], 100)
    np.testing.assert_array_equal(ids_r, want_ids)
    assert_bitexact(sc_r, want_sc)
    assert (ids_r[:, 0] == rows).all()
    big = np.arange(0, 7000, 100, dtype=np.int64)            # 70 needles: the wide kernel
    ids_b, sc_b = ctx.search_rows(big, 20)
    want_ids, want_sc = orc.search_cosine(db, db[big], 20)
    np.testing.assert_array_equal(ids_b, want_ids)
    assert_bitexact(sc_b, want_sc)


def test_search_full_size_properties(ctx):
    """BASELINE config 4 size (1M x 100, 4096 queries, top-20): too big for the CPU oracle, so check
    size-independent properties: every needle finds itself first with score ~1, scores are sorted,
    ids are unique and in range, and splitting the queries differently gives the identical answer."""
    rng = np.random.default_rng(31)
    N, d, Q, k = 1_000_000, 100, 4096, 20
    db = rng.standard_normal(size=(N, d), dtype=np.float32)
    rows = (np.arange(1, Q + 1, dtype=np.int64) * 244)
    ctx.db_set(db)
    ids, sc = ctx.search_rows(rows, k)
    assert (ids[:, 0] == rows).all() and np.abs(sc[:, 0] - 1.0).max() < 1e-5
    assert (np.diff(sc, axis=1) <= 0).all()
    assert ids.min() >= 0 and ids.max() < N
    assert all(len(set(r.tolist())) == k for r in ids[::97])
    ids2, sc2 = ctx.search_cosine(db[rows[:70]], k)           # different query tiling / split count
    np.testing.assert_array_equal(ids2, ids[:70])
    assert_bitexact(sc2, sc[:70])
    # spot-check 3 needles against a float64 brute force
    for qi in (0, 2047, 4095):
        ref = (db @ db[rows[qi]].astype(np.float64)) / (np.linalg.norm(db, axis=1).astype(np.float64) * np.linalg.norm(db[rows[qi]].astype(np.float64)))
        top = np.argsort(-ref)[:k]
        assert len(set(top[:15].tolist()) - set(ids[qi].tolist())) == 0
        assert np.abs(sc[qi] - ref[ids[qi]]).max() < 1e-5


def test_search_ties_nan_zero(orc, ctx):
    """Duplicate rows (exact score ties -> lowest id first), zero rows, NaN rows (sorted last)."""
    rng = np.random.default_rng(7)
    base = rng.normal(size=(40, 16)).astype(np.float32)
    db = base[rng.integers(0, 40, size=4000)]                 # heavy duplication
    db[5] = 0.0
    db[100] = np.nan
    db[101, 3] = np.nan
    q = np.concatenate([base[:6], np.zeros((1, 16), np.float32)])
    for k in (1, 20, 128):
        want_ids, want_sc = orc.search_cosine(db, q, k)
        ctx.db_set(db)
        ids, sc = ctx.search_cosine(q, k)
        np.testing.assert_array_equal(ids, want_ids)
        assert_bitexact(sc, want_sc, f"scores k={k}")
    # every NaN row sorts after every finite row when k covers the whole db
    small = db[:120]
    want_ids, _ = orc.search_cosine(small, q[:2], 120)
    ctx.db_set(small)
    ids, sc = ctx.search_cosine(q[:2], 120)
    np.testing.assert_array_equal(ids, want_ids)
    assert set(ids[0, -2:].tolist()) == {100, 101} and np.isnan(sc[0, -2:]).all()


def test_cosine_pair(orc, ctx):
    rng = np.random.default_rng(3)
    for d in (1, 32, 100, 1024):
        a, b = rng.normal(size=d).astype(np.float32), rng.normal(size=d).astype(np.float32)
        assert np.float32(ctx.cosine(a, b)).view(np.uint32) == np.float32(orc.cosine(a, b)).view(np.uint32)


KMEANS_CASES = [
    # N, d, k, niter
    (10000, 32, 20, 15),     # apply_r.lua:159-161
    (4000, 100, 70, 4),      # several centroid tiles
    (3000, 30, 7, 5),        # k <= 16 path, ragged d
    (129, 8, 3, 3),
    (6000, 256, 33, 2),      # k*d too big for the shared-memory accumulators
]


def _init(k, d, seed=6):
    c = np.random.default_rng(seed).normal(size=(k, d)).astype(np.float32)
    return (c / np.linalg.norm(c, axis=1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("case", KMEANS_CASES, ids=lambda c: "N%d_d%d_k%d_it%d" % c)
def test_kmeans_exact(orc, ctx, case):
    N, d, k, niter = case
    x = _db(N, d, 11)
    init = _init(k, d)
    want_c, want_t, want_l = orc.kmeans(x, k, niter, init)
    ctx.db_set(x)
    cen, tot, lab = ctx.kmeans(k, niter, init)
    np.testing.assert_array_equal(lab, want_l)
    assert_bitexact(tot, want_t, "total counts")
    assert_bitexact(cen, want_c, "centroids")


def test_kmeans_full_size_properties(ctx):
    """1M x 100 rows, k=20 (the shape of BASELINE configs 3/4): too slow for the oracle's sort-free
    checks at this size, so test what must hold at any size: counts add up, the final centroids are the
    float64 means of the rows labelled in the last iteration, and -- because the centroid sums are
    int64 fixed point -- a row permutation changes neither centroids nor labels by a single bit."""
    rng = np.random.default_rng(41)
    N, d, k, niter = 1_000_000, 100, 20, 3
    x = rng.standard_normal(size=(N, d), dtype=np.float32)
    x += (3.0 * rng.integers(0, 2, size=(N, 1))).astype(np.float32)
    init = _init(k, d, 42)
    ctx.db_set(x)
    cen, tot, lab = ctx.kmeans(k, niter, init)
    assert tot.sum() == niter * N and lab.min() >= 0 and lab.max() < k
    for j in range(k):
        m = lab == j
        if m.any():
            assert np.abs(cen[j] - x[m].astype(np.float64).mean(0)).max() < 2e-5
    perm = rng.permutation(N)
    ctx.db_set(x[perm])
    cen_p, tot_p, lab_p = ctx.kmeans(k, niter, init)
    assert_bitexact(cen_p, cen, "centroids under a row permutation")
    assert_bitexact(tot_p, tot)
    np.testing.assert_array_equal(lab_p, lab[perm])
    # cosine-min assignment at the same size against float64 (ties / rounding aside)
    cl, cv = ctx.assign_cosine_min(cen)
    xs = x[perm][:20000].astype(np.float64)
    ref = (xs / np.linalg.norm(xs, axis=1, keepdims=True)) @ (cen / np.linalg.norm(cen, axis=1, keepdims=True)).astype(np.float64).T
    assert (ref.argmin(1) == cl[:20000]).mean() > 0.9995
    assert np.abs(ref.min(1) - cv[:20000]).max() < 1e-5


def test_kmeans_empty_cluster_and_zero_iters(orc, ctx):
    x = _db(500, 16, 12)
    init = _init(5, 16)
    init[3] = -1000.0 * np.abs(init[3]) - 1000.0     # |c|^2 term makes it lose everywhere: stays empty
    want_c, want_t, want_l = orc.kmeans(x, 5, 4, init)
    ctx.db_set(x)
    cen, tot, lab = ctx.kmeans(5, 4, init)
    assert want_t[3] == 0 and (want_c[3] == init[3]).all(), "empty clusters keep their centroid"
    np.testing.assert_array_equal(lab, want_l)
    assert_bitexact(cen, want_c)
    assert_bitexact(tot, want_t)
    cen0, tot0, lab0 = ctx.kmeans(5, 0, init)
    assert_bitexact(cen0, init)
    assert (tot0 == 0).all() and (lab0 == -1).all()


@pytest.mark.parametrize("case", [(10000, 32, 20), (3000, 100, 70), (500, 30, 3)], ids=lambda c: "N%d_d%d_k%d" % c)
def test_assign_cosine_min_and_members_exact(orc, ctx, case):
    N, d, k = case
    x = _db(N, d, 21)
    cen = _db(k, d, 22)
    x[7] = cen[0]                                    # exact duplicates -> ties
    x[8] = cen[0]
    images = np.random.default_rng(23).random((N, 48)).astype(np.float32)
    want_cl, want_cv = orc.assign_cosine_min(x, cen)
    ctx.db_set(x)
    cl, cv = ctx.assign_cosine_min(cen)
    np.testing.assert_array_equal(cl, want_cl)
    assert_bitexact(cv, want_cv, "cos")
    m = 71                                            # nbMaxPerCluster = 64+7, apply_r.lua:161
    want_ids, want_cnt, want_mean = orc.cluster_members(want_cl, want_cv, k, m, images)
    ids, cnt, mean = ctx.cluster_members(k, m, images)
    np.testing.assert_array_equal(cnt, want_cnt)
    np.testing.assert_array_equal(ids, want_ids)
    ok = want_cnt > 0                                 # empty clusters: 0/0 = NaN on both sides
    assert_bitexact(mean[ok], want_mean[ok], "mean images")
    assert np.isnan(mean[~ok]).all()


RTILE_CASES = [
    # N, d, k : the register-tiled path (9 <= k <= 32, d % 4 == 0, d <= 128) at its boundaries
    (5000, 100, 9), (5000, 100, 16), (5000, 100, 17), (4097, 100, 24), (5000, 100, 25), (5000, 100, 32),
    (1000, 4, 20), (1000, 128, 20), (127, 32, 20), (128, 32, 20), (129, 32, 20), (1, 32, 20),
]


# (stream_tc, label_tc, rtile): the TMA -> tf32 filter pipeline (default), the split-bf16 labelling kernel, the register-tiled
# chains, the one-thread-per-row streaming chains
PATHS = ((1, 0, 1), (0, 1, 1), (0, 0, 1), (0, 0, 0))


def _default_paths(ctx):
    ctx.set_option("stream_tc", 1)
    ctx.set_option("label_tc", 0)
    ctx.set_option("rtile", 1)


@pytest.mark.parametrize("case", RTILE_CASES, ids=lambda c: "N%d_d%d_k%d" % c)
def test_rtile_kmeans_and_assign_exact(orc, ctx, case):
    """kmeans labels/centroids/counts and cosine-min labels/values: bit-exact against the oracle, and identical with the
    register-tiled kernels switched off (the one-thread-per-row streaming kernels)."""
    N, d, k = case
    x = _db(N, d, 31)
    init = _init(k, d, seed=32)
    want_c, want_t, want_l = orc.kmeans(x, k, 3, init)
    ctx.db_set(x)
    for stream_tc, label_tc, rtile in PATHS:
        ctx.set_option("stream_tc", stream_tc)
        ctx.set_option("label_tc", label_tc)
        ctx.set_option("rtile", rtile)
        cen, tot, lab = ctx.kmeans(k, 3, init)
        np.testing.assert_array_equal(lab, want_l)
        assert_bitexact(tot, want_t, "total counts")
        assert_bitexact(cen, want_c, "centroids")
    _default_paths(ctx)
    cenq = _db(k, d, 33)
    if N > 9:
        x2 = x.copy(); x2[7] = cenq[0]; x2[8] = cenq[k - 1]          # exact duplicates of the first / last centroid
        ctx.db_set(x2)
    else:
        x2 = x
    want_cl, want_cv = orc.assign_cosine_min(x2, cenq)
    for stream_tc, label_tc, rtile in PATHS:
        ctx.set_option("stream_tc", stream_tc)
        ctx.set_option("label_tc", label_tc)
        ctx.set_option("rtile", rtile)
        cl, cv = ctx.assign_cosine_min(cenq)
        np.testing.assert_array_equal(cl, want_cl)
        assert_bitexact(cv, want_cv, "cos")
    _default_paths(ctx)


def test_rtile_nan_and_tie_rules(orc, ctx):
    """TH max scan (first NaN wins, else first maximum) and cosine-min (a NaN at j = 0 sticks, other NaNs never win,
    lowest j on ties) across the warp-merge of the register-tiled kernels."""
    rng = np.random.default_rng(41)
    N, d, k = 1500, 32, 20
    x = rng.normal(size=(N, d)).astype(np.float32)
    cen = rng.normal(size=(k, d)).astype(np.float32)
    cen[5] = cen[13]                                                   # two identical centroids in different warps: lowest j
    cen[17, 3] = np.nan                                                # a NaN centroid in the last warp
    x[100, 0] = np.nan                                                 # a NaN row: every score NaN
    x[200] = 0.0                                                       # zero row: every cosine 0 -> j = 0
    init = cen.copy()
    want_c, want_t, want_l = orc.kmeans(x, k, 2, init)
    ctx.db_set(x)
    c, t, l = ctx.kmeans(k, 2, init)
    np.testing.assert_array_equal(l, want_l)
    assert_bitexact(t, want_t); assert_bitexact(c, want_c)
    want_cl, want_cv = orc.assign_cosine_min(x, cen)
    cl, cv = ctx.assign_cosine_min(cen)
    np.testing.assert_array_equal(cl, want_cl)
    assert_bitexact(cv, want_cv)
    cen0 = cen.copy(); cen0[0, 1] = np.nan                            # NaN at j = 0 sticks for cosine-min
    want_cl, want_cv = orc.assign_cosine_min(x, cen0)
    cl, cv = ctx.assign_cosine_min(cen0)
    np.testing.assert_array_equal(cl, want_cl)
    assert_bitexact(cv, want_cv)


def test_label_tc_adversarial(orc, ctx):
    """Tensor-core labelling under stress: duplicate centroids (every row a tie), rows equal to centroids, clustered data with tiny
    gaps, a huge dynamic range, a zero row and a zero centroid -- labels, centroids, counts and cosines stay bit-exact because every
    near-tie goes to the exact chains."""
    rng = np.random.default_rng(55)
    N, d, k = 20000, 100, 20
    x = rng.standard_normal(size=(N, d), dtype=np.float32)
    x[:5000] = x[0] + np.float32(1e-4) * rng.standard_normal(size=(5000, d), dtype=np.float32)   # a tight cluster
    x[5000:10000] *= np.exp(rng.uniform(-15, 15, size=(5000, 1))).astype(np.float32)
    x[12345] = 0.0
    init = _init(k, d, seed=56)
    init[7] = init[3]                                             # duplicate centroids: exact ties, lowest index wins
    init[9] = 0.0                                                 # a zero centroid
    x[100] = init[5]; x[101] = init[19]
    want_c, want_t, want_l = orc.kmeans(x, k, 3, init)
    want_cl, want_cv = orc.assign_cosine_min(x, init)
    ctx.db_set(x)
    for stream_tc in (1, 0):                                      # the tf32 pipeline, then the split-bf16 labelling kernel
        ctx.set_option("stream_tc", stream_tc)
        ctx.set_option("label_tc", 1)
        cen, tot, lab = ctx.kmeans(k, 3, init)
        np.testing.assert_array_equal(lab, want_l)
        assert_bitexact(tot, want_t); assert_bitexact(cen, want_c)
        cl, cv = ctx.assign_cosine_min(init)
        np.testing.assert_array_equal(cl, want_cl)
        assert_bitexact(cv, want_cv)
    for kk, dd in ((1, 32), (2, 4), (32, 128), (31, 64), (20, 68)):   # shapes at the limits of the kernel
        xs = rng.standard_normal(size=(3001, dd), dtype=np.float32)
        ini = _init(kk, dd, seed=57)
        w_c, w_t, w_l = orc.kmeans(xs, kk, 2, ini)
        ctx.db_set(xs)
        c2, t2, l2 = ctx.kmeans(kk, 2, ini)
        np.testing.assert_array_equal(l2, w_l)
        assert_bitexact(t2, w_t); assert_bitexact(c2, w_c)
        w_cl, w_cv = orc.assign_cosine_min(xs, ini)
        cl2, cv2 = ctx.assign_cosine_min(ini)
        np.testing.assert_array_equal(cl2, w_cl)
        assert_bitexact(cv2, w_cv)
    _default_paths(ctx)


def test_search_four_needles_variant(orc, ctx):
    """Q <= 4 takes the unpadded 2-queries-per-thread streaming variant: exact at Q = 1..4, ragged N."""
    x = _db(3001, 100, 51)
    ctx.db_set(x)
    for stream_tc in (0, 1):
        ctx.set_option("stream_tc", stream_tc)
        for Q in (1, 2, 3, 4):
            q = x[[5, 77, 1234, 3000][:Q]] + 0.01
            ids, sc = ctx.search_cosine(q, 20)
            wi, ws = orc.search_cosine(x, q, 20)
            np.testing.assert_array_equal(ids, wi)
            assert_bitexact(sc, ws)


def test_db_error_paths(pkg):
    """Misuse fails with an error code and a message, never with a crash or a silent fallback."""
    c = pkg.Context(0)
    try:
        q = np.zeros((2, 8), np.float32)
        with pytest.raises(pkg.GanrevError):
            c.search_cosine(q, 5)                                   # no database
        with pytest.raises(pkg.GanrevError):
            c.kmeans(3, 1, np.zeros((3, 8), np.float32))            # no database
        c.db_set(np.random.default_rng(0).normal(size=(50, 8)).astype(np.float32))
        with pytest.raises(pkg.GanrevError):
            c.search_cosine(q, 129)                                 # k <= 128
        with pytest.raises(pkg.GanrevError):
            c.cluster_members(3, 5, np.zeros((50, 4), np.float32))  # assign_cosine_min has not run
        with pytest.raises(pkg.GanrevError):
            c.nearest_l2(np.zeros((1, 16), np.float32), None, N=10)  # no resident images
        ids, sc = c.search_cosine(q, 5)                             # zero queries: every cosine is 0 -> lowest ids
        assert ids.tolist() == [[0, 1, 2, 3, 4]] * 2 and np.all(sc == 0)
    finally:
        c.close()
