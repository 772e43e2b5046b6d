"""GPU parity of torch.dist (apply_r.lua:366) and the quantile flags (apply_r.lua:370-378):
bit-exact against the oracle's canonical order, and within 4 ulp of TH's sequential order."""
import numpy as np
import pytest

from util import assert_bitexact

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("N,px", [(1024, 1024), (528, 3072), (40, 12288), (33, 1000), (17, 37), (3, 1)])
def test_l2_exact(orc, ctx, N, px):
    rng = np.random.default_rng(N + px)
    a = rng.random((N, px)).astype(np.float32)
    b = (a + rng.normal(scale=0.05, size=(N, px))).astype(np.float32)
    b[0] = a[0]                                       # zero distance
    got = ctx.l2(a, b)
    assert_bitexact(got, orc.l2(a, b), "l2")
    seq = orc.l2(a, b, sequential=True)
    assert np.all(np.abs(got - seq) <= 4 * np.spacing(seq)), "canonical order stays within 4 ulp of TH's order"


@pytest.mark.parametrize("n_calc,n_show,q", [(1024, 528, 0.15), (10000, 10000, 0.15), (7, 7, 0.15), (100, 10, 0.999)])
def test_anomaly_flags_exact(orc, ctx, n_calc, n_show, q):
    rng = np.random.default_rng(n_calc)
    l2 = np.abs(rng.normal(size=n_calc)) * 3.0
    l2[: n_calc // 10] = l2[n_calc // 10: 2 * (n_calc // 10)]     # duplicates around the threshold
    want_flags, want_thr = orc.anomaly_flags(l2, n_calc, n_show, q)
    flags, thr = ctx.anomaly_flags(l2, n_calc, n_show, q)
    assert np.float64(thr).view(np.uint64) == np.float64(want_thr).view(np.uint64)
    np.testing.assert_array_equal(flags, want_flags)


def test_anomaly_flags_full_size(ctx):
    """1M distances (BASELINE configs[3] size): the multi-block radix select returns the exact order statistic."""
    rng = np.random.default_rng(77)
    n = 1_000_000
    l2 = np.abs(rng.normal(size=n)) * 2.0
    l2[::1000] = l2[7]                                              # a run of duplicates
    flags, thr = ctx.anomaly_flags(l2, n, n, 0.15)
    sims = 1.0 - l2
    want_thr = np.partition(sims, int(np.floor(n * 0.15)) - 1)[int(np.floor(n * 0.15)) - 1]
    assert np.float64(thr).view(np.uint64) == np.float64(want_thr).view(np.uint64)
    np.testing.assert_array_equal(flags, (sims <= want_thr).astype(np.uint8))


def test_anomaly_flags_resident_distances(pkg, orc, ctx):
    """l2 == NULL reuses the distances of the last l2 / fix_l2 call; fewer than n_calc resident is an error."""
    rng = np.random.default_rng(3)
    a, b = rng.random((64, 100)).astype(np.float32), rng.random((64, 100)).astype(np.float32)
    l2 = ctx.l2(a, b)
    flags, thr = ctx.anomaly_flags(None, 64, 40, 0.15)
    want_flags, want_thr = orc.anomaly_flags(l2, 64, 40, 0.15)
    np.testing.assert_array_equal(flags, want_flags)
    assert np.float64(thr).view(np.uint64) == np.float64(want_thr).view(np.uint64)
    with pytest.raises(pkg.GanrevError):
        ctx.anomaly_flags(None, 65, 40, 0.15)


def test_anomaly_flags_bad_quantile(pkg, ctx):
    with pytest.raises(pkg.GanrevError):
        ctx.anomaly_flags(np.ones(5), 5, 5, 0.15)     # floor(5*0.15) = 0: Lua would index nil


@pytest.mark.parametrize("Q,N,px", [(3, 1000, 1024), (8, 4096, 1024), (11, 700, 3072), (1, 5, 37), (20, 257, 100)])
def test_nearest_l2_exact(orc, ctx, Q, N, px):
    """SURVEY 8f rank 3 (sample.lua:128-148): ids and distances bit-exact against the oracle."""
    rng = np.random.default_rng(Q * 7 + N)
    ts = rng.random((N, px)).astype(np.float32)
    q = (ts[rng.integers(0, N, size=Q)] + rng.normal(scale=0.02, size=(Q, px))).astype(np.float32)
    if N > 60:
        ts[50] = ts[3]; ts[N - 1] = ts[3]; q[0] = ts[3]                  # exact duplicates: distance 0 three times -> row 3
    oi, od = orc.nearest_l2(q, ts)
    for stream_tc in (1, 0):                                             # tensor-core filter + canonical candidates; every pair canonically
        ctx.set_option("stream_tc", stream_tc)
        ids, dist = ctx.nearest_l2(q, ts)
        np.testing.assert_array_equal(ids, oi)
        assert_bitexact(dist, od, "nearest_l2 distance")
    ctx.set_option("stream_tc", 1)


def test_nearest_l2_nan_and_empty(orc, ctx):
    rng = np.random.default_rng(5)
    ts = rng.random((600, 64)).astype(np.float32)
    q = rng.random((9, 64)).astype(np.float32)
    ts_a = ts.copy(); ts_a[0, 1] = np.nan                                # row 0 is taken unconditionally and then sticks
    ids, dist = ctx.nearest_l2(q, ts_a)
    assert np.all(ids == 0) and np.all(np.isnan(dist))
    ts_b = ts.copy(); ts_b[77, 5] = np.nan                               # NaN rows elsewhere are never taken
    ids_b, dist_b = ctx.nearest_l2(q, ts_b)
    oi, od = orc.nearest_l2(q, ts_b)
    np.testing.assert_array_equal(ids_b, oi)
    assert_bitexact(dist_b, od)
    ids_e, dist_e = ctx.nearest_l2(q, np.zeros((0, 64), np.float32))
    assert np.all(ids_e == -1) and np.all(np.isinf(dist_e))


def test_nearest_l2_resident_images_fullsize(pkg, orc, ctx):
    """200k resident 32x32 faces (819 MB): the nearest neighbour of a face that IS in the set is itself at distance 0,
    and a sampled check against the oracle on a slice."""
    rng = np.random.default_rng(9)
    N, px = 200000, 1024
    imgs = rng.random((N, px), dtype=np.float32)
    ctx.load_G(1, 32, 32, 100, pkg.weights.init_G(1, 32, 32, 100))     # resident buffers take their row shape from the loaded model
    ctx.buffer_put(pkg._lib.BUF_IMAGES, imgs.reshape(N, 1, 32, 32))
    rows = np.array([0, 1, 777, 123456, N - 1])
    ids, dist = ctx.nearest_l2(imgs[rows], None, N=N)
    np.testing.assert_array_equal(ids, rows)
    assert np.all(dist == 0.0)
    q = (imgs[rows] + rng.normal(scale=0.05, size=(5, px))).astype(np.float32)
    ids2, dist2 = ctx.nearest_l2(q, None, N=N)
    np.testing.assert_array_equal(ids2, rows)
    sl = slice(123000, 124000)
    oi, od = orc.nearest_l2(q[[3]], imgs[sl])
    assert oi[0] + 123000 == ids2[3]
    assert_bitexact(dist2[[3]], od)


def test_nearest_l2_filter_adversarial(orc, ctx):
    """The tensor-core filter of nearest_l2 may only pass extra rows: sets whose rows are all within the tf32 resolution of each
    other, many exact ties (lowest row wins), a huge dynamic range, zero rows / zero queries, inf and NaN entries away from row 0,
    and a set smaller than one tile."""
    rng = np.random.default_rng(17)
    px = 256
    base = rng.random((1, px)).astype(np.float32)
    ts = (base + np.float32(1e-6) * rng.standard_normal((9000, px))).astype(np.float32)       # distances differ in the 1e-5s
    ts[4000:6000] = rng.random((2000, px)).astype(np.float32) * np.exp(rng.uniform(-12, 12, size=(2000, 1))).astype(np.float32)
    ts[10] = ts[7]; ts[8999] = ts[7]; ts[123] = 0.0
    ts[200, 3] = np.inf; ts[201, 5] = np.nan
    q = np.stack([ts[7], base[0], ts[4100], np.zeros(px, np.float32), ts[5000] * np.float32(1.0001), ts[123] + np.float32(1e-20)]).astype(np.float32)
    for sub in (ts, ts[:100], ts[:129]):
        oi, od = orc.nearest_l2(q, sub)
        ids, dist = ctx.nearest_l2(q, sub)
        np.testing.assert_array_equal(ids, oi)
        assert_bitexact(dist, od)
    st = ctx.tfs_stats()
    assert st["launches"] >= 3
