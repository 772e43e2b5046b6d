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


def test_anomaly_flags_bad_quantile(pkg, ctx):
    with pytest.raises(pkg.GanrevError):
        ctx.anomaly_flags(np.ones(5), 5, 5, 0.15)     # floor(5*0.15) = 0: Lua would index nil
