"""CPU suite: the C oracle against (a) the committed golden vectors (tests/golden/golden.npz,
made by tests/golden/make_golden.py from PyTorch-CPU / numpy statements independent of the
oracle) and (b) live PyTorch-CPU / numpy restatements.  Parity with Torch7 itself is UNPINNED."""
import math
import os

import numpy as np
import pytest

import torch_ref
from util import assert_bitexact

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden.npz"))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_networks_against_golden(pkg, orc, tag):
    C, H, W, nd, N = (int(v) for v in GOLD[f"net_{tag}_geom"])
    gb = pkg.weights.init_G(C, H, W, nd, seed=101, stress=True)
    rb = pkg.weights.init_R(C, H, W, nd, seed=102, stress=True)
    if abs(gb.astype(np.float64).sum() - float(GOLD[f"net_{tag}_gsum"])) > 1e-9:
        pytest.skip("numpy Generator stream changed: regenerate tests/golden/golden.npz")
    img = orc.forward_G(gb, C, H, W, nd, GOLD[f"net_{tag}_noise"])
    assert np.abs(img - GOLD[f"net_{tag}_img"]).max() < 1e-4
    att = orc.forward_R(rb, C, H, W, nd, GOLD[f"net_{tag}_img"])
    assert np.abs(att - GOLD[f"net_{tag}_att"]).max() < 1e-3
    attm = orc.forward_R(rb, C, H, W, nd, GOLD[f"net_{tag}_img"], GOLD[f"net_{tag}_mask"])
    assert np.abs(attm - GOLD[f"net_{tag}_attm"]).max() < 1e-3
    assert np.abs(attm - att).max() > 1e-2


def test_exact_paths_against_golden(orc):
    ids, sc = orc.search_cosine(GOLD["search_db"], GOLD["search_q"], 12)
    np.testing.assert_array_equal(ids, GOLD["search_ids"])
    assert_bitexact(sc, GOLD["search_sc"], "scores")
    x, init = GOLD["km_x"], GOLD["km_init"]
    assert orc.kmeans_shift(x) == int(GOLD["km_shift"])
    cen, tot, lab = orc.kmeans(x, 4, 3, init)
    assert_bitexact(cen, GOLD["km_cen"]); assert_bitexact(tot, GOLD["km_tot"])
    np.testing.assert_array_equal(lab, GOLD["km_lab"])
    cl, cv = orc.assign_cosine_min(x, GOLD["km_cen"])
    np.testing.assert_array_equal(cl, GOLD["km_cl"]); assert_bitexact(cv, GOLD["km_cv"])
    d = orc.l2(GOLD["l2_a"], GOLD["l2_b"])
    assert_bitexact(d, GOLD["l2_d"])
    flags, thr = orc.anomaly_flags(d, 5, 5, float(GOLD["flag_q"]))
    assert thr == float(GOLD["flag_thr"])
    np.testing.assert_array_equal(flags, GOLD["flag_flags"])


@pytest.mark.parametrize("geom", [(1, 32, 32, 32, 4), (3, 32, 32, 20, 2), (1, 16, 16, 100, 3)])
@pytest.mark.parametrize("stress", [False, True])
def test_networks_against_torch_cpu(pkg, orc, geom, stress):
    C, H, W, nd, N = geom
    Wt = pkg.weights
    gb, rb = Wt.init_G(C, H, W, nd, 1, stress), Wt.init_R(C, H, W, nd, 2, stress)
    assert gb.size == orc.blob_floats_G(C, H, W, nd) == Wt.blob_floats(Wt.g_layout(C, H, W, nd))
    assert rb.size == orc.blob_floats_R(C, H, W, nd) == Wt.blob_floats(Wt.r_layout(C, H, W, nd))
    noise = np.random.default_rng(3).normal(size=(N, nd)).astype(np.float32)
    img = orc.forward_G(gb, C, H, W, nd, noise)
    img_t = torch_ref.forward_G(Wt.unpack(gb, Wt.g_layout(C, H, W, nd)), C, H, W, nd, noise)
    assert np.abs(img - img_t).max() < 1e-5
    assert img.min() >= 0.0 and img.max() <= 1.0
    mask = (np.random.default_rng(5).random(img.shape) >= 0.5).astype(np.uint8)
    for m in (None, mask):
        for tanh in (False, True):
            att = orc.forward_R(rb, C, H, W, nd, img, m, tanh_out=tanh)
            att_t = torch_ref.forward_R(Wt.unpack(rb, Wt.r_layout(C, H, W, nd)), C, H, W, nd, img, m, tanh_out=tanh)
            assert np.abs(att - att_t).max() < 1e-4 * max(1.0, np.abs(att_t).max())


def test_cosine_matches_double_precision(orc):
    rng = np.random.default_rng(0)
    for d in (1, 3, 32, 100, 1024):
        a, b = rng.normal(size=d).astype(np.float32), rng.normal(size=d).astype(np.float32)
        want = float(a.astype(np.float64) @ b.astype(np.float64)) / math.sqrt(float((a.astype(np.float64) ** 2).sum()) * float((b.astype(np.float64) ** 2).sum()))
        assert abs(orc.cosine(a, b) - want) < 1e-5
    z = np.zeros(8, np.float32)
    assert orc.cosine(z, z) == 0.0      # eps keeps 0/0 finite, as nn.CosineDistance does


def test_search_against_numpy(orc):
    rng = np.random.default_rng(1)
    db = rng.normal(size=(3000, 32)).astype(np.float32)
    q = db[[99, 199, 299, 399, 499]]                      # needles i*100, apply_r.lua:268 (0-based here)
    ids, sc = orc.search_cosine(db, q, 100)
    dn = db / np.linalg.norm(db, axis=1, keepdims=True)
    ref = (dn[[99, 199, 299, 399, 499]].astype(np.float64) @ dn.T.astype(np.float64))
    assert (ids[:, 0] == [99, 199, 299, 399, 499]).all(), "a row is its own nearest neighbour"
    for r in range(5):
        assert np.all(np.diff(sc[r]) <= 0)
        top = set(np.argsort(-ref[r])[:90].tolist())
        assert len(top - set(ids[r].tolist())) == 0
        assert np.abs(sc[r] - ref[r][ids[r]]).max() < 1e-5
    ids2, _ = orc.search_cosine(db[:7], q, 10)
    assert (ids2[:, 7:] == -1).all()


def test_kmeans_is_lloyd_and_partition_free(orc):
    rng = np.random.default_rng(2)
    x = (rng.normal(size=(2000, 16)) + 3.0 * rng.integers(0, 2, size=(2000, 1))).astype(np.float32)
    init = rng.normal(size=(5, 16)).astype(np.float32)
    init /= np.linalg.norm(init, axis=1, keepdims=True)
    cen, tot, lab = orc.kmeans(x, 5, 1, init)
    # one Lloyd step in float64: argmax c.x - |c|^2/2 == argmin |x - c|^2
    d2 = ((x[:, None, :].astype(np.float64) - init[None].astype(np.float64)) ** 2).sum(-1)
    ref_lab = d2.argmin(1)
    agree = (ref_lab == lab).mean()
    assert agree > 0.999
    for j in range(5):
        if (lab == j).any():
            assert np.abs(cen[j] - x[lab == j].astype(np.float64).mean(0)).max() < 1e-5
        else:
            assert (cen[j] == init[j]).all()
    assert tot.sum() == 2000
    # the fixed-point sums make the result independent of row order
    perm = rng.permutation(2000)
    shift = orc.kmeans_shift(x)
    cen_p, tot_p, lab_p = orc.kmeans(x[perm], 5, 6, init, shift)
    cen_o, tot_o, lab_o = orc.kmeans(x, 5, 6, init, shift)
    assert_bitexact(cen_p, cen_o); assert_bitexact(tot_p, tot_o)
    np.testing.assert_array_equal(lab_p, lab_o[perm])


def test_assign_min_members_l2_flags(orc):
    rng = np.random.default_rng(3)
    x = rng.normal(size=(400, 8)).astype(np.float32)
    cen = rng.normal(size=(6, 8)).astype(np.float32)
    cl, cv = orc.assign_cosine_min(x, cen)
    xn = x / np.linalg.norm(x, axis=1, keepdims=True)
    cn = cen / np.linalg.norm(cen, axis=1, keepdims=True)
    ref = xn.astype(np.float64) @ cn.T.astype(np.float64)
    assert (ref.argmin(1) == cl).mean() > 0.99              # the MINIMUM cosine, apply_r.lua:211
    imgs = rng.random((400, 12)).astype(np.float32)
    ids, cnt, mean = orc.cluster_members(cl, cv, 6, 71, imgs)
    for j in range(6):
        members = np.where(cl == j)[0]
        assert cnt[j] == min(71, members.size)
        got = ids[j, :cnt[j]]
        assert np.all(np.diff(cv[got]) <= 0)                 # sorted descending, apply_r.lua:224
        if cnt[j]:
            assert np.abs(mean[j] - imgs[got].mean(0)).max() < 1e-5
    a = rng.random((9, 1024)).astype(np.float32)
    b = rng.random((9, 1024)).astype(np.float32)
    d = orc.l2(a, b)
    assert np.abs(d - np.sqrt(((a.astype(np.float64) - b) ** 2).sum(1))).max() < 1e-4
    assert np.all(np.abs(d - orc.l2(a, b, sequential=True)) <= 4 * np.spacing(d))
    flags, thr = orc.anomaly_flags(d, 9, 5, 0.5)
    sims = 1.0 - d
    assert thr == np.sort(sims)[math.floor(9 * 0.5) - 1]
    np.testing.assert_array_equal(flags.astype(bool), sims[:5] <= thr)
    with pytest.raises(RuntimeError):
        orc.anomaly_flags(d, 5, 5, 0.15)


def test_nearest_l2_oracle(orc):
    """sample.lua:128-148: first strictly smallest torch.dist; ties -> lowest row; row-0 NaN sticks; empty set."""
    rng = np.random.default_rng(11)
    ts = rng.random((300, 48)).astype(np.float32)
    q = (ts[[5, 250, 17]] + rng.normal(scale=0.01, size=(3, 48))).astype(np.float32)
    ids, dist = orc.nearest_l2(q, ts)
    ref = np.sqrt(((ts[None].astype(np.float64) - q[:, None].astype(np.float64)) ** 2).sum(-1))
    np.testing.assert_array_equal(ids, ref.argmin(1))
    np.testing.assert_allclose(dist, ref.min(1), rtol=1e-6)
    ts2 = ts.copy(); ts2[40] = ts2[7]; ts2[200] = ts2[7]                # exact duplicates: the first one wins
    ids2, _ = orc.nearest_l2(ts2[[7]], ts2)
    assert ids2[0] == 7
    ts3 = ts.copy(); ts3[0, 3] = np.nan                                  # "closestDist == nil or dist < closestDist"
    ids3, d3 = orc.nearest_l2(q, ts3)
    assert np.all(ids3 == 0) and np.all(np.isnan(d3))
    ts4 = ts.copy(); ts4[9, 0] = np.nan                                  # NaN elsewhere is never taken
    ids4, _ = orc.nearest_l2(q, ts4)
    np.testing.assert_array_equal(ids4, ids)
    ids5, d5 = orc.nearest_l2(q, np.zeros((0, 48), np.float32))
    assert np.all(ids5 == -1) and np.all(np.isinf(d5))
