"""Presentation row: grid layouts restated independently, borders, file writers."""
import math

import numpy as np
import pytest


@pytest.fixture(scope="module")
def pr(pkg):
    return pkg.present


def _grid_ref(x, nrow, padding, fill):
    """Independent restatement: place image k at cell (k // xmaps, k % xmaps)."""
    n, c, h, w = x.shape
    xmaps = min(nrow, n)
    ymaps = -(-n // xmaps)
    g = np.full((c, (h + padding) * ymaps, (w + padding) * xmaps), fill, np.float32)
    for k in range(n):
        r, q = divmod(k, xmaps)
        g[:, r * (h + padding) + padding // 2: r * (h + padding) + padding // 2 + h,
          q * (w + padding) + padding // 2: q * (w + padding) + padding // 2 + w] = x[k]
    return g


@pytest.mark.parametrize("n,nrow,padding", [(1, 6, 0), (7, 3, 0), (9, 3, 2), (10, 4, 0), (5, 8, 4)])
def test_toDisplayTensor_layout(pr, n, nrow, padding):
    x = np.random.default_rng(n).random((n, 3, 5, 4)).astype(np.float32)
    got = pr.toDisplayTensor(x, nrow=nrow, padding=padding, min=0, max=1.0)
    np.testing.assert_array_equal(got, np.clip(_grid_ref(x, nrow, padding, x.max()), 0, 1))
    # values outside [min, max] saturate; without min/max the grid is stretched to [0, 1]
    y = x * 3 - 1
    sat = pr.toDisplayTensor(y, nrow=nrow, padding=padding, min=0, max=1.0)
    assert sat.min() >= 0 and sat.max() <= 1
    auto = pr.toDisplayTensor(y, nrow=nrow, padding=padding)
    assert abs(auto.min()) < 1e-6 and abs(auto.max() - 1) < 1e-6


def test_toRgb(pr):
    y = np.random.default_rng(0).random((4, 1, 3, 3)).astype(np.float32)
    rgb = pr.toRgb(y, "y")
    assert rgb.shape == (4, 3, 3, 3) and all(np.array_equal(rgb[:, c], y[:, 0]) for c in range(3))
    assert pr.toRgb(rgb, "rgb") is rgb or np.array_equal(pr.toRgb(rgb, "rgb"), rgb)
    assert pr.toRgbSingle(y[0], "y").shape == (3, 3, 3)
    with pytest.raises(ValueError):
        pr.toRgb(rgb, "lab")                                  # nn_utils.lua:165: unknown colour space


def test_toRgb_hsl_yuv_known_answers(pr):
    """image.hsl2rgb / image.yuv2rgb behind NN_UTILS.toRgb (nn_utils.lua:152-163): primaries, greys and the analytic inverse."""
    hsl = np.zeros((1, 3, 1, 6), np.float32)
    #            red      green     blue     grey(s=0)  dark red   light cyan
    hsl[0, 0] = [0.0, 1.0 / 3.0, 2.0 / 3.0, 0.7,       0.0,       0.5]
    hsl[0, 1] = [1.0, 1.0,       1.0,       0.0,       1.0,       1.0]
    hsl[0, 2] = [0.5, 0.5,       0.5,       0.25,      0.25,      0.75]
    want = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0.25, 0.25, 0.25], [0.5, 0, 0], [0.5, 1, 1]], np.float32).T
    np.testing.assert_allclose(pr.toRgb(hsl, "hsl")[0, :, 0, :], want, atol=1e-6)
    rgb = np.random.default_rng(2).random((3, 3, 4, 5)).astype(np.float32)
    r, g, b = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    yuv = np.stack([0.299 * r + 0.587 * g + 0.114 * b, -0.14713 * r - 0.28886 * g + 0.436 * b, 0.615 * r - 0.51499 * g - 0.10001 * b], axis=1)
    np.testing.assert_allclose(pr.toRgb(yuv.astype(np.float32), "yuv"), rgb, atol=2e-4)
    assert pr.toRgbSingle(yuv[0].astype(np.float32), "yuv").shape == (3, 4, 5)


def test_apply_r_artefacts(pr, tmp_path):
    rng = np.random.default_rng(1)
    H = W = 8
    imgs = rng.random((30, 1, H, W)).astype(np.float32)
    # cluster grid: mean face first, nrow = ceil(sqrt(1 + members))   (apply_r.lua:247-256)
    members = imgs[[3, 5, 8, 13, 21]]
    g = pr.clusterGrid(members.mean(0), members, "y")
    assert g.shape == (3, 2 * H, 3 * W)                                   # 6 images, 3 per row
    np.testing.assert_allclose(g[:, :H, :W], np.tile(members.mean(0), (3, 1, 1)), atol=1e-7)
    np.testing.assert_array_equal(g[:, H:, 2 * W:], np.tile(members[4], (3, 1, 1)))
    # search grid: blue frame on the first image only   (apply_r.lua:284-295)
    s = pr.searchGrid(imgs[:4], "y")
    assert s.shape == (3, 2 * H, 2 * W)
    assert np.all(s[2, 0, :W] == 1) and np.all(s[0, 0, :W] == 0) and np.all(s[1, :H, 0] == 0) and np.all(s[2, :H, W - 1] == 1)
    np.testing.assert_array_equal(s[:, 1:H - 1, 1:W - 1], np.tile(imgs[0, :, 1:-1, 1:-1], (3, 1, 1)))
    np.testing.assert_array_equal(s[:, :H, W:], np.tile(imgs[1], (3, 1, 1)))
    # fixed pairs: blue background, original | fixed, 4 pairs per row   (apply_r.lua:325-345)
    p = pr.fixedPairsGrid(imgs[:6], imgs[6:12], "y")
    assert p.shape == (3, 2 * (H + 2), 4 * (2 * W + 2))
    assert np.all(p[2, 0, :2 * W + 2] == 1) and np.all(p[0, 0, :2 * W + 2] == 0)
    np.testing.assert_array_equal(p[:, 1:1 + H, 1:1 + W], np.tile(imgs[0], (3, 1, 1)))
    np.testing.assert_array_equal(p[:, 1:1 + H, 1 + W:1 + 2 * W], np.tile(imgs[6], (3, 1, 1)))
    # anomalies: red frame where flagged, floor(sqrt(n)) per row   (apply_r.lua:375-388)
    flags = np.zeros(9, bool); flags[[1, 4]] = True
    a = pr.anomalyGrid(imgs[:9], flags, "y")
    assert a.shape == (3, 3 * (H + 2), 3 * (W + 2))
    cell = a[:, :H + 2, W + 2:2 * (W + 2)]                                 # image 1: flagged
    assert np.all(cell[0, 0] == 1) and np.all(cell[1, 0] == 0) and np.all(cell[2, :, 0] == 0)
    assert np.all(a[:, 0, :W + 2] == 0)                                    # image 0: not flagged, black frame
    # writers: files appear with the reference's names and decode to the grid's size
    PIL = pytest.importorskip("PIL.Image")
    res = {"member_ids": np.array([[3, 5, -1], [-1, -1, -1]]), "member_counts": np.array([2, 0]),
           "average_faces": np.stack([imgs[[3, 5]].mean(0), imgs[0]])}
    out = pr.saveClusterImages(res, imgs, "y", str(tmp_path))
    assert [o.split("/")[-1] for o in out] == ["cluster_01.jpg"]           # the empty cluster writes nothing (apply_r.lua:248)
    assert PIL.open(out[0]).size == (2 * W, 2 * H)
    path = pr.saveAnomalies(imgs, flags, "y", str(tmp_path))
    assert path.endswith("anomalies.jpg") and PIL.open(path).size == (3 * (W + 2), 3 * (H + 2))
    b = pr.to_bytes(np.array([[[0.0, 0.5, 1.0, 2.0, -1.0]]], np.float32))
    assert b.reshape(-1).tolist() == [0, 127, 255, 255, 0]
