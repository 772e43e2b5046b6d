"""Torch7 checkpoint reader: known-answer bytes assembled by hand from the format
description, round trips through the test-only writer, and full G / R checkpoints -> weight blobs."""
import struct

import numpy as np
import pytest

import t7_writer as tw


@pytest.fixture(scope="module")
def t7(pkg):
    return pkg.t7


def test_known_answer_bytes(t7):
    i32 = lambda v: struct.pack("<i", v)
    assert t7.loads(i32(1) + struct.pack("<d", 1.5)) == 1.5
    assert t7.loads(i32(2) + i32(3) + b"abc") == "abc"
    assert t7.loads(i32(5) + i32(1)) is True and t7.loads(i32(0)) is None
    # {1 = "a", x = 7}: table, index 1, 2 pairs
    tbl = i32(3) + i32(1) + i32(2) + (i32(1) + struct.pack("<d", 1.0)) + (i32(2) + i32(1) + b"a") + (i32(2) + i32(1) + b"x") + (i32(1) + struct.pack("<d", 7.0))
    assert t7.loads(tbl) == {1: "a", "x": 7}
    # torch.FloatTensor 2x3 over a 6-element FloatStorage, versioned header, 8-byte longs
    q = lambda v: struct.pack("<q", v)
    s = lambda x: i32(len(x)) + x
    data = np.arange(6, dtype="<f4")
    ten = (i32(4) + i32(1) + s(b"V 1") + s(b"torch.FloatTensor") + i32(2) + q(2) + q(3) + q(3) + q(1) + q(1)
           + i32(4) + i32(2) + s(b"V 1") + s(b"torch.FloatStorage") + q(6) + data.tobytes())
    out = t7.loads(ten)
    assert out.dtype == np.float32 and out.shape == (2, 3)
    np.testing.assert_array_equal(out, data.reshape(2, 3))
    with pytest.raises(t7.T7Error):
        t7.loads(ten[:-4])                                    # truncated storage
    with pytest.raises(t7.T7Error):
        t7.loads(i32(99))                                     # unknown tag


def test_function_records_known_answer(t7):
    """Serialized Lua functions (a fixer checkpoint holds one: models.lua:404 sets drop.evaluate = function() end).
    Bytes assembled by hand from torch7 File.lua's readObject, not from the test writer: tag 6 (TYPE_FUNCTION, legacy)
    is int32 size + dumped chunk + upvalues object with NO object index; tag 8 (TYPE_RECUR_FUNCTION) and 7 (its legacy
    twin) carry an index first and are memoised like tables."""
    i32 = lambda v: struct.pack("<i", v)
    chunk = b"\x1bLJ\x02\x00dummy"
    upv = i32(3) + i32(5) + i32(0)                              # upvalues: empty table with object index 5
    f6 = t7.loads(i32(6) + i32(len(chunk)) + chunk + upv)
    assert f6.dumped == chunk and f6.upvalues == {}
    for tag in (8, 7):
        f = t7.loads(i32(tag) + i32(2) + i32(len(chunk)) + chunk + upv)
        assert f.dumped == chunk and f.upvalues == {}
    # {f = <tag-8 function, index 2>, g = <the same function by index>, h = <tag-6 function>, n = 4}
    s = lambda x: i32(2) + i32(len(x)) + x
    tbl = (i32(3) + i32(1) + i32(4)
           + s(b"f") + i32(8) + i32(2) + i32(len(chunk)) + chunk + i32(3) + i32(3) + i32(0)
           + s(b"g") + i32(8) + i32(2)
           + s(b"h") + i32(6) + i32(len(chunk)) + chunk + i32(0)
           + s(b"n") + i32(1) + struct.pack("<d", 4.0))
    out = t7.loads(tbl)
    assert out["f"] is out["g"] and out["f"].upvalues == {} and out["h"].upvalues is None and out["n"] == 4


@pytest.mark.parametrize("kw", [{}, {"legacy": True}, {"long_size": 4}, {"cuda": True}])
def test_round_trip_structures(t7, kw):
    rng = np.random.default_rng(0)
    w = rng.normal(size=(4, 5)).astype(np.float32)
    shared = {"w": w, "again": w, "n": 3, "s": "str", "flag": False, "none_val": None, "list": [1.5, "b", {"k": 2}]}
    shared["self"] = shared                                   # cyclic table
    long_size = kw.get("long_size", 8)
    out = t7.loads(tw.dumps(shared, **kw), long_size=long_size)
    np.testing.assert_array_equal(out["w"], w)
    assert out["again"] is out["w"]                           # shared by index, not copied twice
    assert out["self"] is out and out["n"] == 3 and out["s"] == "str" and out["flag"] is False
    assert out["list"] == {1: 1.5, 2: "b", 3: {"k": 2}}
    assert "none_val" not in out or out["none_val"] is None
    # strided view: the transposed 3x2 window at offset 1 of a 12-element storage
    st = np.arange(12, dtype=np.float64)
    v = t7.loads(tw.dumps(tw.TensorView(st, [3, 2], [1, 4], 1), **kw), long_size=long_size)
    np.testing.assert_array_equal(v, np.array([[1, 5], [2, 6], [3, 7]], np.float64))
    ints = t7.loads(tw.dumps(np.arange(5, dtype=np.int64), **kw), long_size=long_size)
    assert ints.dtype == np.int64 and ints.tolist() == [0, 1, 2, 3, 4]


def _seq(*mods):
    return tw.Module("nn.Sequential", modules=list(mods), train=False)


def _bn(prefix, p, spatial, std_form=False):
    f = dict(weight=p[prefix + ".g"], bias=p[prefix + ".b"], running_mean=p[prefix + ".m"], eps=1e-5, momentum=0.1, affine=True, train=False)
    if std_form:
        f["running_std"] = (1.0 / np.sqrt(p[prefix + ".v"].astype(np.float64) + 1e-5)).astype(np.float32)
    else:
        f["running_var"] = p[prefix + ".v"]
    return tw.Module("nn.SpatialBatchNormalization" if spatial else "nn.BatchNormalization", **f)


def _conv(prefix, p, cudnn=False, mm_view=False):
    w = p[prefix + ".w"]
    return tw.Module(("cudnn" if cudnn else "nn") + ".SpatialConvolution", weight=w.reshape(w.shape[0], -1) if mm_view else w, bias=p[prefix + ".b"],
                     kW=3, kH=3, dW=1, dH=1, padW=1, padH=1, nInputPlane=w.shape[1], nOutputPlane=w.shape[0],
                     gradWeight=np.zeros((0,), np.float32))


def test_g_and_r_checkpoints_to_blobs(pkg, t7):
    W = pkg.weights
    C, H, Wd, nd = 1, 16, 16, 24
    gb = W.init_G(C, H, Wd, nd, seed=3, stress=True)
    p = W.unpack(gb, W.g_layout(C, H, Wd, nd))
    G = _seq(tw.Module("nn.Linear", weight=p["lin.w"], bias=p["lin.b"]), _bn("bn0", p, False), tw.Module("cudnn.ReLU", inplace=True),
             tw.Module("nn.View", size=np.array([512, 4, 4], np.int64), numElements=512 * 16),
             tw.Module("nn.SpatialUpSamplingNearest", scale_factor=2), _conv("c1", p, cudnn=True), _bn("bn1", p, True), tw.Module("cudnn.ReLU"),
             tw.Module("nn.SpatialUpSamplingNearest", scale_factor=2), _conv("c2", p, cudnn=True, mm_view=True), _bn("bn2", p, True, std_form=False), tw.Module("cudnn.ReLU"),
             _conv("c3", p, cudnn=True), tw.Module("nn.Sigmoid"))
    opt = {"noiseDim": nd, "noiseMethod": "normal", "height": H, "width": Wd, "colorSpace": "y", "scale": 16}
    ck = t7.loads(tw.dumps({"G": G, "opt": opt, "epoch": 12}, cuda=True))
    c2, h2, w2, nd2, blob = t7.g_blob(ck)
    assert (c2, h2, w2, nd2) == (C, H, Wd, nd)
    np.testing.assert_array_equal(blob, gb)                   # bit-exact: what went in comes out in blob order

    rb = W.init_R(C, H, Wd, nd, seed=5, stress=True)
    q = W.unpack(rb, W.r_layout(C, H, Wd, nd))
    conv = _seq(tw.Module("nn.Dropout", p=0.5, v2=False),
                *[m for i in range(1, 7) for m in (_conv(f"c{i}", q), _bn(f"bn{i}", q, True, std_form=(i % 2 == 0)), tw.Module("nn.ELU", alpha=1.0))])
    R = _seq(conv, tw.Module("nn.View", size=np.array([128 * 16], np.int64)), tw.Module("nn.Linear", weight=q["l1.w"], bias=q["l1.b"]),
             _bn("bn7", q, False), tw.Module("nn.ELU"), tw.Module("nn.Dropout", p=0.5), tw.Module("nn.Linear", weight=q["l2.w"], bias=q["l2.b"]))
    ck_r = t7.loads(tw.dumps({"R": R, "opt": opt}))
    blob_r = t7.r_blob(ck_r, C, H, Wd, nd)
    lay = W.r_layout(C, H, Wd, nd)
    got, want = W.unpack(blob_r, lay), q
    for name, _ in lay:
        if name.endswith(".v") and name[:-2] in ("bn2", "bn4", "bn6"):   # running_std form: var recovered through 1/std^2 - eps
            np.testing.assert_allclose(got[name], want[name], rtol=2e-6, atol=1e-7)
        else:
            np.testing.assert_array_equal(got[name], want[name])

    # a checkpoint of the wrong architecture is refused with a message, not mis-packed
    bad = _seq(tw.Module("nn.Linear", weight=p["lin.w"], bias=p["lin.b"]), _conv("c1", p))
    with pytest.raises(t7.T7Error):
        t7.g_blob(t7.loads(tw.dumps({"G": bad, "opt": opt})))


def test_round_trip_property(t7):
    """Random nested tables / tensors survive writer -> reader unchanged (hypothesis)."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    leaves = st.one_of(
        st.booleans(), st.integers(min_value=-2 ** 40, max_value=2 ** 40), st.floats(allow_nan=False, allow_infinity=False, width=64),
        st.text(alphabet=st.characters(min_codepoint=32, max_codepoint=126), max_size=12),
        st.lists(st.floats(allow_nan=False, width=32), min_size=0, max_size=6).map(lambda v: np.asarray(v, np.float32)),
        st.lists(st.integers(-100, 100), min_size=1, max_size=6).map(lambda v: np.asarray(v, np.int64).reshape(1, -1)),
    )
    keys = st.one_of(st.integers(1, 50), st.text(alphabet="abcxyz_", min_size=1, max_size=6))
    tables = st.recursive(leaves, lambda ch: st.dictionaries(keys, ch, max_size=4), max_leaves=12)

    def same(a, b):
        if isinstance(a, np.ndarray):
            return isinstance(b, np.ndarray) and a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
        if isinstance(a, dict):
            return isinstance(b, dict) and a.keys() == b.keys() and all(same(a[k], b[k]) for k in a)
        if isinstance(a, float):
            return float(b) == a
        return a == b and type(a) is type(b) or (isinstance(a, int) and not isinstance(a, bool) and b == a)

    @settings(max_examples=150, deadline=None)
    @given(tables, st.sampled_from([8, 4]), st.booleans())
    def run(obj, long_size, legacy):
        out = t7.loads(tw.dumps(obj, long_size=long_size, legacy=legacy), long_size=long_size)
        if isinstance(obj, float) and obj == int(obj) and abs(obj) < 2 ** 53:
            assert out == obj                                   # integral numbers come back as ints (Lua has one number type)
        else:
            assert same(obj, out), (obj, out)

    run()
