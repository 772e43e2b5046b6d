"""GPU parity of G3 / R_default (models.lua:104-143, 389-464) through the C ABI against the
CPU oracle.  Tolerances are north_star's: generated pixels within 2e-2 max-abs, recovered
vectors cosine >= 0.999 -- plus a relative-L2 bound so a constant offset cannot hide errors."""
import os

import numpy as np
import pytest

from util import rel_l2, row_cosine

pytestmark = pytest.mark.gpu

PIX_TOL = 2e-2      # north_star: "Generated pixels must be within 2e-2 max-abs"
COS_TOL = 0.999     # north_star: "Recovered vectors must reach cosine >= 0.999"
REL_TOL = 3e-2      # our own, stricter companion bound (bf16 operands, fp32 accumulate)

GEOMS = [
    # C, H, W, nd, N
    (1, 32, 32, 32, 37),
    (1, 32, 32, 100, 19),
    (3, 64, 64, 256, 5),
    (1, 16, 16, 32, 21),
    (3, 32, 32, 64, 9),
]


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


def _setup(pkg, ctx, geom, stress, impl):
    C, H, W, nd, N = geom
    # impl: 1 = CUDA-core A/B kernel, 0 = tcgen05 with the default CTA-pair (cta_group::2) layers, 2 = tcgen05 single-CTA
    ctx.set_option("conv_impl", 1 if impl == 1 else 0)
    ctx.set_option("cta_pairs", 0 if impl == 2 else int(os.environ.get("GANREV_CTA_PAIRS", "31")))
    ctx.set_option("chunk", 16)        # several chunks, the last one ragged
    gb = pkg.weights.init_G(C, H, W, nd, seed=1, stress=stress)
    rb = pkg.weights.init_R(C, H, W, nd, seed=2, stress=stress)
    rfb = pkg.weights.init_R(C, H, W, nd, seed=4, stress=stress)
    ctx.load_G(C, H, W, nd, gb)
    ctx.load_R(0, C, H, W, nd, rb)
    ctx.load_R(1, C, H, W, nd, rfb)
    noise = np.random.default_rng(3).normal(size=(N, nd)).astype(np.float32)
    return gb, rb, rfb, noise


@pytest.mark.parametrize("impl", [1, 0, 2], ids=["cudacore", "tcgen05", "tcgen05_1cta"])
@pytest.mark.parametrize("geom", GEOMS, ids=lambda g: "C%dx%dx%d_nd%d_N%d" % g)
def test_G_matches_oracle(pkg, orc, ctx, geom, impl):
    C, H, W, nd, N = geom
    gb, _, _, noise = _setup(pkg, ctx, geom, True, impl)
    want = orc.forward_G(gb, C, H, W, nd, noise)
    got = ctx.forward_G(noise)
    assert got.shape == want.shape
    err = np.abs(got - want).max()
    assert want.std() > 0.05, "stress weights should spread pixels over (0,1)"
    assert err <= PIX_TOL, f"max |pixel diff| {err}"


@pytest.mark.parametrize("impl", [1, 0, 2], ids=["cudacore", "tcgen05", "tcgen05_1cta"])
@pytest.mark.parametrize("geom", GEOMS, ids=lambda g: "C%dx%dx%d_nd%d_N%d" % g)
def test_R_matches_oracle(pkg, orc, ctx, geom, impl):
    C, H, W, nd, N = geom
    gb, rb, rfb, noise = _setup(pkg, ctx, geom, True, impl)
    images = orc.forward_G(gb, C, H, W, nd, noise)          # identical inputs for both sides
    want = orc.forward_R(rb, C, H, W, nd, images)
    got = ctx.forward_R(0, images)
    assert row_cosine(got, want).min() >= COS_TOL
    assert rel_l2(got, want) <= REL_TOL, rel_l2(got, want)
    # fixer: explicit 50% input dropout mask, multiply-no-rescale (models.lua:399-406)
    mask = (np.random.default_rng(5).random(images.shape) >= 0.5).astype(np.uint8)
    want_f = orc.forward_R(rfb, C, H, W, nd, images, mask)
    got_f = ctx.forward_R(1, images, mask)
    assert row_cosine(got_f, want_f).min() >= COS_TOL
    assert rel_l2(got_f, want_f) <= REL_TOL, rel_l2(got_f, want_f)
    assert rel_l2(got_f, want) > 10 * REL_TOL, "mask must change the result"


@pytest.mark.parametrize("pairs", [31, 0], ids=["cta_pairs", "1cta"])
@pytest.mark.parametrize("geom", [(1, 32, 32, 100, 150), (1, 16, 16, 32, 40), (1, 64, 64, 32, 11), (3, 32, 32, 64, 70), (3, 64, 64, 256, 9)], ids=lambda g: "C%dx%dx%d_nd%d_N%d" % g)
def test_G_fused_last_conv_matches_two_pass(pkg, orc, ctx, geom, pairs):
    """The last conv's tap products come out of G conv2's epilogue (fuse_conv3 = 1, the default: fp32 FFMA2 on the
    un-rounded activation, conv_tc.cuh FUSE3) or out of a separate 1x1 tensor-core pass over the bf16 activation
    (fuse_conv3 = 0).  Both against the oracle (models.lua:127-133), and against each other at bf16-rounding level;
    several items per persistent CTA and a ragged last chunk."""
    C, H, W, nd, N = geom
    gb = pkg.weights.init_G(C, H, W, nd, seed=1, stress=True)
    noise = np.random.default_rng(3).normal(size=(N, nd)).astype(np.float32)
    want = orc.forward_G(gb, C, H, W, nd, noise)
    ctx.set_option("conv_impl", 0)
    ctx.set_option("cta_pairs", pairs)
    ctx.set_option("chunk", 64)
    got = {}
    try:
        for fuse in (0, 1):
            ctx.set_option("fuse_conv3", 2 * fuse)      # 2: also for C = 3 (off by default there: slower, not wrong)
            ctx.load_G(C, H, W, nd, gb)
            ctx.profile_reset(); ctx.profile_enable(True)
            got[fuse] = ctx.forward_G(noise)
            ctx.profile_enable(False)
            assert (ctx.profile().get("g_conv3_taps", {"launches": 0})["launches"] > 0) == (fuse == 0), "fuse_conv3 must decide whether the separate pass runs"
            assert np.abs(got[fuse] - want).max() <= PIX_TOL
    finally:
        ctx.set_option("fuse_conv3", 1)
        ctx.set_option("chunk", 16)
        ctx.set_option("cta_pairs", 31)
    assert np.abs(got[1] - want).max() <= np.abs(got[0] - want).max() * 1.5 + 1e-4   # the fused form skips one bf16 rounding
    assert np.abs(got[0] - got[1]).max() <= 1e-2


@pytest.mark.parametrize("opts", [{"xpose2": 2}, {"xpose2": 0}, {"tma_hybrid": 1}, {"tma_store": 0}, {"tma_store": 2}], ids=lambda o: "_".join(f"{k}{v}" for k, v in o.items()))
def test_epilogue_store_variants_match_oracle(pkg, orc, ctx, opts):
    """The A/B knobs of the conv epilogue's store path (ganrev.h: xpose2, tma_hybrid, tma_store) select different code in the shipped
    library: each must give the same G / R results within tolerance (several items per CTA, ragged last chunk)."""
    C, H, W, nd, N = 1, 32, 32, 32, 150
    defaults = {"xpose2": 1, "tma_hybrid": 0, "tma_store": 1}
    gb = pkg.weights.init_G(C, H, W, nd, seed=1, stress=True)
    rb = pkg.weights.init_R(C, H, W, nd, seed=2, stress=True)
    noise = np.random.default_rng(3).normal(size=(N, nd)).astype(np.float32)
    want_img = orc.forward_G(gb, C, H, W, nd, noise)
    want_att = orc.forward_R(rb, C, H, W, nd, want_img)
    ctx.set_option("conv_impl", 0)
    ctx.set_option("cta_pairs", 31)
    ctx.set_option("chunk", 64)
    try:
        for k, v in opts.items():
            ctx.set_option(k, v)
        ctx.load_G(C, H, W, nd, gb)
        ctx.load_R(0, C, H, W, nd, rb)
        assert np.abs(ctx.forward_G(noise) - want_img).max() <= PIX_TOL
        got = ctx.forward_R(0, want_img)
        assert row_cosine(got, want_att).min() >= COS_TOL
        assert rel_l2(got, want_att) <= REL_TOL
    finally:
        for k, v in defaults.items():
            ctx.set_option(k, v)
        ctx.set_option("chunk", 16)


BENCH_SHAPES = [
    # C, H, W, nd, N, rows checked against the oracle: BASELINE configs[3] and configs[4] geometries at the DEFAULT chunk
    (1, 32, 32, 100, 20000, 1000),
    (3, 64, 64, 256, 5000, 256),
]


@pytest.mark.parametrize("shape", BENCH_SHAPES, ids=lambda g: "C%dx%dx%d_nd%d_N%d_check%d" % g)
def test_benchmark_launch_shape_matches_oracle(pkg, orc, ctx, shape):
    """The launch shape bench.py times: default chunk (8192 32x32 faces' worth of pixels, ~55 items per persistent CTA per
    layer, accumulator / smem-ring / transpose-buffer wrap-around at depth), several chunks with a ragged last one.
    Sampled rows of G and of R against the oracle within north_star's tolerances, and the device-resident chain
    (NULL pointers, what bench.py's `value` times) equal to the host-pointer chain bit for bit."""
    C, H, W, nd, N, n_check = shape
    ctx.set_option("conv_impl", 0)
    ctx.set_option("cta_pairs", 31)
    ctx.set_option("chunk", 0)                                  # library default
    gb = pkg.weights.init_G(C, H, W, nd, seed=1, stress=True)
    rb = pkg.weights.init_R(C, H, W, nd, seed=2, stress=True)
    ctx.load_G(C, H, W, nd, gb)
    ctx.load_R(0, C, H, W, nd, rb)
    rng = np.random.default_rng(17)
    noise = rng.standard_normal(size=(N, nd), dtype=np.float32)
    img = ctx.forward_G(noise)
    att = ctx.forward_R(0, img)
    # rows spread over every chunk, always including the first / last rows of the batch and of a chunk boundary
    chunk = max(256, 8192 * 1024 // (H * W))
    rows = np.unique(np.concatenate([rng.choice(N, size=n_check - 8, replace=False),
                                     [0, 1, chunk - 1, chunk, min(N - 1, 2 * chunk), N - 2, N - 1, N // 2]]))
    want_img = orc.forward_G(gb, C, H, W, nd, noise[rows])
    err = np.abs(img[rows] - want_img).max()
    assert err <= PIX_TOL, f"max |pixel diff| {err}"
    want_att = orc.forward_R(rb, C, H, W, nd, img[rows])           # identical inputs for both sides
    assert row_cosine(att[rows], want_att).min() >= COS_TOL
    assert rel_l2(att[rows], want_att) <= REL_TOL, rel_l2(att[rows], want_att)
    # resident chain == host-pointer chain, bit for bit, over ALL rows
    ctx.buffer_put(pkg._lib.BUF_NOISE, noise)
    ctx.forward_G(None, N=N, want_images=False)
    att2 = ctx.forward_R(0, None, N=N)
    np.testing.assert_array_equal(att2, att)
    for r0 in (0, chunk - 4, N - 8):
        np.testing.assert_array_equal(ctx.buffer_get(pkg._lib.BUF_IMAGES, r0, 8), img[r0:r0 + 8])
    # determinism at depth: a second pass over the same input is identical
    np.testing.assert_array_equal(ctx.forward_R(0, None, N=N), att)
    ctx.set_option("chunk", 16)


def test_one_geometry_per_context(pkg, ctx):
    """include/ganrev.h: loading a model of a different geometry unloads the old models and empties the resident buffers
    (R.nd > G.nd wrote past ATTRS; a reload kept stale row counts)."""
    W_ = pkg.weights
    ctx.load_G(1, 32, 32, 32, W_.init_G(1, 32, 32, 32))
    ctx.load_R(0, 1, 32, 32, 32, W_.init_R(1, 32, 32, 32))
    noise = np.zeros((4, 32), np.float32)
    ctx.forward_G(noise)
    ctx.forward_R(0, None, N=4)
    ctx.load_R(1, 1, 32, 32, 100, W_.init_R(1, 32, 32, 100))      # another noise_dim: G and R slot 0 are gone
    with pytest.raises(pkg.GanrevError):
        ctx.forward_R(0, None, N=4)
    with pytest.raises(pkg.GanrevError):
        ctx.forward_G(None, N=4, want_images=False)
    with pytest.raises(pkg.GanrevError):
        ctx.forward_R(1, None, N=4)                               # the resident images were emptied too
    with pytest.raises(pkg.GanrevError):
        ctx.buffer_get(pkg._lib.BUF_IMAGES, 0, 1)
    ctx.load_G(1, 32, 32, 100, W_.init_G(1, 32, 32, 100))
    img = ctx.forward_G(np.zeros((4, 100), np.float32))
    assert ctx.forward_R(1, img).shape == (4, 100)


def test_tanh_output(pkg, orc, ctx):
    C, H, W, nd, N = 1, 32, 32, 32, 8
    ctx.set_option("conv_impl", 0)
    rb = pkg.weights.init_R(C, H, W, nd, seed=2, stress=True)
    ctx.load_G(C, H, W, nd, pkg.weights.init_G(C, H, W, nd, seed=1, stress=True))
    ctx.load_R(0, C, H, W, nd, rb, tanh_out=True)            # noiseMethod ~= "normal", models.lua:452-454
    images = np.random.default_rng(0).random((N, C, H, W)).astype(np.float32)
    want = orc.forward_R(rb, C, H, W, nd, images, None, tanh_out=True)
    got = ctx.forward_R(0, images)
    assert np.abs(got).max() <= 1.0
    assert np.abs(got - want).max() <= 6e-2   # tanh of O(1..3) pre-activations: bf16-level error


def test_default_init_and_tc_vs_cudacore(pkg, orc, ctx):
    """The reference's own init (weight-init.lua 'heuristic'): tiny activations, images ~0.5."""
    geom = (1, 32, 32, 32, 24)
    C, H, W, nd, N = geom
    gb, rb, _, noise = _setup(pkg, ctx, geom, False, 0)
    want_img = orc.forward_G(gb, C, H, W, nd, noise)
    got_img = ctx.forward_G(noise)
    assert np.abs(got_img - want_img).max() <= 1e-3
    # relative accuracy of the (tiny) image-dependent signal around 0.5
    assert rel_l2(got_img - 0.5, want_img - 0.5) <= REL_TOL
    got_att = ctx.forward_R(0, want_img)
    want_att = orc.forward_R(rb, C, H, W, nd, want_img)
    assert row_cosine(got_att, want_att).min() >= COS_TOL
    assert rel_l2(got_att, want_att) <= REL_TOL
    ctx.set_option("conv_impl", 1)
    cc_img = ctx.forward_G(noise)
    cc_att = ctx.forward_R(0, want_img)
    # same bf16 operands, different accumulation order only
    assert rel_l2(got_img - 0.5, cc_img - 0.5) <= 1e-2
    assert rel_l2(got_att, cc_att) <= 1e-2
    ctx.set_option("conv_impl", 0)


def test_resident_chain_and_fix_l2(pkg, orc, ctx):
    """G -> R -> G with device-resident intermediates (NULL pointers) equals the host-pointer
    path bit for bit, and fix_l2 matches the oracle's G(R_fixer(x)) distance within tolerance."""
    geom = (1, 32, 32, 32, 50)
    C, H, W, nd, N = geom
    gb, rb, rfb, noise = _setup(pkg, ctx, geom, True, 0)
    img = ctx.forward_G(noise)
    att = ctx.forward_R(0, img)
    ctx.buffer_put(pkg._lib.BUF_NOISE, noise)
    assert ctx.forward_G(None, N=N, want_images=False) is None
    att2 = ctx.forward_R(0, None, N=N)
    np.testing.assert_array_equal(att, att2)
    np.testing.assert_array_equal(ctx.buffer_get(pkg._lib.BUF_IMAGES, 0, N), img)
    mask = (np.random.default_rng(5).random(img.shape) >= 0.5).astype(np.uint8)
    a1, fixed, l2 = ctx.fix_l2(1, img, mask)
    np.testing.assert_array_equal(a1, ctx.forward_R(1, img, mask))
    np.testing.assert_array_equal(fixed, ctx.forward_G(a1))
    # exact distance of the library's own pair of image sets
    from util import assert_bitexact
    assert_bitexact(l2, orc.l2(img, fixed), "fix_l2 distances")
    # tolerance check against the all-oracle chain
    want_fixed = orc.forward_G(gb, C, H, W, nd, orc.forward_R(rfb, C, H, W, nd, img, mask))
    assert np.abs(fixed - want_fixed).max() <= 3 * PIX_TOL      # two chained bf16 nets
    assert np.abs(fixed - want_fixed).mean() <= 0.25 * PIX_TOL


def test_errors(pkg, ctx):
    with pytest.raises(pkg.GanrevError):
        ctx.load_G(2, 32, 32, 32, np.zeros(10, np.float32))          # C must be 1 or 3
    with pytest.raises(pkg.GanrevError):
        ctx.load_G(1, 48, 48, 32, np.zeros(10, np.float32))          # power-of-two geometry only
    with pytest.raises(pkg.GanrevError):
        ctx.load_G(1, 32, 32, 32, np.zeros(10, np.float32))          # wrong blob size
    c2 = pkg.Context(0)
    with pytest.raises(pkg.GanrevError):
        c2.geom = (1, 32, 32, 32)
        c2.forward_G(np.zeros((2, 32), np.float32))                  # nothing loaded
    c2.close()


def test_t7_checkpoint_loads_like_blob(pkg, tmp_path):
    """a Torch7 `.net` file written in the documented format drives the library exactly like
    the blob it was made from (same bytes in -> same images / vectors out)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    import t7_writer as tw
    from test_t7_cpu import _seq, _bn, _conv
    W = pkg.weights
    C, H, Wd, nd = 1, 32, 32, 100
    gb = W.init_G(C, H, Wd, nd, seed=3, stress=True)
    p = W.unpack(gb, W.g_layout(C, H, Wd, nd))
    G = _seq(tw.Module("nn.Linear", weight=p["lin.w"], bias=p["lin.b"]), _bn("bn0", p, False), _conv("c1", p, cudnn=True), _bn("bn1", p, True),
             _conv("c2", p, cudnn=True), _bn("bn2", p, True), _conv("c3", p, cudnn=True))
    opt = {"noiseDim": nd, "noiseMethod": "normal", "height": H, "width": Wd, "colorSpace": "y"}
    path = tmp_path / "adversarial.net"
    path.write_bytes(tw.dumps({"G": G, "opt": opt}, cuda=True))
    ctx = pkg.Context(0)
    try:
        noise = np.random.default_rng(0).normal(size=(64, nd)).astype(np.float32)
        g1, opt_out = pkg.models.load_G(str(path), ctx=ctx)
        assert opt_out["noiseDim"] == nd
        a = g1.forward(noise)
        b = pkg.models.create_G((C, H, Wd), nd, blob=gb, ctx=ctx).forward(noise)
        np.testing.assert_array_equal(a, b)
    finally:
        ctx.close()


def test_apply_r_main_end_to_end(pkg, tmp_path):
    """apply_r.lua main() through the library (random-init nets, 1,200 faces so every mode's constants fit):
    every artefact of the reference is written and the returned data is consistent."""
    ctx = pkg.Context(0)
    try:
        out = pkg.apply_r.main(writeTo=str(tmp_path), nbImages=1200, ctx=ctx)
        # component variations (apply_r.lua:111-136): row i*16 + j is the base vector with component i set to steps[j] ...
        vn, steps = out["variation_noise"], np.linspace(-3, 3, 16).astype(np.float32)
        assert vn.shape == (100 * 16, 100)
        for i in (0, 37, 99):
            blk = vn[i * 16:(i + 1) * 16]
            np.testing.assert_allclose(blk[:, i], steps, rtol=1e-6)
            others = np.delete(blk, i, axis=1)
            assert (others == others[0]).all()
        base = vn[16].copy(); base[1] = vn[0][1]                  # the base vector: any row with its varied component restored
        assert np.array_equal(np.delete(vn[0], 0), np.delete(base, 0))
        # ... and the images are G of exactly those rows (each image is computed independently of its batch position)
        rows = np.array([0, 15, 16, 600, 1599])
        np.testing.assert_array_equal(ctx.forward_G(vn[rows]), out["variations"][rows])
        assert np.abs(out["variations"][0] - out["variations"][15]).max() > 0   # varying a component changes the face
    finally:
        ctx.close()
    names = sorted(os.path.basename(f) for f in out["files"])
    assert "variations.jpg" in names and "anomalies.jpg" in names and "fixed_pairs.jpg" in names
    assert "fixed_images_528.jpg" in names and "fixed_images_528_unfixed.jpg" in names
    assert sum(n.startswith("similar_attributes_") for n in names) == 5 and sum(n.startswith("similar_pixelwise_") for n in names) == 5
    nonempty = int((out["clusters"]["member_counts"] > 0).sum())
    assert sum(n.startswith("cluster_") for n in names) == nonempty > 0
    assert all(os.path.getsize(f) > 0 for f in out["files"])
    ids, sc = out["similar"]["attributes"]
    assert ids.shape == (5, 100) and np.array_equal(ids[:, 0], np.array([99, 199, 299, 399, 499]))   # a needle's best match is itself
    assert out["anomalies"]["flags"].shape == (528,) and 0 < out["anomalies"]["flags"].sum() < 528
    assert out["variations"].shape == (100 * 16, 1, 32, 32)
