#!/usr/bin/env python
"""Generates tests/golden/*.npz -- small known-answer vectors from statements INDEPENDENT of
the C oracle (PyTorch-CPU for the networks; plain numpy / Python integers for the exact-match
arithmetic).  They pin the oracle, not Torch7: the reference ships no golden vectors and cannot
run here ("parity unpinned").

Run from the repo root:  python tests/golden/make_golden.py
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

f32 = np.float32


def fma32(a, b, c):
    """fp32 fmaf via float64: the 48-bit product is exact in float64; the sum is rounded once to
    float64 and once to float32 (double rounding is possible in principle -- the generated
    vectors are checked against the oracle at generation time)."""
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def dot_seq(a, b):
    acc = f32(0.0)
    for x, y in zip(a, b):
        acc = fma32(x, y, acc)
    return acc


def rnorm(a):
    return f32(1.0) / f32(dot_seq(a, a) + f32(1e-12))


def cosine(a, b):
    return f32(dot_seq(a, b) * f32(np.sqrt(f32(rnorm(a) * rnorm(b)))))


def search(db, qs, k):
    ids, scs = [], []
    for q in qs:
        sc = [cosine(q, r) for r in db]
        order = sorted(range(len(db)), key=lambda j: (1 if math.isnan(sc[j]) else 0, -float(sc[j]) if not math.isnan(sc[j]) else 0.0, j))
        ids.append(order[:k]); scs.append([sc[j] for j in order[:k]])
    return np.array(ids, np.int64), np.array(scs, f32)


def kmeans(x, k, niter, init):
    N, d = x.shape
    mx = float(np.abs(x).max())
    e = math.frexp(mx)[1] if mx > 0 else 0
    n = 0
    while (1 << n) < N:
        n += 1
    shift = max(0, min(60, 62 - n - e))
    cen = init.copy()
    tot = [0] * k
    lab = [-1] * N
    for _ in range(niter):
        c2 = [f32(0.5) * dot_seq(c, c) for c in cen]
        for i in range(N):
            best, bv = 0, None
            for j in range(k):
                v = f32(dot_seq(cen[j], x[i]) - c2[j])
                if bv is None or not (v <= bv):
                    best, bv = j, v
            lab[i] = best
        acc = [[0] * d for _ in range(k)]
        cnt = [0] * k
        for i in range(N):
            for c in range(d):
                acc[lab[i]][c] += int(np.rint(np.float64(x[i, c]) * 2.0 ** shift))
            cnt[lab[i]] += 1
        for j in range(k):
            if cnt[j]:
                for c in range(d):
                    cen[j, c] = f32(np.float64(acc[j][c]) / (np.float64(cnt[j]) * 2.0 ** shift))
            tot[j] += cnt[j]
    return cen, np.array(tot, f32), np.array(lab, np.int32), shift


def assign_min(x, cen):
    cl, cv = [], []
    for xi in x:
        best, bv = 0, None
        for j, c in enumerate(cen):
            v = cosine(xi, c)
            if bv is None or v < bv:
                best, bv = j, v
        cl.append(best); cv.append(bv)
    return np.array(cl, np.int32), np.array(cv, f32)


def l2_canonical(a, b):
    out = []
    for x, y in zip(a, b):
        lanes = [np.float64(0.0)] * 32
        for i in range(len(x)):
            d = f32(x[i] - y[i])
            lanes[(i >> 2) & 31] = lanes[(i >> 2) & 31] + np.float64(f32(d * d))
        off = 16
        while off >= 1:
            lanes = [lanes[l] + lanes[l ^ off] for l in range(32)]
            off >>= 1
        out.append(np.sqrt(lanes[0]))
    return np.array(out, np.float64)


def main():
    from __graft_entry__ import load_package
    from oracle import oracle as orc
    import torch_ref
    pkg = load_package()
    Wt = pkg.weights
    out = {}

    # 1. networks: PyTorch-CPU on stress weights from seeds (blobs are regenerated, not stored)
    for tag, (C, H, W, nd, N) in {"a": (1, 16, 16, 8, 3), "b": (3, 16, 16, 5, 2)}.items():
        gb = Wt.init_G(C, H, W, nd, seed=101, stress=True)
        rb = Wt.init_R(C, H, W, nd, seed=102, stress=True)
        noise = np.random.default_rng(103).normal(size=(N, nd)).astype(f32)
        img = torch_ref.forward_G(Wt.unpack(gb, Wt.g_layout(C, H, W, nd)), C, H, W, nd, noise)
        mask = (np.random.default_rng(104).random(img.shape) >= 0.5).astype(np.uint8)
        att = torch_ref.forward_R(Wt.unpack(rb, Wt.r_layout(C, H, W, nd)), C, H, W, nd, img)
        attm = torch_ref.forward_R(Wt.unpack(rb, Wt.r_layout(C, H, W, nd)), C, H, W, nd, img, mask)
        out.update({f"net_{tag}_geom": np.array([C, H, W, nd, N]), f"net_{tag}_noise": noise, f"net_{tag}_img": img,
                    f"net_{tag}_mask": mask, f"net_{tag}_att": att, f"net_{tag}_attm": attm,
                    f"net_{tag}_gsum": np.float64(gb.astype(np.float64).sum()), f"net_{tag}_rsum": np.float64(rb.astype(np.float64).sum())})
        assert np.abs(orc.forward_G(gb, C, H, W, nd, noise) - img).max() < 1e-4
        assert np.abs(orc.forward_R(rb, C, H, W, nd, img, mask) - attm).max() < 1e-3

    # 2. exact-match arithmetic from plain numpy / Python integers
    rng = np.random.default_rng(201)
    db = rng.normal(size=(60, 8)).astype(f32)
    db[7] = db[3]; db[20] = db[3]; db[11] = 0.0; db[40, 2] = np.nan
    qs = np.concatenate([db[[3, 11]], rng.normal(size=(2, 8)).astype(f32)])
    ids, sc = search(db, qs, 12)
    o_ids, o_sc = orc.search_cosine(db, qs, 12)
    assert (ids == o_ids).all() and (sc.view(np.uint32)[~np.isnan(sc)] == o_sc.view(np.uint32)[~np.isnan(o_sc)]).all()
    out.update(search_db=db, search_q=qs, search_ids=ids, search_sc=sc)

    x = rng.normal(size=(90, 6)).astype(f32)
    init = rng.normal(size=(4, 6)).astype(f32)
    init /= np.linalg.norm(init, axis=1, keepdims=True).astype(f32)
    cen, tot, lab, shift = kmeans(x, 4, 3, init)
    o_cen, o_tot, o_lab = orc.kmeans(x, 4, 3, init)
    assert shift == orc.kmeans_shift(x) and (lab == o_lab).all() and (cen.view(np.uint32) == o_cen.view(np.uint32)).all() and (tot == o_tot).all()
    cl, cv = assign_min(x, cen)
    o_cl, o_cv = orc.assign_cosine_min(x, cen)
    assert (cl == o_cl).all() and (cv.view(np.uint32) == o_cv.view(np.uint32)).all()
    out.update(km_x=x, km_init=init, km_cen=cen, km_tot=tot, km_lab=lab, km_shift=np.int32(shift), km_cl=cl, km_cv=cv)

    a = rng.random((5, 200)).astype(f32)
    b = (a + rng.normal(scale=0.1, size=a.shape)).astype(f32)
    d = l2_canonical(a, b)
    assert (d.view(np.uint64) == orc.l2(a, b).view(np.uint64)).all()
    sims = 1.0 - d
    thr = np.sort(sims)[math.floor(5 * 0.5) - 1]
    out.update(l2_a=a, l2_b=b, l2_d=d, flag_q=np.float64(0.5), flag_thr=np.float64(thr), flag_flags=(sims <= thr).astype(np.uint8))

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    print("wrote", os.path.join(HERE, "golden.npz"), os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes")


if __name__ == "__main__":
    main()
