"""A/B of ganrev_nearest_l2 (sample.lua:128-148): tensor-core filter + canonical candidates (nearest_tc.cuh) against the kernel that
evaluates every (row, query) pair canonically.  200k resident 32x32 faces, 8 / 32 query faces.  python tools/exp_nearest.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
ctx = pkg.Context(0)
rng = np.random.default_rng(11)
N, px = 200_000, 1024
a = rng.random((N, px), dtype=np.float32)
ctx.load_G(1, 32, 32, 100, pkg.weights.init_G(1, 32, 32, 100))
ctx.buffer_put(pkg._lib.BUF_IMAGES, a.reshape(N, 1, 32, 32))
for Q in (8, 32):
    q = (a[rng.integers(0, N, size=Q)] + rng.normal(scale=0.05, size=(Q, px))).astype(np.float32)
    out = {}
    for stc in (1, 0):
        ctx.set_option("stream_tc", stc)
        ctx.nearest_l2(q, None, N=N)
        ctx.tfs_stats()
        ctx.profile_reset(); ctx.profile_enable(True)
        for _ in range(3):
            out[stc] = ctx.nearest_l2(q, None, N=N)
        ctx.profile_enable(False)
        e = ctx.profile()["nearest_l2"]
        ms = e["ms"] / 3
        passes = (Q + 7) // 8
        print(f"Q={Q} stream_tc={stc}: {ms:7.3f} ms per call ({passes} passes over the set), {4.0 * N * px * passes / ms * 1e-6:7.0f} GB/s   {ctx.tfs_stats() if stc else ''}", flush=True)
    assert (out[0][0] == out[1][0]).all() and (out[0][1].view(np.uint64) == out[1][1].view(np.uint64)).all()
ctx.set_option("stream_tc", 1)
ctx.close()
