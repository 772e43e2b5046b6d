"""Profiling driver for the database kernels: python tools/prof_search.py <case> (run under ncu, see profiles/r02/README).
cases: tc (4096-needle search over 1M x 100 N(0,1) rows, tensor-core path), exact (same, fmaf-chain kernels), tc256 (configs[4]: 1.25M x 256,
       top-100), kmeans20 / assign20 (k = 20 over 4M x 100), labeltc20 (the same through label_tc.cuh), kmeans1024 (k = 1024 over 1.25M x 256),
       search4 (4 needles over 4M x 100), l2 (200k pairs), nearest (8 queries)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from __graft_entry__ import load_package

pkg = load_package()
case = sys.argv[1] if len(sys.argv) > 1 else "tc"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = pkg.Context(0)
rng = np.random.default_rng(0)
if case in ("tc", "exact"):
    N, d, Q, k = 1_000_000, 100, 4096, 20
    ctx.db_synthetic(N, d, seed=8)
    ctx.set_option("search_tc", 1 if case == "tc" else 0)
    rows = np.arange(1, Q + 1, dtype=np.int64) * 244
    for _ in range(reps):
        ids, sc = ctx.search_rows(rows, k)
    print(case, "self first:", bool((ids[:, 0] == rows).all()), "tc counters", ctx.tc_counters())
elif case in ("kmeans20", "assign20", "search4"):
    N, d, k = 4_000_000, 100, 20
    ctx.db_synthetic(N, d, seed=11)
    init = rng.standard_normal(size=(k, d), dtype=np.float32)
    init /= np.linalg.norm(init, axis=1, keepdims=True)
    for _ in range(reps):
        if case == "kmeans20":
            ctx.kmeans(k, 2, init, want_labels=False)
        elif case == "assign20":
            ctx.assign_cosine_min(init)
        else:
            ctx.search_rows(np.array([99, 199, 299, 399], np.int64), 20)
elif case == "gr":
    C, H, W, ND, N = 1, 32, 32, 100, 8192                     # one default chunk of G and of R
    ctx.load_G(C, H, W, ND, pkg.weights.init_G(C, H, W, ND))
    ctx.load_R(0, C, H, W, ND, pkg.weights.init_R(C, H, W, ND))
    ctx.buffer_put(pkg._lib.BUF_NOISE, rng.standard_normal(size=(N, ND), dtype=np.float32))
    for _ in range(reps + 1):
        ctx.forward_G(None, N=N, want_images=False)
        ctx.forward_R(0, None, N=N, want_attrs=False)
elif case == "kmeans1024":
    N, d, k = 1_250_000, 256, 1024
    ctx.db_synthetic(N, d, seed=8)
    init = rng.standard_normal(size=(k, d), dtype=np.float32)
    init /= np.linalg.norm(init, axis=1, keepdims=True)
    for _ in range(reps):
        ctx.kmeans(k, 1, init, want_labels=False)
elif case == "tc256":
    N, d, Q, k = 1_250_000, 256, 4096, 100
    ctx.db_synthetic(N, d, seed=8)
    rows = np.arange(1, Q + 1, dtype=np.int64) * 300
    for _ in range(reps):
        ids, sc = ctx.search_rows(rows, k)
    print(case, "self first:", bool((ids[:, 0] == rows).all()), "tc counters", ctx.tc_counters())
elif case == "labeltc20":
    N, d, k = 4_000_000, 100, 20
    ctx.db_synthetic(N, d, seed=11)
    ctx.set_option("label_tc", 1)
    init = rng.standard_normal(size=(k, d), dtype=np.float32)
    init /= np.linalg.norm(init, axis=1, keepdims=True)
    for _ in range(reps):
        ctx.kmeans(k, 2, init, want_labels=False)
        ctx.assign_cosine_min(init)
elif case in ("l2", "nearest"):
    n = 200_000
    a = rng.random((n, 1024), dtype=np.float32)
    b = a[::-1].copy()
    for _ in range(reps):
        if case == "l2":
            ctx.l2(a, b)
        else:
            ctx.nearest_l2(b[:8], a)
ctx.profile_enable(False)
ctx.close()
