"""A/B of the small-Q database kernels: TMA -> tf32 filter pipeline (stream_tc.cuh) against the fmaf-chain kernels, one pass
over 4M x 100 fp32 rows (the bench's hbm_kernels shapes) plus d = 32 / 128 / 256.  python tools/exp_stream_tc.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
ctx = pkg.Context(0)
shapes = [(4_000_000, 100, 20)] if len(sys.argv) < 2 else [(4_000_000, 100, 20), (8_000_000, 32, 20), (3_000_000, 128, 20), (1_500_000, 256, 20), (4_000_000, 100, 8), (4_000_000, 100, 32)]
DBG = int(os.environ.get('TFS_DBG', '0'))
for N, d, k in shapes:
    ctx.db_synthetic(N, d, seed=11)
    init = np.random.default_rng(0).standard_normal(size=(k, d)).astype(np.float32)
    init /= np.linalg.norm(init, axis=1, keepdims=True)
    rows = np.array([99, 199, 299, 399], np.int64)
    for stc in ((1,) if DBG or os.environ.get('TFS_ONLY') else (1, 0)):
        ctx.set_option("stream_tc", stc)
        ctx.set_option("dbg", (DBG << 16) if stc else 0)
        for name, key, fn in (("search_q4", "search_scan", lambda: ctx.search_rows(rows, 20)),
                              ("kmeans", "kmeans_assign", lambda: ctx.kmeans(k, 2, init, want_labels=False)),
                              ("cosmin", "assign_cosine_min", lambda: ctx.assign_cosine_min(init))):
            fn()
            ctx.profile_reset(); ctx.profile_enable(True)
            for _ in range(3):
                fn()
            ctx.profile_enable(False)
            e = ctx.profile()[key]
            ms = e["ms"] / e["launches"]
            st = ctx.tfs_stats()
            print(f"N={N} d={d} k={k} stream_tc={stc} {name:10s} {ms:7.3f} ms  {4.0 * N * d / ms * 1e-6:7.0f} GB/s   {st if stc else ''}", flush=True)
ctx.set_option("stream_tc", 1)
ctx.set_option("dbg", 0)
ctx.close()
