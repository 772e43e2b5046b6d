"""Phase costs of label_tc_kernel: time of one kmeans / cosine-min pass over 4M x 100 rows with phases switched off
(ganrev_set_option("dbg", bits << 8); results are invalid while a bit is set).  python tools/exp_label.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
ctx = pkg.Context(0)
N, d, k = 4_000_000, 100, 20
ctx.db_synthetic(N, d, seed=11)
init = np.random.default_rng(0).standard_normal(size=(k, d)).astype(np.float32)
init /= np.linalg.norm(init, axis=1, keepdims=True)
for mode in ("kmeans", "cosmin"):
    for bits, what in ((0, "full"), (1, "-sums"), (2, "-sort-sums"), (6, "-sort-sums-scan"), (14, "-sort-sums-scan-mma"), (30, "-all but loads")):
        ctx.set_option("dbg", bits << 8)
        fn = (lambda: ctx.kmeans(k, 1, init, want_labels=False)) if mode == "kmeans" else (lambda: ctx.assign_cosine_min(init))
        fn()
        ctx.profile_reset(); ctx.profile_enable(True)
        for _ in range(3):
            fn()
        ctx.profile_enable(False)
        pr = ctx.profile()
        e = pr["kmeans_assign" if mode == "kmeans" else "assign_cosine_min"]
        ms = e["ms"] / e["launches"]
        print(f"{mode:7s} {what:24s} {ms:7.3f} ms  {4.0 * N * d / ms * 1e-6:7.0f} GB/s")
ctx.set_option("dbg", 0)
ctx.close()
