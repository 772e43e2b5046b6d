import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
ctx = pkg.Context(0)
N, d, k = 2_000_000, 100, 20
rng = np.random.default_rng(0)
x = rng.standard_normal(size=(N, d), dtype=np.float32)
ctx.db_set(x)
rows = np.array([99, 199, 299, 399], np.int64)
init = rng.standard_normal(size=(k, d), dtype=np.float32); init /= np.linalg.norm(init, axis=1, keepdims=True)
for _ in range(2):
    ctx.profile_reset(); ctx.profile_enable(True)
    ctx.search_rows(rows, 20)
    ctx.kmeans(k, 1, init, want_labels=False)
    ctx.assign_cosine_min(init)
    ctx.profile_enable(False)
    pr = ctx.profile()
    print({n: (round(v["ms"], 3), round(4.0 * N * d / (v["ms"] * 1e-3) * 1e-9)) for n, v in pr.items() if v["ms"] > 0.05})
