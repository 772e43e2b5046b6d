import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
ctx = pkg.Context(0)
N, d, Q, k = 262144, 100, 4096, 20
db = np.random.default_rng(0).standard_normal(size=(N, d), dtype=np.float32)
ctx.db_set(db)
rows = (np.arange(1, Q + 1, dtype=np.int64) * 61)
for _ in range(2):
    ctx.profile_reset(); ctx.profile_enable(True)
    ctx.search_rows(rows, k)
    ctx.profile_enable(False)
    print({k2: round(v["ms"], 2) for k2, v in ctx.profile().items()})
