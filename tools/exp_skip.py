"""Timing experiment: per-layer time with A and/or B operand loads skipped (garbage results)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
C, H, W, ND, N = 1, 32, 32, 100, 32768
ctx = pkg.Context(0)
if "GANREV_CTA_PAIRS" in os.environ:
    ctx.set_option("cta_pairs", int(os.environ["GANREV_CTA_PAIRS"]))
ctx.load_G(C, H, W, ND, pkg.weights.init_G(C, H, W, ND))
ctx.load_R(0, C, H, W, ND, pkg.weights.init_R(C, H, W, ND))
noise = np.random.default_rng(0).normal(size=(N, ND)).astype(np.float32)
ctx.buffer_put(pkg._lib.BUF_NOISE, noise)
for dbg in (0,):
    ctx.set_option("dbg", dbg)
    for rep in range(2):
        ctx.profile_reset(); ctx.profile_enable(True)
        ctx.forward_G(None, N=N, want_images=False)
        ctx.forward_R(0, None, N=N, want_attrs=False)
        ctx.profile_enable(False)
    prof = ctx.profile()
    print("dbg", dbg, " ".join("%s=%.2f" % (k.replace("_", "")[:9], v["ms"]) for k, v in prof.items() if v["ms"] > 0.5))
