"""Print the in-kernel clock64 timeline of CTA 0 for a tensor-core layer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
C, H, W, ND, N = 1, 32, 32, 100, 4096
ctx = pkg.Context(0)
ctx.load_G(C, H, W, ND, pkg.weights.init_G(C, H, W, ND))
ctx.load_R(0, C, H, W, ND, pkg.weights.init_R(C, H, W, ND))
noise = np.random.default_rng(0).normal(size=(N, ND)).astype(np.float32)
ctx.buffer_put(pkg._lib.BUF_NOISE, noise)
for layer in sys.argv[1:] or ["g_conv2_up"]:
    for dbg in (0, 15):
        ctx.set_option("dbg", dbg)
        ctx.forward_G(None, N=N, want_images=False); ctx.forward_R(0, None, N=N, want_attrs=False)
        ctx.trace_arm(layer)
        ctx.forward_G(None, N=N, want_images=False); ctx.forward_R(0, None, N=N, want_attrs=False)
        t = ctx.trace_read()
        t0 = t[t > 0].min()
        rel = np.where(t > 0, t - t0, -1)
        print(f"=== {layer} dbg={dbg}  (cycles since first event)")
        names = ["P:empty-ok", "P:issued", "M:full-ok", "M:committed", "E:tfull-ok", "E:released", "M:tempty-ok", "E:start-wait"]
        for r in (0, 1, 2, 3):
            print("  %-12s" % names[r], " ".join("%6d" % v for v in rel[r][:28]))
        for r in (6, 7, 4, 5):
            print("  %-12s" % names[r], " ".join("%6d" % v for v in rel[r][:10]))
        ev = rel[1][:200]; ev = ev[ev >= 0]
        if len(ev) > 20:
            print("  producer round period (steady): %.0f cycles" % np.diff(ev[8:]).mean())
        ev = rel[5][:60]; ev = ev[ev >= 0]
        if len(ev) > 6:
            print("  tile period (steady): %.0f cycles" % np.diff(ev[2:]).mean())
