"""Epilogue phase timeline (needs the library built with -DGANREV_EPI_TRACE): per chunk of epilogue
thread 0: accumulators landed / math done / staged in smem / stores issued."""
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
for layer in sys.argv[1:] or ["r_conv2"]:
    ctx.forward_G(None, N=N, want_images=False); ctx.forward_R(0, None, N=N, want_attrs=False)
    ctx.trace_arm(layer)
    ctx.forward_G(None, N=N, want_images=False); ctx.forward_R(0, None, N=N, want_attrs=False)
    t = ctx.trace_read()
    t0 = t[t > 0].min()
    rel = np.where(t > 0, t - t0, -1)
    print(f"=== {layer} epilogue thread 0, per item: [E:start-wait, tfull-ok | per chunk: landed, math, staged, stored | released]")
    for it in range(2, 8):
        ch = []
        for c in range(4):
            if rel[0][it * 4 + c] >= 0:
                ch.append("(%d %d %d %d)" % tuple(rel[r][it * 4 + c] for r in range(4)))
        print("  item %d: wait %d tfull %d  %s  released %d" % (it, rel[7][it], rel[4][it], " ".join(ch), rel[5][it]))
