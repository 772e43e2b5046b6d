"""A/B a ganrev_set_option knob on one box: per-kernel CUDA-event times of G->R over 32768 faces for each value.
usage: [GEOM=C,H,W,nd,N] python tools/ab_option.py tma_store 0 1"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
name, values = sys.argv[1], [int(v) for v in sys.argv[2:]]
C, H, W, ND, N = [int(v) for v in os.environ.get("GEOM", "1,32,32,100,32768").split(",")]
noise = np.random.default_rng(0).normal(size=(N, ND)).astype(np.float32)
for rep in range(2):
    for v in values:
        ctx = pkg.Context(0)
        ctx.set_option(name, v)
        ctx.load_G(C, H, W, ND, pkg.weights.init_G(C, H, W, ND))
        ctx.load_R(0, C, H, W, ND, pkg.weights.init_R(C, H, W, ND))
        ctx.buffer_put(pkg._lib.BUF_NOISE, noise)
        for _ in range(2):
            ctx.forward_G(None, N=N, want_images=False); ctx.forward_R(0, None, N=N, want_attrs=False)
        ctx.profile_reset(); ctx.profile_enable(True)
        for _ in range(3):
            ctx.forward_G(None, N=N, want_images=False); ctx.forward_R(0, None, N=N, want_attrs=False)
        ctx.profile_enable(False)
        pr = ctx.profile()
        tot = sum(e["ms"] for e in pr.values())
        print(f"{name}={v}: total {tot:.2f} ms  " + "  ".join(f"{k} {e['ms']:.2f}" for k, e in pr.items() if e["ms"] > 0.5))
        ctx.close()
