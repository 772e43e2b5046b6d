"""A/B a ganrev_set_option knob by the WALL time of the resident G->R chain (per-kernel profiling off: event records between
launches would defeat e.g. programmatic dependent launch).  usage: [GEOM=C,H,W,nd,N] python tools/ab_total.py pdl 0 1"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
name, values = sys.argv[1], [int(v) for v in sys.argv[2:]]
C, H, W, ND, N = [int(v) for v in os.environ.get("GEOM", "1,32,32,100,65536").split(",")]
noise = np.random.default_rng(0).normal(size=(N, ND)).astype(np.float32)
for rep in range(3):
    for v in values:
        ctx = pkg.Context(0)
        ctx.set_option(name, v)
        ctx.load_G(C, H, W, ND, pkg.weights.init_G(C, H, W, ND))
        ctx.load_R(0, C, H, W, ND, pkg.weights.init_R(C, H, W, ND))
        ctx.buffer_put(pkg._lib.BUF_NOISE, noise)
        for _ in range(2):
            ctx.forward_G(None, N=N, want_images=False); ctx.forward_R(0, None, N=N, want_attrs=False)
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(5):
            ctx.forward_G(None, N=N, want_images=False); ctx.forward_R(0, None, N=N, want_attrs=False)
        ctx.sync()
        dt = (time.perf_counter() - t0) / 5
        print(f"{name}={v}: {dt * 1e3:.2f} ms per G->R over {N} faces = {N / dt:.0f} faces/s")
        ctx.close()
