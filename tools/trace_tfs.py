"""clock64 timeline of CTA 0 of the tf32 filter pipeline (stream_tc.cuh): python tools/trace_tfs.py [search|kmeans|cosmin] [d] [k]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
mode = sys.argv[1] if len(sys.argv) > 1 else "search"
d = int(sys.argv[2]) if len(sys.argv) > 2 else 100
k = int(sys.argv[3]) if len(sys.argv) > 3 else 20
N = 4_000_000 * 100 // d
ctx = pkg.Context(0)
ctx.db_synthetic(N, d, seed=11)
init = np.random.default_rng(0).standard_normal(size=(k, d)).astype(np.float32)
init /= np.linalg.norm(init, axis=1, keepdims=True)
rows = np.array([99, 199, 299, 399], np.int64)
fn = {"search": lambda: ctx.search_rows(rows, 20), "kmeans": lambda: ctx.kmeans(k, 1, init, want_labels=False), "cosmin": lambda: ctx.assign_cosine_min(init)}[mode]
fn()
ctx.trace_arm("tfs")
fn()
t = ctx.trace_read()
t0 = t[t > 0].min()
rel = np.where(t > 0, t - t0, -1)
np.set_printoptions(linewidth=250)
print(f"=== {mode} d={d} k={k}: cycles since the first event, CTA 0")
print("P tile issued   ", rel[0][:40])
print("M tile committed", rel[1][:40])
for g in (0, 1):
    r = rel[2 + g]
    if (r >= 0).any():
        print(f"E{g} acc seen     ", r[0:80:2])
        print(f"E{g} done         ", r[1:80:2])
if (rel[4] >= 0).any():
    print("S labels seen   ", rel[4][0:80:2])
    print("S done          ", rel[4][1:80:2])
for name, ev in (("P", rel[0]), ("M", rel[1]), ("E0 done", rel[2][1::2]), ("E1 done", rel[3][1::2]), ("S done", rel[4][1::2])):
    ev = ev[ev >= 0]
    if len(ev) > 40:
        print(f"{name}: steady period {np.diff(ev[20:]).mean():.0f} cycles per event; busy E/S = done - seen:", end=" ")
        print()
for name, r in (("E0", rel[2]), ("E1", rel[3]), ("S", rel[4])):
    seen, done = r[0::2], r[1::2]
    ok = (seen >= 0) & (done >= 0)
    if ok.sum() > 40:
        print(f"{name}: mean busy {np.mean((done - seen)[ok][20:]):.0f} cycles per tile, mean wait {np.mean((seen[1:] - done[:-1])[ok[1:] & ok[:-1]][20:]):.0f}")


if (rel[5] >= 0).sum() > 64:
    ph = rel[5][:256].reshape(16, 16)
    print("S per warp: cycles from 'labels seen' to 'my rows are summed' (warps 0..7), then to 'barrier passed'")
    for n in range(4, 16):
        print("   tile", n, " ".join("%6d" % (ph[n, w] - ph[n, 9]) for w in range(8)), "| %6d" % (ph[n, 8] - ph[n, 9]))

if (rel[6] >= 0).sum() > 64:
    ph = rel[6][:256].reshape(32, 8)
    print("E0 phases, cycles since the accumulator was released: candidates known, pairs listed, chains done, labels out, bucket buffer free")
    for n in range(6, 18):
        print("   tile", n, " ".join("%6d" % (ph[n, e] - ph[n, 0] if ph[n, e] >= 0 else -1) for e in range(1, 6)), "  (acc seen -> released %d)" % (ph[n, 0] - rel[2][2 * n]))
