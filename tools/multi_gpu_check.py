"""torchrun worker: sharded search / kmeans over NCCL vs the single-shard oracle (bit-exact)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as td
from __graft_entry__ import load_package
from oracle import oracle as orc

pkg = load_package()
world, rank, local = pkg.dist.env_world()
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = pkg.Context(local)
pkg.dist.init_comm(ctx)

rng = np.random.default_rng(11)
N, d, Q, k = 20011, 100, 70, 20
db = rng.normal(size=(N, d)).astype(np.float32)
db[15000:15040] = db[7]                      # ties across the shard boundary
q = np.concatenate([db[[7, 12000]], rng.normal(size=(Q - 2, d)).astype(np.float32)])
lo, hi = pkg.dist.shard_range(N, world, rank)
ctx.db_set(db[lo:hi])
ids, sc = ctx.search_cosine(q, k)
want_ids, want_sc = orc.search_cosine(db, q, k)
assert (ids == want_ids).all(), "sharded top-k ids differ"
assert (sc.view(np.uint32) == want_sc.view(np.uint32)).all(), "sharded top-k scores differ"

kk, niter = 20, 6
init = rng.normal(size=(kk, d)).astype(np.float32)
init /= np.linalg.norm(init, axis=1, keepdims=True)
cen, tot, lab = ctx.kmeans(kk, niter, init)
want_c, want_t, want_l = orc.kmeans(db, kk, niter, init)
assert (lab == want_l[lo:hi]).all(), "sharded kmeans labels differ"
assert (cen.view(np.uint32) == want_c.view(np.uint32)).all(), "sharded kmeans centroids differ"
assert (tot == want_t).all()
td.barrier()
if rank == 0:
    print("multi-gpu check ok: world", world)
ctx.close()
td.destroy_process_group()
