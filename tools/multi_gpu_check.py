"""torchrun worker: every row-sharded operation over NCCL vs the single-shard oracle (bit-exact).

search (allgather + merge), kmeans (int64 allreduce), cosine-min assignment + per-cluster top-71 + mean faces
(allgather-merge with Q = k clusters, member images assembled across ranks), anomaly quantile (radix select with a
histogram allreduce per pass), nearest-L2 over a sharded set.  The library's NCCL communicator is bootstrapped with
the FILE-based unique-id hand-off (dist.init_comm_file), the protocol lua/ganrev.lua uses for one process per GPU.
Prints one JSON line on rank 0: {"sharded_parity": true, ...}.
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as td
from __graft_entry__ import load_package
from oracle import oracle as orc


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a.view(np.uint64)


def run_checks(pkg, ctx, world, rank, seed=11):
    """Returns a dict of named booleans; every rank computes the same thing."""
    res = {}
    rng = np.random.default_rng(seed)
    N, d, Q, k = 20011, 100, 70, 20
    db = rng.normal(size=(N, d)).astype(np.float32)
    db[15000:15040] = db[7]                      # ties across the shard boundary
    q = np.concatenate([db[[7, 12000]], rng.normal(size=(Q - 2, d)).astype(np.float32)])
    lo, hi = pkg.dist.shard_range(N, world, rank)
    ctx.db_set(db[lo:hi])
    ids, sc = ctx.search_cosine(q, k)
    want_ids, want_sc = orc.search_cosine(db, q, k)
    res["search_ids"] = bool((ids == want_ids).all())
    res["search_scores"] = bool((bits(sc) == bits(want_sc)).all())
    bounds = [pkg.dist.shard_range(N, world, r) for r in range(world)]                 # the same needle list on every rank
    rows = np.array([7, 12000, N - 1] + [b[0] for b in bounds] + [b[1] - 1 for b in bounds], np.int64)
    ids_r, sc_r = ctx.search_rows(rows, k)
    wi, ws = orc.search_cosine(db, db[rows], k)
    res["search_rows"] = bool((ids_r == wi).all() and (bits(sc_r) == bits(ws)).all())

    kk, niter = 20, 6
    init = rng.normal(size=(kk, d)).astype(np.float32)
    init /= np.linalg.norm(init, axis=1, keepdims=True)
    cen, tot, lab = ctx.kmeans(kk, niter, init)
    want_c, want_t, want_l = orc.kmeans(db, kk, niter, init)
    res["kmeans_labels"] = bool((lab == want_l[lo:hi]).all())
    res["kmeans_centroids"] = bool((bits(cen) == bits(want_c)).all() and (tot == want_t).all())

    # cosine-min assignment, per-cluster top-71 and mean faces (apply_r.lua:206-243) over sharded rows + images
    px, m = 48, 71
    images = rng.random((N, px)).astype(np.float32)
    cl, cv = ctx.assign_cosine_min(want_c)
    want_cl, want_cv = orc.assign_cosine_min(db, want_c)
    res["assign"] = bool((cl == want_cl[lo:hi]).all() and (bits(cv) == bits(want_cv[lo:hi])).all())
    mids, mcnt, mean = ctx.cluster_members(kk, m, images[lo:hi])
    w_ids, w_cnt, w_mean = orc.cluster_members(want_cl, want_cv, kk, m, images)
    ok = w_cnt > 0
    res["cluster_members"] = bool((mids == w_ids).all() and (mcnt == w_cnt).all())
    res["cluster_means"] = bool((bits(mean[ok]) == bits(w_mean[ok])).all() and np.isnan(mean[~ok]).all())

    # anomaly quantile over sharded distances (apply_r.lua:370-378)
    l2 = np.abs(rng.normal(size=N)) * 3.0
    l2[:500] = l2[500:1000]
    flags, thr = ctx.anomaly_flags(l2[lo:hi], hi - lo, hi - lo, 0.15)
    w_flags, w_thr = orc.anomaly_flags(l2, N, N, 0.15)
    res["anomaly"] = bool(np.float64(thr).view(np.uint64) == np.float64(w_thr).view(np.uint64) and (flags == w_flags[lo:hi]).all())

    # nearest training image over a sharded set (sample.lua:128-148), incl. the "row 0 sticks" quirk
    ts = rng.random((3001, 256)).astype(np.float32)
    qq = (ts[rng.integers(0, 3001, size=9)] + rng.normal(scale=0.02, size=(9, 256))).astype(np.float32)
    ts[2900] = ts[3]; qq[0] = ts[3]
    lo2, hi2 = pkg.dist.shard_range(3001, world, rank)
    nid, nd_ = ctx.nearest_l2(qq, ts[lo2:hi2])
    oi, od = orc.nearest_l2(qq, ts)
    res["nearest_l2"] = bool((nid == oi).all() and (bits(nd_) == bits(od)).all())
    ts_nan = ts.copy(); ts_nan[0, 1] = np.nan
    nid, nd_ = ctx.nearest_l2(qq, ts_nan[lo2:hi2])
    res["nearest_l2_row0_nan"] = bool((nid == 0).all() and np.isnan(nd_).all())
    return res


def main():
    pkg = load_package()
    world, rank, local = pkg.dist.env_world()
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pkg.Context(local)
    uid_path = os.path.join(tempfile.gettempdir(), f"ganrev_uid_{os.environ.get('MASTER_PORT', '0')}")
    if rank == 0 and os.path.exists(uid_path):
        os.remove(uid_path)
    td.barrier()
    pkg.dist.init_comm_file(ctx, world, rank, uid_path)
    res = run_checks(pkg, ctx, world, rank)
    flag = torch.tensor([1 if all(res.values()) else 0], device="cuda")
    td.all_reduce(flag, op=td.ReduceOp.MIN)
    ok = bool(flag.item())
    if rank == 0:
        print(json.dumps({"sharded_parity": ok, "world": world, "checks": res}))
        if ok:
            print("multi-gpu check ok: world", world)
    ctx.close()
    td.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
