"""Host-side logic that needs no GPU: weight blob layout / init, shard arithmetic, the
torch.distributed (gloo, world size 2) plumbing used to bootstrap the library's communicator,
and the shard -> merge rule of the multi-GPU search checked with the oracle as the checker."""
import os
import socket

import numpy as np
import pytest


def test_blob_layout_roundtrip(pkg):
    Wt = pkg.weights
    for (C, H, W, nd) in [(1, 32, 32, 32), (3, 64, 64, 256)]:
        for lay, init in ((Wt.g_layout(C, H, W, nd), Wt.init_G), (Wt.r_layout(C, H, W, nd), Wt.init_R)):
            blob = init(C, H, W, nd, seed=9, stress=True)
            assert blob.dtype == np.float32 and blob.size == Wt.blob_floats(lay)
            p = Wt.unpack(blob, lay)
            np.testing.assert_array_equal(Wt.pack(p, lay), blob)
    # parameter counts (weights + biases, without BN) S: G 2.56M, R 4.65M
    p = Wt.unpack(Wt.init_G(1, 32, 32, 32), Wt.g_layout(1, 32, 32, 32))
    n = sum(v.size for k, v in p.items() if not k.startswith("bn"))
    assert abs(n - 2.56e6) < 0.02e6


def test_heuristic_init_distribution(pkg):
    """weight-init.lua:14-16, 70-72: U(+-1/sqrt(fan_in)) weights, every bias (and BN beta) zero."""
    Wt = pkg.weights
    p = Wt.unpack(Wt.init_R(1, 32, 32, 32, seed=2), Wt.r_layout(1, 32, 32, 32))
    for name, fan_in in (("c2.w", 576), ("c5.w", 1152), ("l1.w", 8192), ("l2.w", 512)):
        w = p[name]
        b = 1.0 / np.sqrt(fan_in)
        assert np.abs(w).max() <= b and np.abs(w).max() > 0.95 * b
        assert abs(w.std() - b / np.sqrt(3)) < 0.05 * b
    for k, v in p.items():
        if k.endswith(".b") and not k.startswith("bn") or k.startswith("bn") and k.endswith((".b", ".m")):
            assert not v.any(), k
        if k.startswith("bn") and k.endswith(".v"):
            assert (v == 1).all()
        if k.startswith("bn") and k.endswith(".g"):
            assert v.min() >= 0 and v.max() <= 1


def test_shard_range(pkg):
    sr = pkg.dist.shard_range
    for n in (0, 1, 7, 10_000, 1_000_000):
        for world in (1, 2, 3, 8):
            parts = [sr(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_noise_and_batch_helpers(pkg):
    nu = pkg.NN_UTILS
    z = nu.createNoiseInputs(1000, 32, "normal", np.random.default_rng(1))
    assert z.shape == (1000, 32) and z.dtype == np.float32 and abs(z.std() - 1) < 0.05
    u = nu.createNoiseInputs(1000, 32, "uniform", np.random.default_rng(1))
    assert u.min() >= -1 and u.max() <= 1
    with pytest.raises(ValueError):
        nu.createNoiseInputs(1, 1, "bogus")
    assert nu.toBatch(np.zeros((3, 4))).shape == (1, 3, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as td
    from __graft_entry__ import load_package
    from oracle import oracle as orc
    pkg = load_package()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the unique-id hand-off used by dist.init_comm (payload made on rank 0 only)
        payload = pkg.dist.broadcast_bytes(b"\x01\x02uid-128-bytes" * 8 if rank == 0 else None, src=0)
        assert payload == b"\x01\x02uid-128-bytes" * 8
        # 2. row-sharded search: local top-k on each shard + allgather + merge by (score desc, id asc)
        #    must equal the single-shard answer (the rule ganrev_search_cosine implements over NCCL)
        rng = np.random.default_rng(5)
        db = rng.normal(size=(999, 16)).astype(np.float32)
        db[500:520] = db[10]                          # ties across the shard boundary
        q = db[[10, 700]]
        k = 25
        lo, hi = pkg.dist.shard_range(999, world, rank)
        ids, sc = orc.search_cosine(db[lo:hi], q, k)
        ids = np.where(ids >= 0, ids + lo, -1)
        box = [None] * world
        td.all_gather_object(box, (ids, sc))
        all_ids = np.concatenate([b[0] for b in box], axis=1)
        all_sc = np.concatenate([b[1] for b in box], axis=1)
        merged = np.empty((2, k), np.int64)
        for r in range(2):
            order = sorted(range(all_ids.shape[1]), key=lambda j: (all_ids[r, j] < 0, np.isnan(all_sc[r, j]), -all_sc[r, j], all_ids[r, j]))
            merged[r] = all_ids[r, order[:k]]
        want, _ = orc.search_cosine(db, q, k)
        assert (merged == want).all()
        # 3. kmeans: per-shard int64 sums + allreduce(sum) == single-shard sums (order-free accumulators)
        shift = orc.kmeans_shift(db, 999)
        part = np.rint(db[lo:hi].astype(np.float64) * 2.0 ** shift).astype(np.int64).sum(0)
        import torch
        t = torch.from_numpy(part.copy())
        td.all_reduce(t)
        full = np.rint(db.astype(np.float64) * 2.0 ** shift).astype(np.int64).sum(0)
        assert (t.numpy() == full).all()
        ret[rank] = "ok"
    finally:
        td.destroy_process_group()


def test_gloo_world2_plumbing():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours) needs no GPU: one JSON line with the
    contract's keys, rank 0 only under torchrun."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-images", "32"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "g2r_images_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--ref-images", "32"],
                          capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert out1.returncode == 0 and out1.stdout.strip() == ""     # the other ranks exit 0 without work


def test_integration_doc_lists_every_entry_point():
    import os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = re.sub(r"/\*.*?\*/", " ", open(os.path.join(root, "include", "ganrev.h")).read(), flags=re.S)
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    names = set(re.findall(r"\b(ganrev_[a-z0-9_A-Z]+)\s*\(", hdr))
    missing = sorted(n for n in names if n not in doc)
    assert not missing, f"INTEGRATION.md does not mention {missing}"


def test_file_unique_id_handoff(pkg, tmp_path):
    """The NCCL unique-id hand-off without torch.distributed (mirrored by lua/ganrev.lua
    comm_init_file): rank 0 publishes the id atomically, later ranks poll for it, a missing file times out."""
    import threading

    class FakeCtx:
        def comm_unique_id(self):
            return bytes(range(128))

    path = str(tmp_path / "uid")
    got = {}
    t = threading.Thread(target=lambda: got.setdefault("r1", pkg.dist.file_unique_id(None, 2, 1, path, timeout_s=10)))
    t.start()
    assert pkg.dist.file_unique_id(FakeCtx(), 2, 0, path) == bytes(range(128))
    t.join()
    assert got["r1"] == bytes(range(128)) and not os.path.exists(path + ".tmp")
    with pytest.raises(TimeoutError):
        pkg.dist.file_unique_id(None, 2, 1, str(tmp_path / "never"), timeout_s=0.1)


def test_train_r_host_loop_and_torch_restatement(pkg):
    """train_r.py (mirror of train_r.lua:129-170) against a recording stand-in for the context: one init, nbBatches steps, masks
    in the module order include/ganrev.h documents, hyper-parameters passed through; and the PyTorch-CPU restatement used as
    the GPU tests' checker (oracle/torch_cpu.py) agrees with a finite difference of its own loss."""
    torch = pytest.importorskip("torch")
    C, H, W, nd, B = 1, 16, 16, 8, 4

    class Rec:
        def __init__(self):
            self.calls = []

        def train_R_init(self, *a, **k):
            self.calls.append(("init", a, k))

        def train_R_step(self, noise, masks, **k):
            self.calls.append(("step", noise.shape, [m.shape for m in masks], [m.dtype for m in masks], k))
            return 1.0 / len(self.calls), 0.0

        def train_R_state(self, what):
            return np.zeros(3, np.float32)

    rec = Rec()
    blob, losses = pkg.train_r.train(rec, (C, H, W), nd, nbBatches=3, batchSize=B, fixer=True, R_L2=2e-4, learningRate=5e-4)
    assert [c[0] for c in rec.calls] == ["init", "step", "step", "step"] and len(losses) == 3
    assert rec.calls[0][2] == {"tanh_out": False, "fixer": True}
    _, nshape, mshapes, mdtypes, hyper = rec.calls[1]
    assert nshape == (B, nd) and hyper == {"lr": 5e-4, "l1": 0.0, "l2": 2e-4, "clamp": 1.0}
    assert mshapes == [(B, C, H, W), (B, 64, H, W), (B, 64, H, W), (B, 64, H // 2, W // 2), (B, 128, H // 2, W // 2),
                       (B, 128, H // 2, W // 2), (B, 128), (B, 512)]
    assert all(d == np.uint8 for d in mdtypes)
    # the checker itself: d loss / d (one bias) by central differences
    from oracle.torch_cpu import train_masks, train_R_step
    rng = np.random.default_rng(0)
    rb = pkg.weights.init_R(C, H, W, nd, seed=2, stress=True)
    lay = pkg.weights.r_layout(C, H, W, nd)
    images = rng.random((B, C, H, W)).astype(np.float32)
    noise = rng.standard_normal((B, nd)).astype(np.float32)
    masks = train_masks(rng, B, C, H, W, False)
    loss, f, grads, _ = train_R_step(pkg, rb, C, H, W, nd, images, noise, masks, False, False, 0.0, 0.0, 0.0)
    assert f == pytest.approx(loss)
    off = 0
    for k, shp in lay:
        if k == "l2.b":
            break
        off += int(np.prod(shp))
    eps = 1e-2
    up, dn = rb.copy(), rb.copy()
    up[off] += eps; dn[off] -= eps
    lu = train_R_step(pkg, up, C, H, W, nd, images, noise, masks, False, False, 0.0, 0.0, 0.0)[0]
    ld = train_R_step(pkg, dn, C, H, W, nd, images, noise, masks, False, False, 0.0, 0.0, 0.0)[0]
    assert grads[off] == pytest.approx((lu - ld) / (2 * eps), rel=2e-2, abs=1e-5)
