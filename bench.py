#!/usr/bin/env python
"""bench.py -- G->R reversal + cosine top-k search throughput (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (libganrev_cuda.so)
  python bench.py --impl reference --gpus N ...             # CPU arm: the oracle port on host cores

Workload (BASELINE.json configs[3]): per GPU, `--images` (default 1,000,000) generated 32x32
grayscale faces from 100-dim N(0,1) noise, batched G -> R, then a 4096-query cosine top-20
search over the recovered vectors (row-sharded, one NCCL allgather + merge when N > 1).
One step = one such pass.  Weak scaling: every rank works on its own `--images` faces.

`value`  : images/s with the noise already resident in HBM (device-timed, CUDA events on the
           library's stream, max over ranks).
`e2e`    : the same through the C ABI with HOST buffers (pinned noise in, recovered vectors and
           top-k ids/scores out), copies inside the timed region.
Synthetic data, random-init ("heuristic", weight-init.lua) weights.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C, H, W, ND = 1, 32, 32, 100
Q, TOPK = 4096, 20
# SURVEY.md section 8d / BASELINE.md: algorithmic FLOPs per image, direct formulation, nd=100
FLOP_G, FLOP_R = 1.2169e9, 0.3494e9
# direct-form FLOPs per image of the two dominant layers (models.lua:121-122, 127-128)
FLOP_DIRECT = {"g_conv1_up": 2 * 302.0e6, "g_conv2_up": 2 * 302.0e6}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1353.9), d.get("hbm_gbs", 6552.0), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port(n_img, n_db, n_q, threads=None):
    """Time the CPU oracle (kind 'port') on a bounded sample: G->R over n_img faces, then n_q
    queries top-20 over n_db recovered rows.  Returns (images/s, queries/s, cores, description)."""
    from oracle import oracle as orc
    from __graft_entry__ import load_package
    pkg = load_package()
    if threads:
        orc.set_num_threads(threads)
    cores = orc.num_threads()
    gb = pkg.weights.init_G(C, H, W, ND, seed=1)
    rb = pkg.weights.init_R(C, H, W, ND, seed=2)
    noise = np.random.default_rng(7).normal(size=(n_img, ND)).astype(np.float32)
    orc.forward_G(gb, C, H, W, ND, noise[:cores])           # warm the threads / caches
    t0 = time.perf_counter()
    img = orc.forward_G(gb, C, H, W, ND, noise)
    att = orc.forward_R(rb, C, H, W, ND, img)
    t_gr = time.perf_counter() - t0
    db = np.random.default_rng(8).normal(size=(n_db, ND)).astype(np.float32)
    db[: min(n_img, n_db)] = att[: min(n_img, n_db)]
    qs = db[np.arange(n_q) * max(1, n_db // n_q)]
    t0 = time.perf_counter()
    orc.search_cosine(db, qs, TOPK)
    t_s = time.perf_counter() - t0
    sample = f"G->R over {n_img} faces ({t_gr:.1f} s) + {n_q}-query top-{TOPK} over {n_db} rows ({t_s:.1f} s), oracle C port, OpenMP"
    return n_img / t_gr, n_q / t_s, cores, sample, t_gr, t_s


def hbm_kernels(ctx, pkg, peak_gbs):
    """Bandwidth rooflines of the HBM-bound kernels of the path (north_star: L2, kmeans, small-Q search), each on
    inputs larger than L2, timed with CUDA events on the library stream (ganrev_profile_*)."""
    out = {}
    rng = np.random.default_rng(11)

    def timed(name, fn, alg_bytes, reps):
        fn()                                                       # warm-up
        ctx.profile_reset(); ctx.profile_enable(True)
        for _ in range(reps):
            fn()
        ctx.profile_enable(False)
        e = ctx.profile()[name]
        ms = e["ms"] / max(e["launches"], 1)
        gbs = alg_bytes / (ms * 1e-3) * 1e-9
        return {"bound": "hbm", "achieved": gbs, "peak": peak_gbs, "unit": "GB/s", "frac": gbs / peak_gbs, "ms_per_launch": ms}

    # torch.dist over 200k pairs of 32x32 faces (apply_r.lua:366): 8*C*H*W algorithmic bytes per pair
    n = 200_000
    a = rng.random((n, 1024), dtype=np.float32)
    b = a[::-1].copy()
    out["l2_pairs"] = timed("l2_pairs", lambda: ctx.l2(a, b), 8.0 * n * 1024, 2)
    out["l2_pairs"]["workload"] = f"{n} pairs of 32x32 fp32 faces, 8*C*H*W algorithmic bytes per pair"
    # nearest training image by torch.dist for 8 query faces (sample.lua:128-148): the set is read once
    out["nearest_l2_q8"] = timed("nearest_l2", lambda: ctx.nearest_l2(b[:8], a), 4.0 * n * 1024, 3)   # explicit set: independent of --geom
    out["nearest_l2_q8"]["workload"] = (f"8 query faces against {n} 32x32 fp32 faces, 4*C*H*W algorithmic bytes per set row; "
                                        "the canonical fp32-square / fp64-sum arithmetic needs one fp32->fp64 conversion per (query, element), "
                                        "so with 8 queries per pass the conversion unit (16 lanes/clk/SM), not HBM, is the nearer bound")
    # kmeans k=20 over 4M x 100 rows (recovered-vector shape): 4*N*d algorithmic bytes per iteration
    N, d, k = 4_000_000, ND, 20
    x = rng.standard_normal(size=(N, d), dtype=np.float32)
    init = rng.standard_normal(size=(k, d), dtype=np.float32)
    init /= np.linalg.norm(init, axis=1, keepdims=True)
    ctx.db_set(x)
    out["kmeans_assign_k20"] = timed("kmeans_assign", lambda: ctx.kmeans(k, 2, init, want_labels=False), 4.0 * N * d, 2)
    out["kmeans_assign_k20"]["workload"] = f"k=20 over {N} x {d} fp32 rows, 4*N*d algorithmic bytes per iteration"
    # cosine top-20 for 4 needles (BASELINE config 1 shape) over the same rows: 4*N*d algorithmic bytes
    rows = np.array([99, 199, 299, 399], np.int64)
    out["search_q4"] = timed("search_scan", lambda: ctx.search_rows(rows, TOPK), 4.0 * N * d, 3)
    out["search_q4"]["workload"] = f"4 needles, top-{TOPK}, over {N} x {d} fp32 rows, 4*N*d algorithmic bytes"
    return out


def run_reference(args, world, rank):
    """--impl reference: the reference's CPU implementation of the path on host cores.  Torch7
    cannot run here (no lua/luajit/th, un-vendored rocks), so this is the oracle port."""
    if rank != 0:
        return
    n_img, n_db, n_q = args.ref_images, 20000, 64
    vals, qps = [], []
    for i in range(args.warmup + args.steps):
        ips, q, cores, sample, t_gr, t_s = cpu_port(n_img, n_db, n_q)
        if i >= args.warmup:
            vals.append(ips); qps.append(q)
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "g2r_images_per_sec", "value": v, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_img / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[3]: 32x32 gray, nd=100, G->R + 4096-query cosine top-20", "images_per_step_sample": n_img,
                   "note": "CPU arm: bounded sample of the same per-image work; Torch7 itself cannot run here"},
        "search_queries_per_sec": float(np.mean(qps)),
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=1_000_000, help="faces per GPU per step")
    ap.add_argument("--chunk", type=int, default=0, help="pipeline chunk (0 = library default)")
    ap.add_argument("--ref-images", type=int, default=1536, help="CPU-arm sample size per step (~8 s of 16-thread CPU work)")
    ap.add_argument("--cpu-images", type=int, default=2048, help="cpu_baseline sample size (~10 s of 16-thread CPU work)")
    ap.add_argument("--geom", default="", help="C,H,W,nd of another geometry (e.g. 3,64,64,256 = BASELINE configs[4]'s G/R); default configs[3]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hbm-kernels", action="store_true", help="skip the L2 / kmeans / small-Q search bandwidth rooflines")
    args = ap.parse_args()

    global C, H, W, ND, FLOP_G, FLOP_R, FLOP_DIRECT
    workload_name = "configs[3]: G->R reversal of 32x32 grayscale faces (nd=100) + 4096-query cosine top-20 over the recovered set"
    if args.geom:
        C, H, W, ND = (int(v) for v in args.geom.split(","))
        px = H * W
        # direct-form MACs per image: Linear, Up+Conv 512->256 @ (H/2)^2, Up+Conv 256->128 @ H*W, Conv 128->C
        FLOP_G = 2.0 * (ND * 512 * px / 16 + (px / 4) * 256 * 4608 + px * 128 * 2304 + px * C * 1152)
        FLOP_R = 2.0 * (px * 64 * 9 * C + 2 * px * 64 * 576 + (px / 4) * (128 * 576 + 2 * 128 * 1152) + (128 * px / 16) * 512 + 512 * ND)
        FLOP_DIRECT = {"g_conv1_up": 2.0 * (px / 4) * 256 * 4608, "g_conv2_up": 2.0 * px * 128 * 2304}
        workload_name = f"G->R reversal of {C}x{H}x{W} faces (nd={ND}) + 4096-query cosine top-20 over the recovered set"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, world, rank)
        return

    import torch
    import torch.distributed as td
    from __graft_entry__ import load_package
    pkg = load_package()

    torch.cuda.set_device(local_rank)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = pkg.Context(local_rank)
    if world > 1:
        pkg.dist.init_comm(ctx)
    if args.chunk:
        ctx.set_option("chunk", args.chunk)
    N = args.images
    ctx.load_G(C, H, W, ND, pkg.weights.init_G(C, H, W, ND, seed=1))
    ctx.load_R(0, C, H, W, ND, pkg.weights.init_R(C, H, W, ND, seed=2))

    # synthetic inputs: N(0,1) noise (seed 7 + rank), pinned on the host
    noise_t = torch.empty((N, ND), dtype=torch.float32).pin_memory()
    noise = noise_t.numpy()
    rng = np.random.default_rng(7 + rank)
    for lo in range(0, N, 1 << 18):
        hi = min(N, lo + (1 << 18))
        noise[lo:hi] = rng.standard_normal(size=(hi - lo, ND), dtype=np.float32)
    attrs_t = torch.empty((N, ND), dtype=torch.float32).pin_memory()
    attrs = attrs_t.numpy()
    qrows = (np.arange(1, Q + 1) * max(1, N // (Q + 1))).clip(0, N - 1)     # rows i*244 at N = 1M (SURVEY 8d cfg 4)

    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))

    qrows64 = qrows.astype(np.int64)     # needles = recovered rows of rank 0's shard (global ids: rank 0 owns [0, N))

    def step_resident():
        ctx.forward_G(None, N=N, want_images=False)
        ctx.forward_R(0, None, N=N, want_attrs=False)
        ctx.db_set(None, N=N, d=ND)
        return ctx.search_rows(qrows64, TOPK)                 # queries gathered on the device from the database rows

    def step_e2e():
        ctx.forward_G(noise, want_images=False)              # H2D noise
        ctx.forward_R(0, None, N=N, out=attrs)               # D2H recovered vectors
        ctx.db_set(None, N=N, d=ND)
        q = np.ascontiguousarray(attrs[qrows]) if rank == 0 else None
        if world > 1:
            box = [q]
            td.broadcast_object_list(box, src=0)
            q = box[0]
        return ctx.search_cosine(q, TOPK)                    # H2D queries, D2H ids + scores

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        e0.record(stream)
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            td.all_reduce(t, op=td.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ctx.launch_count() - l0, out

    ctx.buffer_put(pkg._lib.BUF_NOISE, noise)
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.profile_reset()
    ctx.profile_enable(True)
    ms_res, launches, _ = timed(step_resident, args.steps)
    ctx.profile_enable(False)
    prof = ctx.profile()
    clocks = sampler.stop()
    for _ in range(max(1, args.warmup // 3)):
        step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)

    if rank != 0:
        ctx.close()
        if world > 1:
            td.destroy_process_group()
        return

    total_images = N * world * args.steps
    value = total_images / (ms_res * 1e-3)
    e2e = total_images / (ms_e2e * 1e-3)
    peak_tf, peak_gbs, peak_src = load_peaks()
    # dominant kernel = the conv GEMM with the largest share of device time
    conv_names = [k for k in prof if k.startswith(("g_conv1", "g_conv2", "r_conv", "g_linear", "r_linear"))]
    dom = max(conv_names, key=lambda k: prof[k]["ms"]) if conv_names else None
    roofline = None
    if dom:
        e = prof[dom]
        ms_launch = e["ms"] / max(e["launches"], 1)
        imgs_launch = N * args.steps / max(e["launches"], 1)
        exec_tf = e["flops"] / max(e["ms"], 1e-9) * 1e-9
        alg = FLOP_DIRECT.get(dom)
        alg_tf = (alg * imgs_launch) / (ms_launch * 1e-3) * 1e-12 if alg else exec_tf
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01", "ncu_traffic.json")))
            if dom in tj:
                traffic = tj[dom]["dram_bytes_per_image"] * imgs_launch     # bytes per launch, from the committed ncu capture
        except Exception:
            pass
        roofline = {"bound": "tensor", "kernel": dom, "achieved": alg_tf, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": alg_tf / peak_tf, "traffic": traffic,
                    "achieved_executed": exec_tf, "frac_executed": exec_tf / peak_tf,
                    "note": "achieved = direct-form (algorithmic) FLOPs of the layer / CUDA-event time per launch; "
                            "executed = FLOPs actually issued (upsample folded into 4 phase convs = 2.25x fewer); peak = " + peak_src,
                    "avg_launch_ms": ms_launch, "images_per_launch": imgs_launch}
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    kernels = {k: {"launches": v["launches"], "ms": round(v["ms"], 3), "share": round(v["ms"] / tot_ms, 4),
                   "tflops_executed": round(v["flops"] / max(v["ms"], 1e-9) * 1e-9, 2),
                   "gbs": round(v["bytes"] / max(v["ms"], 1e-9) * 1e-6, 1)} for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    srch_ms = sum(prof[k]["ms"] for k in prof if k.startswith("search")) / args.steps
    line = {
        "metric": "g2r_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name,
                   "images_per_gpu": N, "queries": Q, "top_k": TOPK, "noise_dim": ND, "weights": "random-init (weight-init.lua heuristic)",
                   "l2_flush": "inputs larger than L2 (activation stream per chunk >> 126 MB)", "arith": "bf16 operands, fp32 accumulate (conv GEMMs); fp32 fmaf (search)"},
        "search_queries_per_sec": Q * world / (srch_ms * 1e-3) if srch_ms > 0 else None,
        "algorithmic_tflops": value * (FLOP_G + FLOP_R) * 1e-12,
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": int(N * ND * 4 + Q * ND * 4),
                "d2h_bytes_per_step": int(N * ND * 4 + Q * TOPK * 12), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kernels,
    }
    if world == 1 and not args.no_hbm_kernels:
        line["hbm_kernels"] = hbm_kernels(ctx, pkg, peak_gbs)
    if world == 1 and not args.no_cpu_baseline:
        ips, qps, cores, sample, _, _ = cpu_port(args.cpu_images, 20000, 64)
        line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample,
                                "search_queries_per_sec": qps}
    print(json.dumps(line))
    ctx.close()
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
