// nearest_tc.cuh -- nearest training image by torch.dist (sample.lua:128-148) as a TMA -> tf32 tcgen05 filter pipeline.
//
// The exact kernel (layers.cuh nearest_l2_kernel) pays one fp32->fp64 conversion and one DADD per (query, pixel): conversion-unit
// bound at 0.12 of HBM.  Here, as in stream_tc.cuh, the tensor core looks at every (row, query) pair and the canonical
// arithmetic (fp32 difference, fp32 square, double sums in the oracle's lane order, double sqrt) runs only for the rows that
// can still be the nearest:
//   * warp 0 streams the set through a ring of 16 KB slots with TMA (box = [128 rows][32 fp32 pixels], 128B-swizzled);
//   * warp 1 issues tcgen05.mma.kind::tf32 (M = 128 rows, N = 16 query columns, K = 8): x.q for every pair, within
//     tfs_eps(px) |x||q| of the real dot product (stream_tc.cuh);
//   * warps 2-5 add up |x|^2 of their row from the same slots (any order: it only feeds bounds), then the slot goes back to
//     the producer;
//   * warps 6-9 (thread = row = TMEM lane): with s = |x|^2 + |q|^2 and c = tfs_eps(px) + 2e-4 (AM-GM: 2|x||q| <= s),
//         (1 - c) s - 2 x.q  <=  canonical squared distance  <=  (1 + c) s - 2 x.q .
//     The smallest UPPER bound seen so far (per query, shared between blocks through one atomicMin word) bounds the nearest
//     distance; a row is a candidate iff its LOWER bound does not exceed it (ties included; NaN / inf always; row 0 always --
//     the reference takes it unconditionally).  Candidates are queued and evaluated in batches: one warp per (row, query) runs
//     the canonical arithmetic on the row from global memory, tightens the bound with the exact value and keeps the warp's best
//     (distance, row) per query.  The per-warp records go through nearest_l2_merge_kernel like those of the exact kernel.
// Results are bit-identical to nearest_l2_kernel / the oracle (tests/test_gpu_l2.py runs both).
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"
#include "layers.cuh"
#include "stream_tc.cuh"

namespace ganrev {
namespace ntc {

using namespace tc;
using tfs::kRows; using tfs::kBoxCols; using tfs::kSlotBytes; using tfs::kMaxSlots; using tfs::kAcc;

constexpr int kNQP = 16;                        // query columns (UMMA N); NL2_QB = 8 queries per pass
constexpr int kThreads = 64 + 128 + 128;        // producer, MMA, 4 norm warps, 4 epilogue warps
constexpr int kQueue = 512;
constexpr int kTrig = 64;                       // queued (row, query) pairs that start a batch of canonical evaluations
constexpr int kBar = 512;

struct NParams {
    const float* set;                           // [n_rows][px]
    long long n_rows;
    int px;
    const float* q;                             // [Q][px] on the device
    int q0, nq;                                 // this pass: queries q0 .. q0 + nq - 1, nq <= NL2_QB
    NearestRec* partial;                        // [grid * 4][NL2_QB]
    unsigned char* row0_nan;                    // [Q]
    unsigned* gthr;                             // [NL2_QB] smallest upper bound published by any block (float bits), +inf before the launch
    unsigned long long* stats;                  // [0] canonical evaluations
    long long n_tiles;
    int nbox, nslots;
    int* err_flag;
};

__global__ void nearest_init_kernel(unsigned* gthr) { if (threadIdx.x < NL2_QB) gthr[threadIdx.x] = 0x7f800000u; }

// torch.dist(set[row], query) in the oracle's canonical order: element i -> lane (i/4) % 32, ascending i per lane, fp32 difference
// and square, double sums, xor-butterfly, double sqrt.  Every lane returns the squared sum t (dist = sqrt(t)).
__device__ __forceinline__ double canonical_sq(const float* __restrict__ x, const float* __restrict__ y, int px, int lane) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    double s = 0.0;
    const int n4 = px >> 2;
    for (int i = lane; i < n4; i += 32) {
        const float4 xv = __ldg(x4 + i), yv = __ldg(y4 + i);
        const float d0 = __fsub_rn(xv.x, yv.x), d1 = __fsub_rn(xv.y, yv.y), d2 = __fsub_rn(xv.z, yv.z), d3 = __fsub_rn(xv.w, yv.w);
        s += static_cast<double>(__fmul_rn(d0, d0));
        s += static_cast<double>(__fmul_rn(d1, d1));
        s += static_cast<double>(__fmul_rn(d2, d2));
        s += static_cast<double>(__fmul_rn(d3, d3));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, off);
    return s;
}

__global__ void __launch_bounds__(kThreads, 1)
nearest_tc_kernel(const __grid_constant__ CUtensorMap tmX, const NParams np) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int px = np.px, nbox = np.nbox, nslots = np.nslots;
    constexpr int bblk = kNQP * 128;
    const uint32_t sB = smem_base + nslots * kSlotBytes;
    uint8_t* Bp = smem + nslots * kSlotBytes;
    uint8_t* tail = Bp + nbox * bblk;
    const uint32_t tail_u32 = sB + nbox * bblk;
    const uint32_t bar_full = tail_u32, bar_empty = tail_u32 + 8 * kMaxSlots;
    const uint32_t bar_accf = tail_u32 + 16 * kMaxSlots, bar_acce = bar_accf + 8 * kAcc;
    const uint32_t bar_nxf = bar_acce + 8 * kAcc, bar_nxe = bar_nxf + 8 * kAcc;
    const uint32_t tmem_slot = bar_nxe + 8 * kAcc;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(tail + (tmem_slot - tail_u32));
    float* qn = reinterpret_cast<float*>(tail + kBar);         // [16] |q|^2 (fp32, bounds only)
    unsigned* thr = reinterpret_cast<unsigned*>(qn + 16);      // [16] smallest upper bound of the squared distance so far (float bits)
    int* qcount = reinterpret_cast<int*>(thr + 16);            // [4]
    float* nxbuf = reinterpret_cast<float*>(qcount + 4);       // [kAcc][128] |x|^2 of the tile's rows
    unsigned* queue = reinterpret_cast<unsigned*>(nxbuf + kAcc * kRows);   // [kQueue]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < nslots; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 5); }   // MMA commit + 4 norm warps
        for (int a = 0; a < kAcc; ++a) {
            mbar_init(bar_accf + 8 * a, 1); mbar_init(bar_acce + 8 * a, 4);
            mbar_init(bar_nxf + 8 * a, 4); mbar_init(bar_nxe + 8 * a, 4);
        }
        fence_barrier_init();
        prefetch_tmap(&tmX);
    }
    if (warp == 0) tmem_alloc<kAcc * kNQP>(tmem_slot);
    for (int i = tid; i < nbox * bblk / 16; i += kThreads) reinterpret_cast<uint4*>(Bp)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < kQueue; i += kThreads) queue[i] = 0xFFFFFFFFu;
    if (tid < 16) { thr[tid] = 0x7f800000u; qn[tid] = 0.0f; }
    if (tid < 4) qcount[tid] = 0;
    __syncthreads();
    {
        const int p4 = px >> 2;
        for (int g = tid; g < np.nq * p4; g += kThreads) {
            const int j = g / p4, c = (g - j * p4) * 4;
            const float4 v = __ldg(reinterpret_cast<const float4*>(np.q + static_cast<long long>(np.q0 + j) * px) + (c >> 2));
            *reinterpret_cast<uint4*>(Bp + (c >> 5) * bblk + tfs::swz_off(j, c & 31)) =
                make_uint4(tfs::tf32_rne(v.x), tfs::tf32_rne(v.y), tfs::tf32_rne(v.z), tfs::tf32_rne(v.w));
        }
        if (warp < np.nq) {                                    // |q_j|^2, warp j
            const float* qr = np.q + static_cast<long long>(np.q0 + warp) * px;
            float s = 0.0f;
            for (int c = lane; c < px; c += 32) { const float v = __ldg(qr + c); s = __fmaf_rn(v, v, s); }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) qn[warp] = s;
        }
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const long long first = blockIdx.x, step = gridDim.x;
    const float cb = tfs::tfs_eps(px) + 2.0e-4f;

    if (warp == 0) {
        // ================================================================ TMA producer
        if (elect_one_sync()) {
            int slot = 0;
            uint32_t ph = 0;
            for (long long tile = first; tile < np.n_tiles; tile += step) {
                const int row0 = static_cast<int>(tile * kRows);
                for (int b = 0; b < nbox; ++b) {
                    tfs::mbar_wait_q(bar_empty + 8 * slot, ph ^ 1u, np.err_flag, 501);
                    mbar_expect_tx(bar_full + 8 * slot, kSlotBytes);
                    tma_load_2d(smem_base + slot * kSlotBytes, &tmX, bar_full + 8 * slot, b * kBoxCols, row0);
                    if (++slot == nslots) { slot = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        if (elect_one_sync()) {
            constexpr uint32_t idesc = tfs::make_idesc_tf32<kNQP>();
            int slot = 0;
            uint32_t ph = 0;
            int it = 0;
            for (long long tile = first; tile < np.n_tiles; tile += step, ++it) {
                const int a = it & (kAcc - 1);
                tfs::mbar_wait_q(bar_acce + 8 * a, ((static_cast<uint32_t>(it) >> 2) & 1u) ^ 1u, np.err_flag, 502);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + a * kNQP;
                for (int b = 0; b < nbox; ++b) {
                    tfs::mbar_wait_q(bar_full + 8 * slot, ph, np.err_flag, 503);
                    tcgen05_fence_after();
                    const int steps = min(4, (px - b * kBoxCols + 7) >> 3);
                    const uint64_t ad = make_smem_desc(smem_base + slot * kSlotBytes), bd = make_smem_desc(sB + b * bblk);
                    for (int k = 0; k < steps; ++k) tfs::umma_tf32(tmem_d, ad + 2u * k, bd + 2u * k, idesc, (b | k) ? 1u : 0u);
                    umma_commit(bar_empty + 8 * slot);
                    if (++slot == nslots) { slot = 0; ph ^= 1u; }
                }
                umma_commit(bar_accf + 8 * a);
            }
        }
        __syncwarp();
    } else if (warp < 6) {
        // ================================================================ row norms: thread = row
        const int r = (warp - 2) * 32 + lane, rs = r & 7;
        int slot = 0;
        uint32_t ph = 0;
        int it = 0;
        for (long long tile = first; tile < np.n_tiles; tile += step, ++it) {
            const int a = it & (kAcc - 1);
            float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
            for (int b = 0; b < nbox; ++b) {
                tfs::mbar_wait_q(bar_full + 8 * slot, ph, np.err_flag, 504);
                const uint8_t* xrow = smem + slot * kSlotBytes + r * 128;
#pragma unroll
                for (int pc = 0; pc < 8; ++pc) {
                    const float4 v = *reinterpret_cast<const float4*>(xrow + ((pc ^ rs) << 4));
                    s0 = __fmaf_rn(v.x, v.x, s0); s1 = __fmaf_rn(v.y, v.y, s1); s2 = __fmaf_rn(v.z, v.z, s2); s3 = __fmaf_rn(v.w, v.w, s3);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8 * slot);
                if (++slot == nslots) { slot = 0; ph ^= 1u; }
            }
            tfs::mbar_wait_q(bar_nxe + 8 * a, ((static_cast<uint32_t>(it) >> 2) & 1u) ^ 1u, np.err_flag, 505);
            nxbuf[a * kRows + r] = (s0 + s1) + (s2 + s3);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_nxf + 8 * a);
        }
    } else {
        // ================================================================ filter + canonical evaluation: thread = row
        const int q4 = warp & 3, ew = warp - 6;
        const int rit = q4 * 32 + lane, gtid = ew * 32 + lane;
        double bd = 0.0;                                       // lane v < nq: this warp's best (distance, row) of query v
        long long bi = -1;
        unsigned long long n_eval = 0;
        auto process_queue = [&]() {
            const int n = min(*reinterpret_cast<volatile int*>(qcount), kQueue);
            for (int idx = ew; idx < n; idx += 4) {
                const unsigned e = queue[idx];
                if (e == 0xFFFFFFFFu) continue;                // (warp-uniform) a reservation that did not fit
                const int j = e & 31;
                const long long row = (first + static_cast<long long>(e >> 12) * step) * kRows + ((e >> 5) & 127);
                const double t = canonical_sq(np.set + row * px, np.q + static_cast<long long>(np.q0 + j) * px, px, lane);
                const double dist = __dsqrt_rn(t);
                ++n_eval;
                if (lane == j && !(dist != dist) && (bi < 0 || dist < bd || (dist == bd && row < bi))) { bd = dist; bi = row; }   // NaN never wins here
                if (lane == 0) {
                    if (t == t) atomicMin(thr + j, __float_as_uint(fmaxf(__double2float_ru(t), 0.0f)));
                    if (row == 0) np.row0_nan[np.q0 + j] = (dist != dist) ? 1 : 0;
                }
            }
            named_bar_sync(1, 128);
            for (int idx = gtid; idx < n; idx += 128) queue[idx] = 0xFFFFFFFFu;
            if (gtid == 0) *reinterpret_cast<volatile int*>(qcount) = 0;
            named_bar_sync(1, 128);
        };
        int it = 0;
        for (long long tile = first; tile < np.n_tiles; tile += step, ++it) {
            const int a = it & (kAcc - 1);
            const uint32_t par = (static_cast<uint32_t>(it) >> 2) & 1u;
            const long long row = tile * kRows + rit;
            const bool live = row < np.n_rows;
            if (ew == 0 && (it & 3) == 0 && lane < np.nq) {    // bounds published by the other blocks
                const unsigned g = *reinterpret_cast<volatile unsigned*>(np.gthr + lane);
                atomicMin(thr + lane, g);
                atomicMin(np.gthr + lane, thr[lane]);
            }
            tfs::warp_mbar_wait(bar_accf + 8 * a, par, lane, np.err_flag, 506);
            tfs::warp_mbar_wait(bar_nxf + 8 * a, par, lane, np.err_flag, 507);
            tcgen05_fence_after();
            uint32_t r0[32];
            tmem_ld16(tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + a * kNQP, r0);
            tmem_ld_wait();
            const float nx = nxbuf[a * kRows + rit];
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_acce + 8 * a); mbar_arrive(bar_nxe + 8 * a); }
            // bounds of the squared distance of this row to every query of the pass
            float lb[NL2_QB];
#pragma unroll
            for (int j = 0; j < NL2_QB; ++j) {
                const float s = nx + qn[j], dt = __uint_as_float(r0[j]);
                lb[j] = __fmaf_rn(-2.0f, dt, (1.0f - cb) * s);
                float ub = __fmaf_rn(-2.0f, dt, (1.0f + cb) * s);
                ub = (live && ub == ub) ? fmaxf(ub, 0.0f) * 1.000001f : __uint_as_float(0x7f800000u);
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) ub = fminf(ub, __shfl_xor_sync(0xffffffffu, ub, o));
                if (lane == 0 && j < np.nq) atomicMin(thr + j, __float_as_uint(ub));
            }
            named_bar_sync(1, 128);                            // every warp's upper bounds are in
            unsigned m = 0u;
#pragma unroll
            for (int j = 0; j < NL2_QB; ++j) {
                const float t = __uint_as_float(thr[j]);
                if (j < np.nq && live && (!(lb[j] > t) || row == 0)) m |= 1u << j;
            }
            for (;;) {
                const int cnt = __popc(m);
                int pos = 0;
                if (cnt) pos = atomicAdd(qcount, cnt);
                const bool over = cnt && pos + cnt > kQueue;
                if (cnt && !over) {
                    unsigned e = m;
                    while (e) { const int j = __ffs(e) - 1; e &= e - 1u; queue[pos++] = (static_cast<unsigned>(it) << 12) | (static_cast<unsigned>(rit) << 5) | j; }
                    m = 0u;
                }
                if (!tfs::named_bar_or(1, 128, over || (cnt && pos >= kTrig))) break;
                process_queue();
            }
        }
        named_bar_sync(1, 128);
        if (*reinterpret_cast<volatile int*>(qcount) > 0) process_queue();
        if (lane < NL2_QB) {
            NearestRec rec;
            rec.d = bd; rec.id = bi;
            np.partial[(static_cast<long long>(blockIdx.x) * 4 + ew) * NL2_QB + lane] = rec;
        }
        if (np.stats) {
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) n_eval = max(n_eval, __shfl_xor_sync(0xffffffffu, n_eval, o));
            if (lane == 0 && n_eval) atomicAdd(np.stats, n_eval);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<kAcc * kNQP>(tmem_base); }
}

inline size_t nearest_fixed_bytes(int px) {
    const int nbox = (px + kBoxCols - 1) / kBoxCols;
    return static_cast<size_t>(nbox) * kNQP * 128 + kBar + 16 * 4 + 16 * 4 + 16 + kAcc * kRows * 4 + kQueue * 4 + 64;
}

}  // namespace ntc
}  // namespace ganrev
