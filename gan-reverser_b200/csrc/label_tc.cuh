// label_tc.cuh -- kmeans labelling + centroid sums (unsup.kmeans at apply_r.lua:198) and cosine-min assignment
// (apply_r.lua:206-218) for k <= 32 centroids in the HBM-bound regime, with the FMAs taken off the fp32 pipe.
//
// At k = 20, d = 100 a row costs 4000 FLOPs per 400 bytes: 65 TFLOP/s of exact fmaf chains would be needed to follow HBM,
// more than the SM's fp32 pipe delivers (rtile_kernel: 0.19-0.29 of HBM).  The labels, however, only need the chain where
// two centroids are nearly tied:
//   * every thread block streams 128-row tiles of the fp32 database ONCE from HBM (coalesced 16-byte loads, two tiles in flight
//     in registers), keeps an fp32 copy of the tile in shared memory, splits each value into bf16 hi + lo (x = hi + lo to 2^-16) and writes the K-major rows straight
//     into the 128B-swizzled UMMA layout (as conv1_tc.cuh does for R's first conv);
//   * one thread issues tcgen05.mma (M = 128 rows, N = 32 centroids) for xh*ch + xl*ch + xh*cl into a TMEM accumulator;
//   * the epilogue thread of a row reads its 32 approximate scores, forms the mode's objective, and compares the gap between
//     the best and the runner-up with the error bound of the approximation (search_tc.cuh: eps(d) = 2^-13 + d*2^-20 relative
//     to |x||c|).  Gap above the bound: the exact argmax / argmin IS that centroid -- no chain.  Otherwise (and for any
//     NaN / inf) the row goes to a list that label_exact_list_kernel resolves with the sequential fmaf chains and the
//     reference's comparator (first NaN wins / a NaN at j = 0 sticks / lowest index on ties) -- ~1 % of the rows.
//   * MODE 1 then adds the tile's certain rows to the int64 fixed-point centroid sums (counting sort by label, one owner per
//     (label, column), rows from the shared-memory fp32 copy of the tile written when it was split); the listed rows are added by the exact kernel.  Integer sums are associative, so
//     the split changes nothing.  MODE 2 computes the winner's cosine with ONE exact chain per row (the value is an output).
// Results are bit-identical to rtile_kernel / stream_kernel / the oracle.
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"
#include "scan.cuh"
#include "search_tc.cuh"

namespace ganrev {
namespace ltc {

using namespace tc;

constexpr int kThreads = 512;                   // 16 warps: two tiles of prefetch in registers, every phase has warps to hide its latencies
constexpr int LR = 128;                         // rows per tile = MMA M
constexpr int LN = 32;                          // centroid columns = MMA N
constexpr int kABlk = LR * 128;                 // one [128 x 64] bf16 block
constexpr int kBBlk = LN * 128;                 // one [32 x 64] bf16 block
constexpr int kMaxPre = 8;                      // float4 per thread per tile (d <= 128)
constexpr int kAmbStage = 1024;                 // ambiguous rows staged per block before one global reservation

struct LabelParams {
    scan::ScanParams s;          // db, rdb, n_rows, d, q (centroids), rq / c2, nq, labels, cosv, acc, cnt, sc
    long long n_tiles;
    unsigned* amb_rows;          // [n_rows] rows that need the exact chains
    unsigned* amb_count;
    int* err_flag;
    int dbg;                     // timing experiments only (results invalid): 1 skip centroid sums, 2 skip sort + sums, 4 skip epilogue scan, 8 skip MMAs, 16 skip the bf16 build
};

// Database rows: hi = the value TRUNCATED to bf16 (its top 16 bits), lo = (value - hi) truncated to bf16 -- integer / fp32-pipe
// instructions only, because the conversion unit (16 lanes per clock per SM) is this kernel's scarcest pipe (the fixed-point
// centroid sums need one F2I per element).  |value - hi| < 2^-7 |value|, |value - hi - lo| < 2^-14 |value|.
__device__ __forceinline__ uint32_t split_trunc(float a, float b, uint32_t& lo_out) {
    const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
    const float la = a - __uint_as_float(ua & 0xFFFF0000u), lb = b - __uint_as_float(ub & 0xFFFF0000u);   // exact
    lo_out = __byte_perm(__float_as_uint(la), __float_as_uint(lb), 0x7632);
    return __byte_perm(ua, ub, 0x7632);
}
// Centroids (built once per block): round-to-nearest split, |value - hi - lo| <= 2^-16 |value|.
__device__ __forceinline__ uint32_t split_hi_lo(float a, float b, uint32_t& lo_out) {
    const uint32_t hi = pack_bf16x2(a, b);
    const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xFFFF0000u);
    lo_out = pack_bf16x2(a - ah, b - bh);
    return hi;
}
// Error of the approximate score relative to |x||c| with these splits: residuals 2^-14 (x) + 2^-16 (c), dropped xl*cl term
// 2^-7 * 2^-8 = 2^-15, tensor-core accumulation and the fmaf chain's own rounding as in search_tc.cuh:
// < 2^-12 + d*2^-20 for every d.
__host__ __device__ __forceinline__ float label_eps(int d) { return 2.44140625e-4f + static_cast<float>(d) * 9.5367431640625e-7f; }

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
label_tc_kernel(const LabelParams lp) {
    static_assert(MODE == 1 || MODE == 2, "labelling modes");
    const scan::ScanParams& p = lp.s;
    const scan::FixScale fx = scan::make_fix_scale(MODE == 1 ? p.sc : 1.0);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int d = p.d, d4 = d >> 2, nsl = (d + 63) >> 6;
    const unsigned d4magic = 0xFFFFFFFFu / static_cast<unsigned>(d4) + 1u;
    const int a_buf = nsl * 2 * kABlk;                       // one tile: [slice][hi | lo] blocks
    // layout: A (one tile: tile i's MMAs have retired before tile i+1 is built) | B[nsl][2] | xs[2] | tail
    const uint32_t sB = smem_base + a_buf;
    const int S = scan::wide4_stride(d);                      // fp32 tile row stride: S % 32 == 4, 16-byte row reads are conflict-free
    float* xs0 = reinterpret_cast<float*>(smem + a_buf + nsl * 2 * kBBlk);   // [2][128][S] fp32 copies of the tile being labelled and the tile being built
    const int xs_bytes = LR * S * 4;
    uint8_t* tail = smem + a_buf + nsl * 2 * kBBlk + 2 * xs_bytes;
    const uint32_t tail_u32 = smem_base + a_buf + nsl * 2 * kBBlk + 2 * xs_bytes;
    const uint32_t bar = tail_u32;                            // bar, bar + 8: MMAs of the even / odd tiles retired
    const uint32_t tmem_slot = tail_u32 + 16;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(tail + 16);
    float* colA = reinterpret_cast<float*>(tail + 32);        // [32] c2 (MODE 1) / sqrt(rq) (MODE 2) per column; padded columns lose
    float* colB = colA + 32;                                  // [32] rq (MODE 2)
    int* slab = reinterpret_cast<int*>(colB + 32);            // [128] label of each row of the tile, -1 dead, -2 listed
    int* perm = slab + LR;                                    // [128]
    int* lstart = perm + LR;                                  // [33] first sorted position of each label, [nq] = certain rows of the tile
    int* wcnt = lstart + 36;                                  // [32][4] rows per (label, row warp), then their start positions
    int* wrank = wcnt + 128;                                  // [128] rank of a row among its warp's rows of the same label
    unsigned* amb_n = reinterpret_cast<unsigned*>(wrank + LR);     // [1] staged ambiguous rows (+3 pad)
    unsigned* amb_buf = amb_n + 4;                            // [kAmbStage]
    float* cen = reinterpret_cast<float*>(amb_buf + kAmbStage);    // MODE 2: fp32 centroids [32][cstride]
    const int cstride = d | 1;
    unsigned long long* sacc = reinterpret_cast<unsigned long long*>(
        reinterpret_cast<uintptr_t>(cen + (MODE == 2 ? LN * cstride : 0)) + 7 & ~static_cast<uintptr_t>(7));   // MODE 1: [nq*d + nq]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); fence_barrier_init(); *amb_n = 0u; }
    if (warp == 0) tmem_alloc<64>(tmem_slot);                 // two 32-column accumulators
    // ---- centroid operand: rows j < nq split into hi / lo, K-major, 128B-swizzled; rows >= nq are zero
    for (int i = tid; i < nsl * 2 * LN * 8; i += kThreads) reinterpret_cast<uint4*>(smem + a_buf)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    for (int g = tid; g < p.nq * d4; g += kThreads) {
        const int j = g / d4, c = (g - j * d4) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.q + static_cast<long long>(j) * d) + (c >> 2));
        uint32_t lo0, lo1;
        const uint32_t hi0 = split_hi_lo(v.x, v.y, lo0), hi1 = split_hi_lo(v.z, v.w, lo1);
        const int sl = c >> 6, cs = c & 63;                   // slice, column inside the slice
        uint8_t* blk = smem + a_buf + (sl * 2) * kBBlk;
        const int off = j * 128 + ((((cs >> 3) ^ (j & 7)) << 4) | ((cs & 7) << 1));
        *reinterpret_cast<uint2*>(blk + off) = make_uint2(hi0, hi1);
        *reinterpret_cast<uint2*>(blk + kBBlk + off) = make_uint2(lo0, lo1);
        if (MODE == 2) { float* cr = cen + j * cstride + c; cr[0] = v.x; cr[1] = v.y; cr[2] = v.z; cr[3] = v.w; }
    }
    float cmax2 = 0.0f;                                        // max_j |c_j|^2 (MODE 1 margin); NaN / inf poison it on purpose
    __shared__ int s_cbad;                                     // MODE 2: a centroid without a positive finite norm -> every row is listed
    if (tid == 0) s_cbad = 0;
    __syncthreads();
    if (tid < 32) {
        if (MODE == 1) {
            const float h = tid < p.nq ? __ldg(p.c2 + tid) : __uint_as_float(0x7f800000u);   // padded column: score -inf
            colA[tid] = h;
            float m = tid < p.nq ? 2.0f * h : 0.0f;
            if (m != m) m = __uint_as_float(0x7f800000u);
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            colB[0] = m;
        } else {
            const float r = tid < p.nq ? __ldg(p.rq + tid) : 0.0f;
            colA[tid] = tid < p.nq ? __fsqrt_rn(r) : 0.0f;
            colB[tid] = r;
            if (tid < p.nq && (!(r > 0.0f) || !(r < 3.0e38f))) atomicOr(&s_cbad, 1);
        }
    }
    if (MODE == 1)
        for (int i = tid; i < p.nq * d + p.nq; i += kThreads) sacc[i] = 0ull;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    if (MODE == 1) cmax2 = colB[0];
    const bool cbad = MODE == 2 && s_cbad != 0;
    constexpr uint32_t idesc = make_idesc<LN, LR>();
    const float eps_d = label_eps(d);

    // ---- tile streaming: float4 idx = tid + 256*u of the tile's contiguous 128*d floats
    const long long first = blockIdx.x, step = gridDim.x;
    float4 preA[kMaxPre], preB[kMaxPre];                       // two tiles in flight from HBM (one tile per HBM latency is 0.4 of the bandwidth)
    auto prefetch = [&](long long tile, float4 (&pre)[kMaxPre]) {
        const long long row0 = tile * LR;
        const long long live4 = min(static_cast<long long>(LR), p.n_rows - row0) * d4;
        const float4* src = reinterpret_cast<const float4*>(p.db + row0 * d);
#pragma unroll
        for (int u = 0; u < kMaxPre; ++u) {
            const int idx = tid + kThreads * u;
            pre[u] = (idx < LR * d4 && idx < live4) ? __ldg(src + idx) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);   // (not evict-first: the tile is read again from L2)
        }
    };
    // registers -> hi / lo -> swizzled A blocks of buffer b, then one thread issues the tile's MMAs into accumulator b
    auto build_and_issue = [&](int b, const float4 (&pre)[kMaxPre]) {
        uint8_t* A = smem;
        float* xb = xs0 + b * (LR * S);
#pragma unroll
        for (int u = 0; u < kMaxPre; ++u) {
            const int idx = tid + kThreads * u;
            if (idx < LR * d4 && !(lp.dbg & 16)) {
                const int r = d4 == 1 ? idx : static_cast<int>(__umulhi(static_cast<unsigned>(idx), d4magic));
                const int c = (idx - r * d4) * 4;
                uint32_t lo0, lo1;
                const uint32_t hi0 = split_trunc(pre[u].x, pre[u].y, lo0), hi1 = split_trunc(pre[u].z, pre[u].w, lo1);
                const int sl = c >> 6, cs = c & 63;
                uint8_t* blk = A + (sl * 2) * kABlk;
                const int off = r * 128 + ((((cs >> 3) ^ (r & 7)) << 4) | ((cs & 7) << 1));
                *reinterpret_cast<uint2*>(blk + off) = make_uint2(hi0, hi1);
                *reinterpret_cast<uint2*>(blk + kABlk + off) = make_uint2(lo0, lo1);
                *reinterpret_cast<float4*>(xb + r * S + c) = pre[u];   // the exact values stay on chip for the centroid sums / the winner's chain
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        tcgen05_fence_before();
        __syncthreads();
        if (warp == 4 && elect_one_sync()) {
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + b * LN;
            bool first_mma = true;
            for (int sl = 0; sl < ((lp.dbg & 8) ? 0 : nsl); ++sl) {
                const int steps = min(4, (d - sl * 64 + 15) >> 4);
                const uint64_t xh = make_smem_desc(smem_base + (sl * 2) * kABlk), xl = make_smem_desc(smem_base + (sl * 2 + 1) * kABlk);
                const uint64_t ch = make_smem_desc(sB + (sl * 2) * kBBlk), cl = make_smem_desc(sB + (sl * 2 + 1) * kBBlk);
                for (int k = 0; k < steps; ++k) { umma_bf16(tmem_d, xh + 2u * k, ch + 2u * k, idesc, first_mma ? 0u : 1u); first_mma = false; }
                for (int k = 0; k < steps; ++k) umma_bf16(tmem_d, xl + 2u * k, ch + 2u * k, idesc, 1u);
                for (int k = 0; k < steps; ++k) umma_bf16(tmem_d, xh + 2u * k, cl + 2u * k, idesc, 1u);
            }
            umma_commit(bar + 8 * b);
        }
        __syncwarp();
    };
    // Rows of a tile whose zero-padded part of the last A block was written by an earlier, wider use: none (d is fixed per launch);
    // the unused 16-byte chunks of the last slice are never read (steps covers only real columns, rounded up to 16: clear them once).
    for (int i = tid; i < a_buf / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();

    // prologue: tile 0 built from set A; tiles 1 (set B) and 2 (set A) in flight
    if (first < lp.n_tiles) {
        prefetch(first, preA);
        if (first + step < lp.n_tiles) prefetch(first + step, preB);
        build_and_issue(0, preA);
        if (first + 2 * step < lp.n_tiles) prefetch(first + 2 * step, preA);
    }
    int it = 0;
    for (long long tile = first; tile < lp.n_tiles; tile += step, ++it) {
        const int b = it & 1;
        const float* xs = xs0 + b * (LR * S);
        // tile i's MMAs (issued one iteration ago) have retired: its accumulator can be read and the A tile rebuilt
        if (tid == 0) mbar_wait(bar + 8 * b, static_cast<uint32_t>(it >> 1) & 1u, lp.err_flag, 301);
        __syncthreads();                                        // (also: the previous tile's sums / chains are done with xs[b ^ 1], slab, perm)
        tcgen05_fence_after();
        if (tile + step < lp.n_tiles) {
            // tile i+1 sits in set B when i is even, in set A when i is odd; once built, that set takes tile i+3
            if (b == 0) { build_and_issue(1, preB); if (tile + 3 * step < lp.n_tiles) prefetch(tile + 3 * step, preB); }
            else        { build_and_issue(0, preA); if (tile + 3 * step < lp.n_tiles) prefetch(tile + 3 * step, preA); }
        }
        const long long row0 = tile * LR;
        if (warp < 4) {
            // ---- epilogue: lane = row, 32 centroid columns
            const long long row = row0 + warp * 32 + lane;
            const bool live = row < p.n_rows;
            const float rx = live ? __ldg(p.rdb + row) : 0.0f;
            uint32_t r[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + b * LN, r);
            tmem_ld_wait();
            tcgen05_fence_before();
            float best, second;
            int bj = 0;
            bool certain;
            if (lp.dbg & 4) { best = 1.0f; second = 0.0f; bj = lane % p.nq; certain = true; }
            else if (MODE == 1) {
                // objective c.x - 0.5|c|^2, larger wins
                best = __uint_as_float(0xff800000u); second = best;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float t = __uint_as_float(r[j]) - colA[j];
                    if (t > best) { second = best; best = t; bj = j; }
                    else if (t > second) second = t;
                }
                // |x| <= sqrt(1/rx) (rx = 1/(|x|^2 + 1e-12)); both scores carry an error of at most eps_d |x| max|c| (+ the rounding of
                // the fp32 subtraction), and they must differ by more than the sum
                const float xn = __fsqrt_rn(__fdividef(1.0f, rx)), cm = __fsqrt_rn(cmax2);
                const float margin = 2.0f * (eps_d * xn * cm * 1.001f + 2.4e-7f * (xn * cm + 0.5f * cmax2));
                certain = (best - second > margin) && (fabsf(best) < 3.0e38f) && (rx > 0.0f) && (rx < 3.0e38f) && (margin < 3.0e38f);
            } else {
                // objective cos, smaller wins
                best = __uint_as_float(0x7f800000u); second = best;
                const float w = __fsqrt_rn(rx);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float t = j < p.nq ? __uint_as_float(r[j]) * w * colA[j] : __uint_as_float(0x7f800000u);
                    if (t < best) { second = best; best = t; bj = j; }
                    else if (t < second) second = t;
                }
                certain = (second - best > 2.0f * eps_d * 1.001f) && (fabsf(best) < 3.0e38f) && (rx > 0.0f) && (rx < 3.0e38f) && !cbad;
                if (p.nq == 1) certain = (fabsf(best) < 3.0e38f) && (rx > 0.0f) && (rx < 3.0e38f) && !cbad;
            }
            if (MODE == 1 && p.nq == 1) certain = (fabsf(best) < 3.0e38f) && (rx > 0.0f) && (rx < 3.0e38f);
            int lab = -1;
            if (live) {
                if (certain) {
                    lab = bj;
                    p.labels[row] = bj;
                    if (MODE == 2) {
                        // the value is an output: the winner's exact cosine, one sequential fmaf chain
                        const float* xr = xs + (warp * 32 + lane) * S;
                        const float* cr = cen + bj * cstride;
                        float acc = 0.0f;
                        for (int c = 0; c < d; c += 4) {
                            const float4 x4 = *reinterpret_cast<const float4*>(xr + c);
                            acc = __fmaf_rn(cr[c + 0], x4.x, acc);
                            acc = __fmaf_rn(cr[c + 1], x4.y, acc);
                            acc = __fmaf_rn(cr[c + 2], x4.z, acc);
                            acc = __fmaf_rn(cr[c + 3], x4.w, acc);
                        }
                        p.cosv[row] = scan::cos_from(acc, rx, colB[bj]);
                    }
                } else {
                    lab = -2;
                    const unsigned slot = atomicAdd(amb_n, 1u);                     // shared memory
                    if (slot < static_cast<unsigned>(kAmbStage)) amb_buf[slot] = static_cast<unsigned>(row);
                    else lp.amb_rows[atomicAdd(lp.amb_count, 1u)] = static_cast<unsigned>(row);   // burst: straight to the list
                }
            }
            slab[warp * 32 + lane] = lab;
        }
        __syncthreads();
        if (*amb_n >= static_cast<unsigned>(kAmbStage / 2)) {                      // block-uniform
            const unsigned n = min(*amb_n, static_cast<unsigned>(kAmbStage));
            __shared__ unsigned s_base;
            if (tid == 0) s_base = atomicAdd(lp.amb_count, n);
            __syncthreads();
            for (unsigned i = tid; i < n; i += kThreads) lp.amb_rows[s_base + i] = amb_buf[i];
            __syncthreads();
            if (tid == 0) *amb_n = 0u;
        }
        if (MODE == 1 && !(lp.dbg & 2)) {
            // ---- centroid sums of the certain rows: counting sort by label, then thread (g, c4) walks its labels' rows with
            // int64 run sums in registers (rows re-read from L2) and adds them to the block accumulator -- one owner per
            // (label, column), no atomics (rtile_kernel's scheme).
            // rows of warp w (32 rows each) ranked inside their label by ballots; wcount[L][w] -> starts by one warp
            if (warp < 4) {
                const int mine = slab[tid];                      // -1 dead / -2 listed: not ranked
                int myrank = 0;
                for (int L = 0; L < p.nq; ++L) {
                    const unsigned m = __ballot_sync(0xffffffffu, mine == L);
                    if (mine == L) myrank = __popc(m & ((1u << lane) - 1u));
                    if (lane == L) wcnt[L * 4 + warp] = __popc(m);
                }
                wrank[tid] = myrank;
            }
            __syncthreads();
            if (warp == 0) {
                // lane L: rows of label L per warp -> exclusive starts, labels in order; lstart[nq] = all certain rows
                int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                if (lane < p.nq) { c0 = wcnt[lane * 4]; c1 = wcnt[lane * 4 + 1]; c2 = wcnt[lane * 4 + 2]; c3 = wcnt[lane * 4 + 3]; }
                const int tot = c0 + c1 + c2 + c3;
                int incl = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                const int start = incl - tot;
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (lane < p.nq) {
                    wcnt[lane * 4] = start; wcnt[lane * 4 + 1] = start + c0; wcnt[lane * 4 + 2] = start + c0 + c1; wcnt[lane * 4 + 3] = start + c0 + c1 + c2;
                    lstart[lane] = start;
                }
                if (lane == 0) lstart[p.nq] = total;
            }
            __syncthreads();
            if (warp < 4) {
                const int mine = slab[tid];
                if (mine >= 0) perm[wcnt[mine * 4 + warp] + wrank[tid]] = tid;
            }
            __syncthreads();
            const int tpg = d4;
            const int g = tid / tpg, c = (tid - g * tpg) * 4;
            const int groups = max(1, kThreads / tpg);
            if (g < groups && !(lp.dbg & 1)) {
                for (int L = g; L < p.nq; L += groups) {
                    const int pos0 = lstart[L], pos1 = lstart[L + 1];
                    if (pos0 == pos1) continue;
                    long long run0 = 0, run1 = 0, run2 = 0, run3 = 0;
                    int pos = pos0;
                    for (; pos + 4 <= pos1; pos += 4) {
                        float4 x4[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) x4[e] = *reinterpret_cast<const float4*>(xs + perm[pos + e] * S + c);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            run0 += scan::fix64(x4[e].x, fx); run1 += scan::fix64(x4[e].y, fx);
                            run2 += scan::fix64(x4[e].z, fx); run3 += scan::fix64(x4[e].w, fx);
                        }
                    }
                    for (; pos < pos1; ++pos) {
                        const float4 x4 = *reinterpret_cast<const float4*>(xs + perm[pos] * S + c);
                        run0 += scan::fix64(x4.x, fx); run1 += scan::fix64(x4.y, fx);
                        run2 += scan::fix64(x4.z, fx); run3 += scan::fix64(x4.w, fx);
                    }
                    unsigned long long* a = sacc + L * d + c;
                    a[0] += static_cast<unsigned long long>(run0);
                    a[1] += static_cast<unsigned long long>(run1);
                    a[2] += static_cast<unsigned long long>(run2);
                    a[3] += static_cast<unsigned long long>(run3);
                    if (c == 0) sacc[p.nq * d + L] += static_cast<unsigned long long>(pos1 - pos0);
                }
            }
        }
    }
    __syncthreads();
    {   // last flush of the staged ambiguous rows, block sums -> global
        const unsigned n = min(*amb_n, static_cast<unsigned>(kAmbStage));
        __shared__ unsigned s_base2;
        if (tid == 0 && n > 0) s_base2 = atomicAdd(lp.amb_count, n);
        __syncthreads();
        for (unsigned i = tid; i < n; i += kThreads) lp.amb_rows[s_base2 + i] = amb_buf[i];
    }
    if (MODE == 1) {
        const int per = p.nq * d + p.nq;
        for (int i = tid; i < per; i += kThreads) {
            const unsigned long long v = sacc[i];
            if (v != 0ull) {
                if (i < p.nq * d) atomicAdd(&p.acc[i], v);
                else atomicAdd(&p.cnt[i - p.nq * d], v);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<64>(tmem_base); }
}

// The listed rows: exact sequential fmaf chains against every centroid (lane j = centroid j), the reference's comparator, and
// (MODE 1) the row's contribution to the fixed-point sums.  One warp per row.
template <int MODE>
__global__ void __launch_bounds__(256)
label_exact_list_kernel(const scan::ScanParams p, const unsigned* __restrict__ rows, const unsigned* __restrict__ count) {
    const scan::FixScale fx = scan::make_fix_scale(MODE == 1 ? p.sc : 1.0);
    const int lane = threadIdx.x & 31;
    const long long gw = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const unsigned n = *count;
    const bool has = lane < p.nq;
    const float aux = has ? __ldg((MODE == 1 ? p.c2 : p.rq) + lane) : 0.0f;
    const float* cr = p.q + static_cast<long long>(has ? lane : 0) * p.d;
    for (long long i = gw; i < n; i += nw) {
        const long long row = rows[i];
        const float* xr = p.db + row * p.d;
        float acc = 0.0f;
        for (int c = 0; c < p.d; ++c) acc = __fmaf_rn(__ldg(cr + c), __ldg(xr + c), acc);
        scan::Best b;
        b.j = has ? lane : -1;
        b.v = MODE == 1 ? __fsub_rn(acc, aux) : scan::cos_from(acc, __ldg(p.rdb + row), aux);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            scan::Best ob;
            ob.v = __shfl_xor_sync(0xffffffffu, b.v, o);
            ob.j = __shfl_xor_sync(0xffffffffu, b.j, o);
            if (scan::better<MODE>(ob, b)) b = ob;
        }
        if (lane == 0) {
            p.labels[row] = b.j;
            if (MODE == 2) p.cosv[row] = b.v;
            if (MODE == 1) atomicAdd(&p.cnt[b.j], 1ull);
        }
        if (MODE == 1)
            for (int c = lane; c < p.d; c += 32)
                atomicAdd(&p.acc[static_cast<long long>(b.j) * p.d + c], static_cast<unsigned long long>(scan::fix64(__ldg(xr + c), fx)));
    }
}

}  // namespace ltc
}  // namespace ganrev
