// kmeans_tc.cuh -- kmeans labelling for MANY centroids (unsup.kmeans at apply_r.lua:198 at BASELINE configs[4]'s k = 1024,
// d = 256) on the tensor cores, bit-exact.
//
// At k = 1024 an iteration is a dense contraction (2*N*k*d FLOPs against 4*N*d bytes, AI ~ k/2 FLOP/B: SURVEY 8d) -- 0.66 PFLOP
// per 1.25M rows, 34 ms of exact fmaf chains.  As in search_tc.cuh the chains are only needed where the decision is close:
//   * operands: the database rows packed by search_tc's pack_kernel (x^ = x * sqrt(1/(|x|^2+1e-12)) as bf16 hi | lo slices,
//     shared with the search path) and the centroids packed the same way WITHOUT scaling; TMA feeds both;
//   * M = 128 rows (TMEM lanes) x N = 256 centroids per MMA, three product chains per 64-column slice; a CTA walks the
//     ceil(k/256) centroid chunks of its row tile, so the epilogue thread of a row keeps the row's best and second-best
//     objective  t_j = x^.c_j - 0.5|c_j|^2 * sqrt(rx)  (= the reference's c.x - 0.5|c|^2 divided by |x|) in registers;
//   * |t_j - exact_j/|x|| <= eps(d) |c_j| (search_tc.cuh's bound with |x^| <= 1), so  best - second > 2 eps(d) max|c|  (+ fp32
//     rounding) proves the label; everything else (about 1 % of the rows, and every row with a NaN / inf anywhere) goes to a list
//     resolved with the sequential fmaf chains and TH's max scan (first NaN wins, first maximum): rows whose THIRD-best is outside
//     the margin need two chains (exact_two_kernel); three-way ties and non-finite rows get every centroid's chain
//     (full_scan_kernel, one warp per 32 centroids, winners combined by an atomic max on an order-preserving key);
//   * update_kernel adds the certain rows to the int64 fixed-point centroid sums (one RED per element; 1024 x 256 addresses, no
//     contention), the exact kernels add the listed rows: integer sums are associative, the split changes nothing.
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"
#include "scan.cuh"
#include "search_tc.cuh"

namespace ganrev {
namespace ktc {

using namespace tc;
using stc::kEpi; using stc::kThr; using stc::QM; using stc::RN; using stc::kAB; using stc::kBB; using stc::kStage; using stc::kStages;

constexpr int kSmem = kStages * kStage + 1024 /*barriers*/ + 2 * RN * 4 /*c2 of the current chunk, double-buffered*/ + 2 * QM * 16 /*half merges*/ + 1024;   // = 2 float4 per row

struct KParams {
    long long n_rows;
    int d, nslices, k, nchunks;
    long long r_tiles;
    const float* rdb;          // [n_rows]
    const float* c2;           // [k] 0.5|c|^2 (the fmaf-chain value the exact kernels subtract)
    const float* cm;           // [2] device: {2 * eps(d) * max|c| * 1.001, max 0.5|c|^2}; NaN / inf centroids make both infinite
    int* labels;               // certain: label; listed: -2
    uint4* amb_rows;           // near-ties between two centroids: (row, best centroid, second centroid, 0)
    unsigned* full_rows;       // rows where every centroid must be scored exactly (three-way ties, NaN / inf)
    unsigned* amb_count;       // [0] near-ties, [1] full rows
    int* err_flag;
};

// the three largest objectives of a row and the centroids of the first two
struct BS { float v1, v2, v3; int j1, j2; };
__device__ __forceinline__ void bs_take(BS& s, float t, int j) {
    if (t > s.v1) { s.v3 = s.v2; s.v2 = s.v1; s.j2 = s.j1; s.v1 = t; s.j1 = j; }
    else if (t > s.v2) { s.v3 = s.v2; s.v2 = t; s.j2 = j; }
    else if (t > s.v3) s.v3 = t;
}

// grid = min(r_tiles, SMs); tmX = packed rows (box 64 x 128), tmC = packed centroids (box 64 x 256)
__global__ void __launch_bounds__(kThr, 1)
label_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmC, const __grid_constant__ KParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + kStages * kStage;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + kStages * kStage + 8 * (2 * kStages + 4));
    float* c2s = reinterpret_cast<float*>(smem + kStages * kStage + 1024);            // [2][256]
    float4* mrg = reinterpret_cast<float4*>(smem + kStages * kStage + 1024 + 2 * RN * 4);   // [128] upper half's (best, second, bj)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmX); prefetch_tmap(&tmC); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kEpi); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (elect_one_sync()) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long rt = blockIdx.x; rt < p.r_tiles; rt += gridDim.x)
                for (int ch = 0; ch < p.nchunks; ++ch)
                    for (int j = 0; j < p.nslices; ++j) {
                        mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 401);
                        const uint32_t sb = smem_base + stage * kStage;
                        mbar_expect_tx(full_bar(stage), kStage);
                        tma_load_2d(sb, &tmX, full_bar(stage), j * 128, static_cast<int>(rt * QM));
                        tma_load_2d(sb + kAB, &tmX, full_bar(stage), j * 128 + 64, static_cast<int>(rt * QM));
                        tma_load_2d(sb + 2 * kAB, &tmC, full_bar(stage), j * 128, ch * RN);
                        tma_load_2d(sb + 2 * kAB + kBB, &tmC, full_bar(stage), j * 128 + 64, ch * RN);
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            constexpr uint32_t idesc = make_idesc<RN, QM>();
            const uint64_t desc_base = make_smem_desc(0);
            auto desc_at = [&](uint32_t addr) { return desc_base | static_cast<uint64_t>((addr & 0x3FFFFu) >> 4); };
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (long long rt = blockIdx.x; rt < p.r_tiles; rt += gridDim.x)
                for (int ch = 0; ch < p.nchunks; ++ch, ++it) {
                    const int acc = it & 1;
                    mbar_wait(tempty_bar(acc), ((it >> 1) & 1u) ^ 1u, p.err_flag, 402);
                    tcgen05_fence_after();
                    const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * RN);
                    for (int j = 0; j < p.nslices; ++j) {
                        mbar_wait(full_bar(stage), phase, p.err_flag, 403);
                        tcgen05_fence_after();
                        const uint32_t sb = smem_base + stage * kStage;
                        const uint64_t xh = desc_at(sb), xl = desc_at(sb + kAB), ch_ = desc_at(sb + 2 * kAB), cl = desc_at(sb + 2 * kAB + kBB);
                        const int steps = min(4, (p.d - j * 64 + 15) >> 4);
                        for (int k = 0; k < steps; ++k) umma_bf16(tmem_d, xh + 2u * k, ch_ + 2u * k, idesc, (j == 0 && k == 0) ? 0u : 1u);
                        for (int k = 0; k < steps; ++k) umma_bf16(tmem_d, xl + 2u * k, ch_ + 2u * k, idesc, 1u);
                        for (int k = 0; k < steps; ++k) umma_bf16(tmem_d, xh + 2u * k, cl + 2u * k, idesc, 1u);
                        umma_commit(empty_bar(stage));
                        if (j == p.nslices - 1) umma_commit(tfull_bar(acc));
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                }
        }
    } else {
        // ------------------------------------------------------------ epilogue: lane = row, columns = centroids
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int etid = threadIdx.x - 64;
        const float ninf = __uint_as_float(0xff800000u);
        const float margin_c = __ldg(p.cm), c2max = __ldg(p.cm + 1);
        int it = 0;
        for (long long rt = blockIdx.x; rt < p.r_tiles; rt += gridDim.x) {
            const long long row = rt * QM + quarter * 32 + lane;
            const bool live = row < p.n_rows;
            const float rx = live ? __ldg(p.rdb + row) : 0.0f;
            const float w = __fsqrt_rn(rx);
            BS s;
            s.v1 = ninf; s.v2 = ninf; s.v3 = ninf; s.j1 = 0; s.j2 = 0;
            for (int ch = 0; ch < p.nchunks; ++ch, ++it) {
                const int acc = it & 1;
                float* c2c = c2s + (it & 1) * RN;
                // this chunk's 0.5|c|^2 (columns past k lose: +inf); the barrier below publishes it
                {
                    const int j = ch * RN + etid;
                    c2c[etid] = j < p.k ? __ldg(p.c2 + j) : __uint_as_float(0x7f800000u);
                }
                if (etid == 0) mbar_wait(tfull_bar(acc), (it >> 1) & 1u, p.err_flag, 404);
                named_bar_sync(1, 32 * kEpi);
                tcgen05_fence_after();
                const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * RN + half * 128);
#pragma unroll
                for (int c2i = 0; c2i < 2; ++c2i) {
                    uint32_t ra[32], rb[32];
                    tmem_ld32(tq + c2i * 64, ra);
                    tmem_ld32(tq + c2i * 64 + 32, rb);
                    tmem_ld_wait();
                    if (c2i == 1) {
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty_bar(acc));
                    }
                    const int col0 = half * 128 + c2i * 64;
#pragma unroll
                    for (int j = 0; j < 32; ++j) bs_take(s, __fmaf_rn(-c2c[col0 + j], w, __uint_as_float(ra[j])), ch * RN + col0 + j);
#pragma unroll
                    for (int j = 0; j < 32; ++j) bs_take(s, __fmaf_rn(-c2c[col0 + 32 + j], w, __uint_as_float(rb[j])), ch * RN + col0 + 32 + j);
                }
            }
            // merge the two column halves of the row, decide
            if (half == 1) {
                mrg[2 * (quarter * 32 + lane)] = make_float4(s.v1, s.v2, s.v3, 0.0f);
                mrg[2 * (quarter * 32 + lane) + 1] = make_float4(__int_as_float(s.j1), __int_as_float(s.j2), 0.0f, 0.0f);
            }
            named_bar_sync(2, 32 * kEpi);
            if (half == 0) {
                const float4 ov = mrg[2 * (quarter * 32 + lane)], oj = mrg[2 * (quarter * 32 + lane) + 1];
                bs_take(s, ov.x, __float_as_int(oj.x));
                bs_take(s, ov.y, __float_as_int(oj.y));
                bs_take(s, ov.z, 0);                              // (can only land in third place or lower)
                const float margin = margin_c + 4.8e-7f * (fabsf(s.v1) + c2max * w);
                const bool sane = (fabsf(s.v1) < 3.0e38f) && (rx > 0.0f) && (rx < 3.0e38f) && (margin < 3.0e38f);
                const bool certain = sane && (s.v1 - s.v2 > margin);
                if (live) {
                    if (certain) p.labels[row] = s.j1;
                    else {
                        // near-tie: only the centroids within the margin of the best can win.  Two of them: the usual case, two exact
                        // chains decide.  A third one, or anything not finite: every centroid is scored exactly.
                        const bool two = sane && (s.v1 - s.v3 > margin);
                        p.labels[row] = -2;
                        if (two) p.amb_rows[atomicAdd(p.amb_count, 1u)] = make_uint4(static_cast<unsigned>(row), static_cast<unsigned>(s.j1), static_cast<unsigned>(s.j2), 0u);
                        else p.full_rows[atomicAdd(p.amb_count + 1, 1u)] = static_cast<unsigned>(row);
                    }
                }
            }
            named_bar_sync(3, 32 * kEpi);                       // mrg is free again
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) { tcgen05_fence_after(); tmem_dealloc<512>(tmem_base); }
}

// max_j 0.5|c_j|^2 -> the margin constants (one warp)
__global__ void cmax_kernel(const float* __restrict__ c2, int k, float eps_d, float* __restrict__ cm) {
    const int lane = threadIdx.x;
    float m = 0.0f;
    for (int j = lane; j < k; j += 32) {
        float v = c2[j];
        if (v != v) v = __uint_as_float(0x7f800000u);
        m = fmaxf(m, v);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) { cm[0] = 2.0f * eps_d * __fsqrt_rn(2.0f * m) * 1.001f; cm[1] = m; }
}

// centroid sums of the certain rows: one warp per row, one RED per element
__global__ void __launch_bounds__(256)
update_kernel(const float* __restrict__ db, long long n_rows, int d, const int* __restrict__ labels, double sc,
              unsigned long long* __restrict__ acc, unsigned long long* __restrict__ cnt) {
    const scan::FixScale fx = scan::make_fix_scale(sc);
    const int lane = threadIdx.x & 31;
    const long long gw = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long row = gw; row < n_rows; row += nw) {
        const int j = __ldg(labels + row);
        if (j < 0) continue;
        const float* xr = db + row * d;
        for (int c = lane; c < d; c += 32)
            atomicAdd(&acc[static_cast<long long>(j) * d + c], static_cast<unsigned long long>(scan::fix64(__ldcs(xr + c), fx)));
        if (lane == 0) atomicAdd(&cnt[j], 1ull);
    }
}

// one sequential fmaf chain over d (index order), 16-byte loads eight columns ahead when the rows allow it
__device__ __forceinline__ float chain_dot(const float* __restrict__ cr, const float* __restrict__ xr, int d) {
    float acc = 0.0f;
    if ((d & 7) == 0 && ((reinterpret_cast<uintptr_t>(cr) | reinterpret_cast<uintptr_t>(xr)) & 15) == 0) {
        const float4* c4 = reinterpret_cast<const float4*>(cr);
        const float4* x4 = reinterpret_cast<const float4*>(xr);
        for (int i = 0; i < (d >> 2); i += 2) {
            const float4 ca = __ldg(c4 + i), cb = __ldg(c4 + i + 1), xa = __ldg(x4 + i), xb = __ldg(x4 + i + 1);
            acc = __fmaf_rn(ca.x, xa.x, acc); acc = __fmaf_rn(ca.y, xa.y, acc); acc = __fmaf_rn(ca.z, xa.z, acc); acc = __fmaf_rn(ca.w, xa.w, acc);
            acc = __fmaf_rn(cb.x, xb.x, acc); acc = __fmaf_rn(cb.y, xb.y, acc); acc = __fmaf_rn(cb.z, xb.z, acc); acc = __fmaf_rn(cb.w, xb.w, acc);
        }
    } else {
        for (int c = 0; c < d; ++c) acc = __fmaf_rn(__ldg(cr + c), __ldg(xr + c), acc);
    }
    return acc;
}

// Near-ties between two centroids: two exact sequential fmaf chains (lanes 0 and 1), TH's comparator, the row's sums.  One warp per row.
__global__ void __launch_bounds__(256)
exact_two_kernel(const scan::ScanParams p, const uint4* __restrict__ rows, const unsigned* __restrict__ count) {
    const scan::FixScale fx = scan::make_fix_scale(p.sc);
    const int lane = threadIdx.x & 31;
    const long long gw = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const unsigned n = *count;
    for (long long i = gw; i < n; i += nw) {
        const uint4 e = rows[i];
        const long long row = e.x;
        const float* xr = p.db + row * p.d;
        scan::Best b;
        b.v = 0.0f; b.j = -1;
        if (lane < 2 && (lane == 0 || e.z != e.y)) {
            const int j = static_cast<int>(lane == 0 ? e.y : e.z);
            const float acc = chain_dot(p.q + static_cast<long long>(j) * p.d, xr, p.d);
            b.j = j; b.v = __fsub_rn(acc, __ldg(p.c2 + j));
        }
        scan::Best ob;
        ob.v = __shfl_xor_sync(0xffffffffu, b.v, 1);
        ob.j = __shfl_xor_sync(0xffffffffu, b.j, 1);
        if (scan::better<1>(ob, b)) b = ob;
        b.v = __shfl_sync(0xffffffffu, b.v, 0);
        b.j = __shfl_sync(0xffffffffu, b.j, 0);
        if (lane == 0) { p.labels[row] = b.j; atomicAdd(&p.cnt[b.j], 1ull); }
        for (int c = lane; c < p.d; c += 32)
            atomicAdd(&p.acc[static_cast<long long>(b.j) * p.d + c], static_cast<unsigned long long>(scan::fix64(__ldg(xr + c), fx)));
    }
}
// Full rows: every centroid's exact chain.  Work item = (row, group of 32 centroids): lane = one centroid, one chain; the group's
// winner goes into the row's 64-bit key with an atomic max.  TH's max scan as a key: NaN highest (first NaN wins), else the value's
// order-preserving bits, lowest index first.
__device__ __forceinline__ unsigned long long th_key(float v, int j) {
    const uint32_t hi = (v != v) ? 0xFFFFFFFFu : min(scan::score_key32(v), 0xFFFFFFFEu);
    return (static_cast<unsigned long long>(hi) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<uint32_t>(j));
}
__global__ void __launch_bounds__(256)
full_scan_kernel(const scan::ScanParams p, const unsigned* __restrict__ rows, const unsigned* __restrict__ count, unsigned long long* __restrict__ keys) {
    const int lane = threadIdx.x & 31;
    const long long gw = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const int ngrp = (p.nq + 31) >> 5;
    const long long items = static_cast<long long>(*count) * ngrp;
    for (long long it = gw; it < items; it += nw) {
        const long long slot = it / ngrp;
        const int j = static_cast<int>(it - slot * ngrp) * 32 + lane;
        const float* xr = p.db + static_cast<long long>(rows[slot]) * p.d;
        unsigned long long key = 0ull;
        if (j < p.nq) {
            const float acc = chain_dot(p.q + static_cast<long long>(j) * p.d, xr, p.d);
            key = th_key(__fsub_rn(acc, __ldg(p.c2 + j)), j);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) key = max(key, __shfl_xor_sync(0xffffffffu, key, o));
        if (lane == 0) atomicMax(keys + slot, key);
    }
}
__global__ void __launch_bounds__(256)
full_finalize_kernel(const scan::ScanParams p, const unsigned* __restrict__ rows, const unsigned* __restrict__ count, unsigned long long* __restrict__ keys) {
    const scan::FixScale fx = scan::make_fix_scale(p.sc);
    const int lane = threadIdx.x & 31;
    const long long gw = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const unsigned n = *count;
    for (long long i = gw; i < n; i += nw) {
        const long long row = rows[i];
        const int j = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(keys[i] & 0xFFFFFFFFull));
        __syncwarp();
        if (lane == 0) { p.labels[row] = j; atomicAdd(&p.cnt[j], 1ull); }
        const float* xr = p.db + row * p.d;
        for (int c = lane; c < p.d; c += 32)
            atomicAdd(&p.acc[static_cast<long long>(j) * p.d + c], static_cast<unsigned long long>(scan::fix64(__ldg(xr + c), fx)));
    }
}

}  // namespace ktc
}  // namespace ganrev
