// search_tc.cuh -- many-query cosine search (apply_r.lua:265-282 at 4096 needles) on the 5th-gen tensor cores,
// with EXACT results: a split-bf16 tcgen05 GEMM produces approximate cosines with a proven error bound, only the
// (query, row) pairs that can still reach the top-k are re-scored with the canonical sequential fp32 fmaf chain
// (oracle/ganrev_oracle.c orc_search_cosine), and the top-k is selected with the same total-order
// keys as every other search kernel.  Ids and scores are bit-identical to the fmaf-chain kernels in scan.cuh.
//
// 1. PACK (pack_kernel).  x^ = x * sqrt(1/(|x|^2 + 1e-12)) (the factors of nn.CosineDistance's formula), split into
//    bf16 hi = RN(x^), lo = RN(x^ - hi): |x^ - hi - lo| <= 2^-16 |x^|.  Row layout [hi_0 | lo_0 | hi_1 | lo_1 | ...] in
//    64-column slices of d (128 B each, one TMA 128B-swizzle box), queries and database rows alike.
// 2. FILTER (filter_kernel).  S~ = Q^ X^T ~ qh*xh + ql*xh + qh*xl: three tcgen05.mma chains per 64-column slice into one
//    fp32 TMEM accumulator (the dropped ql*xl term is <= 2^-16 |q^||x^|).  M = 128 queries (TMEM lanes), N = 256 database
//    rows (TMEM columns).  The epilogue thread of lane q holds that query's threshold in a register and scans its
//    columns: pairs with S~ >= thr_q are appended to the query's candidate list (global atomics; rare).
//    Error bound used: |S~ - s_exact| <= eps(d) = 2^-13 + d*2^-20, where s_exact is the canonical fmaf-chain score:
//      split residuals + dropped term           <= 3 * 2^-16                       (Cauchy-Schwarz, |q^|,|x^| <= 1)
//      fp32 accumulation in the tensor core     <= (3*ceil(d/16)) * 6 * 2^-23      (per K=16 step, truncating adds of |terms| <= 3)
//      fmaf-chain score vs the true cosine      <= (4d + 12) * 2^-24               (dot product, two norms, sqrt, product)
//    (sum < eps(d) for every d; tests/test_gpu_search_tc.py measures the observed maximum on the device: <= eps/8).
// 3. THRESHOLDS without a sequential dependency: levels of strided samples of the database (stride s_L > ... > s_1 > 1 =
//    s_0, at most 24x apart, the coarsest ~1024 rows).  The exact k-th best score over a SUBSET of the rows is a lower
//    bound of the k-th best over all rows, so a threshold derived from level l+1 (see 4.) can never drop a top-k member of
//    level l; it passes ~ k * s_{l+1}/s_l candidates per query.  The coarsest level keeps every pair (its approximate scores
//    are written densely).  The strided sample is a TMA tensor map with a larger row pitch: no copy.
// 4. SELECT + RESCORE (rescore_kernel, one block per query).  The pairs carry their approximate scores.  With a_k = the k-th
//    largest approximate score among a query's candidates, the exact k-th best is >= a_k - eps, so only candidates with
//    approx >= a_k - 2 eps can be in the top-k: about k of the ~32 k candidates.  Intermediate levels therefore re-score
//    nothing (next threshold = a_k - 2 eps); the final level runs the exact chain (one thread = one sequential fmaf chain)
//    on the survivors and the special rows, and their total-order keys (score desc, NaN last, lowest id) go through the same
//    warp-level sorted-list insertion as merge_kernel into the `partial` list search_finish merges (also across ranks).
// Rows whose norm is not a positive finite number (NaN / inf entries) are "special": packed as zeros and appended to
// every query's candidates at the final level.  Candidate-list overflow (adversarial duplicates), special queries or
// too many special rows raise a flag and the caller re-runs the search with the fmaf-chain kernels: never a wrong
// answer, only a slower one.
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"
#include "scan.cuh"

namespace ganrev {
namespace stc {

using namespace tc;   // PTX wrappers (mbarrier, TMA, tcgen05) of conv_tc.cuh

constexpr int kEpi = 8;                        // epilogue warps: two per TMEM lane quarter (each takes 128 of the 256 columns)
constexpr int kThr = 64 + 32 * kEpi;           // warp 0 = TMA producer, warp 1 = MMA issuer
constexpr int QM = 128;                        // queries per tile  = MMA M = TMEM lanes
constexpr int RN = 256;                        // database rows per tile = MMA N = TMEM columns
constexpr int kSlice = 64;                     // d-columns per slice (one 128-byte swizzle span of hi, one of lo)
constexpr int kAB = QM * 128;                  // bytes of one query block  [128 rows x 64 bf16]
constexpr int kBB = RN * 128;                  // bytes of one row block    [256 rows x 64 bf16]
constexpr int kStage = 2 * kAB + 2 * kBB;      // q_hi, q_lo, x_hi, x_lo of one slice: 96 KB
constexpr int kStages = 2;
constexpr int kWB = 128;                       // candidate pairs staged per epilogue warp before one global reservation
constexpr int kSmem = kStages * kStage + 1024 /*barriers*/ + kEpi * kWB * 16 /*pair staging*/ + 1024 /*alignment*/;
constexpr int kMaxSpecial = 1024;              // special rows handled exactly; more -> the fmaf-chain kernels take over
constexpr int FLAG_OVERFLOW = 1, FLAG_SPECIAL_QUERY = 2, FLAG_SPECIAL_ROWS = 4;

__host__ __device__ __forceinline__ float tc_eps(int d) { return 1.220703125e-4f + static_cast<float>(d) * 9.5367431640625e-7f; }   // 2^-13 + d*2^-20
__host__ __device__ __forceinline__ int packed_cols(int d) { return 2 * kSlice * ((d + kSlice - 1) / kSlice); }

// ------------------------------------------------------------------ 1. pack
// one thread per (row, pair of columns); is_query: a special vector raises FLAG_SPECIAL_QUERY, else it is listed
__global__ void pack_kernel(const float* __restrict__ x, const float* __restrict__ rnorm, long long n, int d, bf16* __restrict__ out,
                            int is_query, unsigned* __restrict__ special_rows, unsigned* __restrict__ special_count, int* __restrict__ flags) {
    const int dp = kSlice * ((d + kSlice - 1) / kSlice);
    const int half = dp >> 1;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= n * half) return;
    const long long row = idx / half;
    const int c = static_cast<int>(idx - row * half) * 2;
    const bool noscale = rnorm == nullptr;                            // kmeans centroids (kmeans_tc.cuh): split as they are
    const float rn = noscale ? 1.0f : rnorm[row];
    const bool special = !(rn > 0.0f) || !(rn < 3.0e38f);            // NaN / zero (|x|^2 overflowed) / inf
    if (special && c == 0) {
        if (is_query) atomicOr(flags, FLAG_SPECIAL_QUERY);
        else {
            const unsigned slot = atomicAdd(special_count, 1u);
            if (slot < kMaxSpecial) special_rows[slot] = static_cast<unsigned>(row);
            else atomicOr(flags, FLAG_SPECIAL_ROWS);
        }
    }
    const float s = special ? 0.0f : __fsqrt_rn(rn);
    float v0 = 0.0f, v1 = 0.0f;
    if (!special) {
        if (c < d) v0 = __fmul_rn(x[row * d + c], s);
        if (c + 1 < d) v1 = __fmul_rn(x[row * d + c + 1], s);
        if (!noscale && !(fabsf(v0) <= 2.0f)) v0 = 0.0f;             // an inf / NaN entry under a finite norm cannot happen; belt and braces
        if (!noscale && !(fabsf(v1) <= 2.0f)) v1 = 0.0f;
    }
    const bf16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
    const bf16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
    const int slice = c / kSlice, cs = c - slice * kSlice;
    bf16* o = out + row * (2 * dp) + slice * (2 * kSlice) + cs;
    __nv_bfloat162 hv; hv.x = h0; hv.y = h1;
    __nv_bfloat162 lv; lv.x = l0; lv.y = l1;
    *reinterpret_cast<__nv_bfloat162*>(o) = hv;
    *reinterpret_cast<__nv_bfloat162*>(o + kSlice) = lv;
}

// ------------------------------------------------------------------ 2. filter
struct FilterParams {
    int nq;                    // valid queries
    long long n_rows;          // rows of this level's strided sample
    int stride;                // global row = sample row * stride
    int d, nslices;
    int q_tiles;
    long long items;           // q_tiles * ceil(n_rows / RN)
    const float* thr;          // [nq] pass iff approx >= thr
    unsigned* cnt;             // [nq] candidates per query (fire-and-forget REDs)
    uint4* pairs;              // [pair_cap] (query, global row, approximate score bits, 0) in arrival order; grouped by query afterwards
    unsigned* total;           // pairs written so far
    unsigned pair_cap;
    int* flags;
    int* err_flag;
    float* dump;               // debug: [nq][n_rows] approximate scores (or nullptr)
};

template <bool DUMP>
__global__ void __launch_bounds__(kThr, 1)
filter_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ FilterParams p) {
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

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // A flag raised by an EARLIER kernel of this search (special query, predicted or actual candidate overflow): the search is
    // going to be answered by the fmaf-chain kernels, do no work.  (Stream order makes the value uniform across the grid.)
    if (!DUMP && *reinterpret_cast<const volatile int*>(p.flags) != 0) return;
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmQ); prefetch_tmap(&tmX); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kEpi); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);                      // two 256-column accumulators
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (elect_one_sync()) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
                const int qt = static_cast<int>(item % p.q_tiles);                    // query tile fastest: neighbouring CTAs share the row tile in L2
                const long long rt = item / p.q_tiles;
                for (int j = 0; j < p.nslices; ++j) {
                    mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 201);
                    const uint32_t sb = smem_base + stage * kStage;
                    mbar_expect_tx(full_bar(stage), kStage);
                    tma_load_2d(sb, &tmQ, full_bar(stage), j * 128, qt * QM);
                    tma_load_2d(sb + kAB, &tmQ, full_bar(stage), j * 128 + 64, qt * QM);
                    tma_load_2d(sb + 2 * kAB, &tmX, full_bar(stage), j * 128, static_cast<int>(rt * RN));
                    tma_load_2d(sb + 2 * kAB + kBB, &tmX, full_bar(stage), j * 128 + 64, static_cast<int>(rt * RN));
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (elect_one_sync()) {
            constexpr uint32_t idesc = make_idesc<RN, QM>();
            const uint64_t desc_base = make_smem_desc(0);
            auto desc_at = [&](uint32_t addr) { return desc_base | static_cast<uint64_t>((addr & 0x3FFFFu) >> 4); };
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (long long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(tempty_bar(acc), ((it >> 1) & 1u) ^ 1u, p.err_flag, 202);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * RN);
                for (int j = 0; j < p.nslices; ++j) {
                    mbar_wait(full_bar(stage), phase, p.err_flag, 203);
                    tcgen05_fence_after();
                    const uint32_t sb = smem_base + stage * kStage;
                    const uint64_t qh = desc_at(sb), ql = desc_at(sb + kAB), xh = desc_at(sb + 2 * kAB), xl = desc_at(sb + 2 * kAB + kBB);
                    const int steps = min(4, (p.d - j * kSlice + 15) >> 4);          // K = 16 steps holding real columns
                    for (int k = 0; k < steps; ++k) umma_bf16(tmem_d, qh + 2u * k, xh + 2u * k, idesc, (j == 0 && k == 0) ? 0u : 1u);
                    for (int k = 0; k < steps; ++k) umma_bf16(tmem_d, ql + 2u * k, xh + 2u * k, idesc, 1u);
                    for (int k = 0; k < steps; ++k) umma_bf16(tmem_d, qh + 2u * k, xl + 2u * k, idesc, 1u);
                    umma_commit(empty_bar(stage));
                    if (j == p.nslices - 1) umma_commit(tfull_bar(acc));
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: lane = query, columns = rows
        // Candidates are rare (~k * 32 per query and level) but a global atomic with a return value costs ~1000 cycles of
        // this warp's time, and all eight warps hand over per item: so pairs go to a per-warp shared-memory buffer
        // (shared atomics), the per-query counts are fire-and-forget REDs, and ONE global reservation per ~64 pairs
        // moves the buffer out.
        const int quarter = warp & 3;               // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;           // columns [128*half, 128*half + 128)
        const int etid = threadIdx.x - 64;
        const int ew = warp - 2;
        const float pinf = __uint_as_float(0x7f800000u);
        uint4* wbuf = reinterpret_cast<uint4*>(smem + kStages * kStage + 1024) + ew * kWB;
        unsigned* wcnt = reinterpret_cast<unsigned*>(smem + kStages * kStage + 512) + ew;
        if (lane == 0) *wcnt = 0u;
        __syncwarp();
        auto flush = [&]() {                          // whole warp
            __syncwarp();
            const unsigned n = min(*wcnt, static_cast<unsigned>(kWB));
            unsigned base = 0;
            if (lane == 0 && n > 0) base = atomicAdd(p.total, n);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (unsigned i = lane; i < n; i += 32) {
                if (base + i < p.pair_cap) p.pairs[base + i] = wbuf[i];
                else atomicOr(p.flags, FLAG_OVERFLOW);
            }
            __syncwarp();
            if (lane == 0) *wcnt = 0u;
            __syncwarp();
        };
        int it = 0;
        for (long long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const int qt = static_cast<int>(item % p.q_tiles);
            const long long rt = item / p.q_tiles;
            const int acc = it & 1;
            const int qidx = qt * QM + quarter * 32 + lane;
            const float thr = qidx < p.nq ? __ldg(p.thr + qidx) : pinf;
            if (etid == 0) mbar_wait(tfull_bar(acc), (it >> 1) & 1u, p.err_flag, 204);   // the only poller
            named_bar_sync(1, 32 * kEpi);
            tcgen05_fence_after();
            const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * RN + half * 128);
            const long long row_base = rt * RN + half * 128;
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                uint32_t ra[32], rb[32];
                tmem_ld32(tq + c2 * 64, ra);
                tmem_ld32(tq + c2 * 64 + 32, rb);
                tmem_ld_wait();
                if (c2 == 1) {                       // all of this warp's columns are in registers: hand the accumulator back
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
                }
                // One bit per passing column; the rare emission path is then a SHORT loop over set bits that needs only the
                // column index (an unrolled 64-way 'if' ladder is ~100 KB of code run by 73% of the chunks: the first version
                // of this kernel spent 46% of its issue slots waiting for instruction fetches).
                uint32_t m0 = 0u, m1 = 0u;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    m0 |= (__uint_as_float(ra[j]) >= thr) ? (1u << j) : 0u;
                    m1 |= (__uint_as_float(rb[j]) >= thr) ? (1u << j) : 0u;
                }
                unsigned long long mm = (static_cast<unsigned long long>(m1) << 32) | m0;
                while (mm != 0ull) {                 // rare: a (query, row) pair that may still reach the top-k
                    const int j = __ffsll(static_cast<long long>(mm)) - 1;
                    mm &= mm - 1ull;
                    const long long r = row_base + c2 * 64 + j;
                    if (r < p.n_rows) {
                        atomicAdd(p.cnt + qidx, 1u);                       // result unused: a RED, no round trip
                        // the approximate score of column j travels with the pair (it selects what gets re-scored exactly):
                        // a 6-level select tree over the 64 registers, ~63 SELs, only on this rare path
                        uint32_t t[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) t[i] = (j & 32) ? rb[i] : ra[i];
#pragma unroll
                        for (int w = 16; w >= 1; w >>= 1)
#pragma unroll
                            for (int i = 0; i < w; ++i) t[i] = (j & w) ? t[i + w] : t[i];
                        const uint4 pr = make_uint4(static_cast<unsigned>(qidx), static_cast<unsigned>(r * p.stride), t[0], 0u);
                        const unsigned slot = atomicAdd(wcnt, 1u);         // shared memory
                        if (slot < static_cast<unsigned>(kWB)) wbuf[slot] = pr;
                        else {                                             // staging full (a burst): straight to the global list
                            const unsigned g = atomicAdd(p.total, 1u);
                            if (g < p.pair_cap) p.pairs[g] = pr; else atomicOr(p.flags, FLAG_OVERFLOW);
                        }
                    }
                }
                if (DUMP) {
                    if (qidx < p.nq) {
#pragma unroll
                        for (int j = 0; j < 64; ++j) {
                            const long long r = row_base + c2 * 64 + j;
                            if (r < p.n_rows) p.dump[static_cast<long long>(qidx) * p.n_rows + r] = __uint_as_float(j < 32 ? ra[j & 31] : rb[j & 31]);
                        }
                    }
                }
            }
            __syncwarp();
            if (*wcnt >= static_cast<unsigned>(kWB / 2)) flush();
        }
        flush();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------ 3b. group the pairs by query
// offsets[q] = exclusive prefix sum of cnt[q] (one block); a query with more candidates than rescore_kernel can hold raises
// the overflow flag.  cursor[] is cleared for scatter_kernel.
__global__ void __launch_bounds__(1024)
offsets_kernel(const unsigned* __restrict__ cnt, int nq, unsigned cap, unsigned* __restrict__ offsets, unsigned* __restrict__ cursor, int* __restrict__ flags) {
    __shared__ unsigned part[1024];
    const int tid = threadIdx.x;
    const int per = (nq + 1023) / 1024;
    unsigned s = 0;
    for (int i = 0; i < per; ++i) {
        const int q = tid * per + i;
        if (q < nq) {
            const unsigned c = cnt[q];
            if (c > cap) atomicOr(flags, FLAG_OVERFLOW);
            s += c;
            cursor[q] = 0u;
        }
    }
    part[tid] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {                  // Hillis-Steele inclusive scan
        const unsigned v = tid >= off ? part[tid - off] : 0u;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    unsigned run = part[tid] - s;
    for (int i = 0; i < per; ++i) {
        const int q = tid * per + i;
        if (q < nq) { offsets[q] = run; run += cnt[q]; }
    }
}
__global__ void scatter_kernel(const uint4* __restrict__ pairs, const unsigned* __restrict__ total, unsigned pair_cap, const unsigned* __restrict__ offsets,
                               unsigned* __restrict__ cursor, unsigned* __restrict__ cand, float* __restrict__ cand_score, const int* __restrict__ flags) {
    if (*reinterpret_cast<const volatile int*>(flags) != 0) return;
    const unsigned n = min(*total, pair_cap);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 pr = pairs[i];
        const unsigned pos = offsets[pr.x] + atomicAdd(cursor + pr.x, 1u);
        if (pos < pair_cap) { cand[pos] = pr.y; cand_score[pos] = __uint_as_float(pr.z); }
    }
}

// ------------------------------------------------------------------ 4. exact re-score + select
struct RescoreParams {
    const float* db;           // [n][d] fp32 rows
    const float* rdb;          // [n]
    long long n;
    int d;
    const float* q;            // [nq][d]
    const float* rq;           // [nq]
    int nq, k;
    // candidates: the whole strided sample of the coarsest level (implicit_stride > 0: cand_score is the dense [nq][n_implicit]
    // matrix filter_kernel<true> wrote), or the per-query lists of the filter with their approximate scores
    int implicit_stride;
    int n_implicit;
    const unsigned* cnt;       // [nq]
    const unsigned* offsets;   // [nq] first candidate of each query in cand
    const unsigned* cand;      // candidates grouped by query (global row ids)
    const float* cand_score;   // their approximate scores
    int cap;                   // most candidates one query may have
    const unsigned* special_rows;    // final level: appended to every query's candidates
    const unsigned* special_count;
    int final_level;
    float eps;
    int* flags;
    float next_ratio;                // rows of the next (finer) level per row of this one
    unsigned long long* keys_out;    // [nq][k] sorted descending: the `partial` list of the final level
    float* thr_out;                  // [nq] next (finer) level's threshold (not final)
};

// k-th largest of n keys held in shared memory (one warp; the sorted-list insertion of merge_kernel); returns the list too
template <int E>
__device__ __forceinline__ unsigned long long warp_topk(const unsigned long long* keys, int n, int k, int lane, unsigned long long (&L)[E]) {
#pragma unroll
    for (int j = 0; j < E; ++j) L[j] = 0ull;
    unsigned long long kth = 0ull;
    for (int base = 0; base < n; base += 32) {
        const unsigned long long c = base + lane < n ? keys[base + lane] : 0ull;
        unsigned hit = __ballot_sync(0xffffffffu, c > kth);
        while (hit) {
            const int t = __ffs(hit) - 1;
            hit &= hit - 1;
            const unsigned long long cc = __shfl_sync(0xffffffffu, c, t);
            if (cc > kth) {                                       // warp-uniform (kth may have risen since the ballot)
                scan::list_insert<E>(L, cc, lane);
                kth = scan::list_kth<E>(L, k);
            }
        }
    }
    return kth;
}

// k-th largest (1-based) of n 32-bit order-preserving keys in shared memory, whole block of 256 threads: MSB radix select,
// four 8-bit passes over a shared histogram.  Returns 0 when n < k.  Every thread gets the result.
__device__ __forceinline__ uint32_t block_kth_largest(const uint32_t* keys, int n, int k, unsigned* hist /*[256]*/, unsigned* state /*[2]*/) {
    const int tid = threadIdx.x;
    if (n < k) return 0u;
    if (tid == 0) { state[0] = 0u; state[1] = static_cast<unsigned>(k); }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        hist[tid] = 0u;
        __syncthreads();
        const uint32_t prefix = state[0];
        const uint32_t himask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (int i = tid; i < n; i += 256) {
            const uint32_t v = keys[i];
            if ((v & himask) == prefix) atomicAdd(&hist[(v >> shift) & 0xFFu], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned r = state[1];
            int bin = 255;
            for (; bin > 0; --bin) {
                if (r <= hist[bin]) break;
                r -= hist[bin];
            }
            state[1] = r;
            state[0] = prefix | (static_cast<uint32_t>(bin) << shift);
        }
        __syncthreads();
    }
    return state[0];
}

// One block per query.  a_k = the k-th largest APPROXIMATE score among the query's candidates (coarsest level: every row of
// the strided sample, scores written densely by filter_kernel<true>; other levels: the filter's candidate list).  Each of those
// k rows has an exact score >= a_k - eps, so the exact k-th best of the level is >= a_k - eps, and a row that reaches the top-k
// of this (or any finer) level has an approximate score >= a_k - 2 eps.  Hence
//   not final: thr_out = a_k - 2 eps, nothing is scored exactly;
//   final    : only the candidates with approx >= a_k - 2 eps (about k, not k * 32) and the special rows get the exact fmaf
//              chain (one thread = one sequential chain, oracle orc_search_cosine); their total-order keys give the answer.
__global__ void __launch_bounds__(256)
rescore_kernel(const RescoreParams p) {
    constexpr int E = 4;                                          // k <= 128
    extern __shared__ __align__(16) uint8_t sm[];
    const int q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (*reinterpret_cast<const volatile int*>(p.flags) != 0) return;   // raised by an earlier kernel: the fmaf-chain kernels will answer
    const bool dense = p.implicit_stride > 0;
    const int n_list = dense ? p.n_implicit : static_cast<int>(min(p.cnt[q], static_cast<unsigned>(p.cap)));
    const int n_sp = p.final_level ? static_cast<int>(min(*p.special_count, static_cast<unsigned>(kMaxSpecial))) : 0;
    const int n_max = max(n_list, 1) + n_sp;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(sm);                  // [n_list] approximate keys, later [n_sel + n_sp] exact keys
    unsigned* sel = reinterpret_cast<unsigned*>(keys + ((n_max + 2) & ~1));            // even key count: everything behind stays 16-byte aligned
    float* qs = reinterpret_cast<float*>(sel + ((n_max + 3) & ~3));
    __shared__ int s_nsel;
    const float ninf = __uint_as_float(0xff800000u);
    // ---- phase A: the k-th largest approximate score (block radix select on order-preserving 32-bit keys)
    const float* ascore = dense ? p.cand_score + static_cast<long long>(q) * p.n_implicit : p.cand_score + p.offsets[q];
    uint32_t* akeys = reinterpret_cast<uint32_t*>(keys);
    __shared__ unsigned s_hist[256];
    __shared__ unsigned s_state[2];
    for (int i = tid; i < n_list; i += 256) akeys[i] = scan::score_key32(ascore[i]);
    if (tid == 0) s_nsel = 0;
    __syncthreads();
    const uint32_t hi = block_kth_largest(akeys, n_list, p.k, s_hist, s_state);
    __syncthreads();
    const float a_k = scan::score_unkey32(hi);
    const bool usable = hi != 0u && fabsf(a_k) < 3.0e38f;         // else fewer than k candidates: no bound
    const float cut = usable ? a_k - 2.0f * p.eps : ninf;
    if (!p.final_level) {
        // The next level's threshold, and a prediction of its candidate count from THIS level's scores: every row within
        // 4*eps of a_k stands for next_ratio rows that will pass the next filter.  Scores packed closer than the filter's
        // resolution (e.g. the recovered vectors of an untrained R: all cosines within 1e-5 of 1) cannot be pruned by any
        // approximation -- say so now, before the expensive levels run, and let the fmaf-chain kernels answer.
        const float band = usable ? a_k - 4.0f * p.eps : ninf;
        int mine = 0;
        for (int i = tid; i < n_list; i += 256) mine += ascore[i] >= band ? 1 : 0;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if (lane == 0 && mine) atomicAdd(&s_nsel, mine);
        __syncthreads();
        if (tid == 0) {
            p.thr_out[q] = cut;
            if (!usable || static_cast<float>(s_nsel) * p.next_ratio > 0.75f * static_cast<float>(p.cap)) atomicOr(p.flags, FLAG_OVERFLOW);
        }
        return;
    }
    // ---- final level: compact the candidates that can still reach the top-k
    const unsigned* crow = p.cand + p.offsets[q];
    for (int i = tid; i < n_list; i += 256)
        if (ascore[i] >= cut) sel[atomicAdd(&s_nsel, 1)] = crow[i];
    for (int i = tid; i < p.d; i += 256) qs[i] = __ldg(p.q + static_cast<long long>(q) * p.d + i);
    __syncthreads();
    const int n_sel = s_nsel;
    // ---- exact scores
    const float rqv = __ldg(p.rq + q);
    const bool vec = (p.d & 3) == 0;
    const int n_c = n_sel + n_sp;
    for (int i = tid; i < n_c; i += 256) {
        const long long row = i >= n_sel ? p.special_rows[i - n_sel] : sel[i];
        const float rx = __ldg(p.rdb + row);
        const bool dup = i < n_sel && (!(rx > 0.0f) || !(rx < 3.0e38f));   // a special row: scored once, through the special list
        const float* xr = p.db + row * p.d;
        float acc = 0.0f;
        if (vec) {
            for (int c = 0; c < p.d; c += 4) {
                const float4 x4 = __ldg(reinterpret_cast<const float4*>(xr + c));
                const float4 q4 = *reinterpret_cast<const float4*>(qs + c);
                acc = __fmaf_rn(q4.x, x4.x, acc);
                acc = __fmaf_rn(q4.y, x4.y, acc);
                acc = __fmaf_rn(q4.z, x4.z, acc);
                acc = __fmaf_rn(q4.w, x4.w, acc);
            }
        } else {
            for (int c = 0; c < p.d; ++c) acc = __fmaf_rn(qs[c], __ldg(xr + c), acc);
        }
        keys[i] = dup ? 0ull : scan::make_key(scan::cos_from(acc, rqv, rx), static_cast<uint32_t>(row));
    }
    __syncthreads();
    if (warp != 0) return;
    unsigned long long L[E];
    warp_topk<E>(keys, n_c, p.k, lane, L);
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int t = lane * E + j;
        if (t < p.k) p.keys_out[static_cast<long long>(q) * p.k + t] = L[j];
    }
}

}  // namespace stc
}  // namespace ganrev
