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
//    s_0, about 32x apart, the coarsest ~1024 rows).  The exact k-th best score tau over a SUBSET of the rows is a lower
//    bound of the k-th best over all rows, so thr_q = tau_q(level l+1) - eps can never drop a top-k member of level l;
//    it passes ~ k * s_{l+1}/s_l candidates per query.  The coarsest level is re-scored exhaustively.  The strided
//    sample is a TMA tensor map with a larger row pitch: no copy.
// 4. RESCORE + SELECT (rescore_kernel, one block per query): every candidate gets the exact score (one thread = one
//    sequential fmaf chain), keys (score desc, NaN last, lowest id) go through the same warp-level sorted-list insertion
//    as merge_kernel.  Output: the level's top-k keys (final level: the `partial` list search_finish merges, also across
//    ranks) and the next level's thresholds.
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
constexpr int kSmem = kStages * kStage + 1024 /*barriers*/ + kEpi * kWB * 8 /*pair staging*/ + 1024 /*alignment*/;
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
    const float rn = rnorm[row];
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
        if (!(fabsf(v0) <= 2.0f)) v0 = 0.0f;                         // an inf / NaN entry under a finite norm cannot happen; belt and braces
        if (!(fabsf(v1) <= 2.0f)) v1 = 0.0f;
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
    uint2* pairs;              // [pair_cap] (query, global row) in arrival order; grouped by query afterwards (group_kernel)
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
        uint2* wbuf = reinterpret_cast<uint2*>(smem + kStages * kStage + 1024) + ew * kWB;
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
                        const uint2 pr = make_uint2(static_cast<unsigned>(qidx), static_cast<unsigned>(r * p.stride));
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
__global__ void scatter_kernel(const uint2* __restrict__ pairs, const unsigned* __restrict__ total, unsigned pair_cap, const unsigned* __restrict__ offsets,
                               unsigned* __restrict__ cursor, unsigned* __restrict__ cand, const int* __restrict__ flags) {
    if (*reinterpret_cast<const volatile int*>(flags) != 0) return;
    const unsigned n = min(*total, pair_cap);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 pr = pairs[i];
        const unsigned pos = offsets[pr.x] + atomicAdd(cursor + pr.x, 1u);
        if (pos < pair_cap) cand[pos] = pr.y;
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
    // candidates: either the implicit strided sample (stride > 0: rows i*stride, i < n_implicit) or the per-query lists
    int implicit_stride;
    int n_implicit;
    const unsigned* cnt;       // [nq]
    const unsigned* offsets;   // [nq] first candidate of each query in cand
    const unsigned* cand;      // candidates grouped by query (global row ids)
    int cap;                   // most candidates one query may have
    const unsigned* special_rows;    // final level: appended to every query's candidates
    const unsigned* special_count;
    int use_special;
    float eps;
    int* flags;
    float next_ratio;                // rows of the next (finer) level per row of this one
    unsigned long long* keys_out;    // [nq][k] sorted descending (the `partial` list of the final level)
    float* thr_out;                  // [nq] next (finer) level's threshold, or nullptr
};

__global__ void __launch_bounds__(256)
rescore_kernel(const RescoreParams p) {
    constexpr int E = 4;                                          // k <= 128
    extern __shared__ __align__(16) uint8_t sm[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(sm);
    const int q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (*reinterpret_cast<const volatile int*>(p.flags) != 0) return;   // raised by an earlier kernel: the fmaf-chain kernels will answer
    const int n_list = p.implicit_stride > 0 ? p.n_implicit : static_cast<int>(min(p.cnt[q], static_cast<unsigned>(p.cap)));
    const int n_sp = p.use_special ? static_cast<int>(min(*p.special_count, static_cast<unsigned>(kMaxSpecial))) : 0;
    const int n_c = n_list + n_sp;
    float* qs = reinterpret_cast<float*>(keys + ((n_c + 1) & ~1));
    for (int i = tid; i < p.d; i += 256) qs[i] = __ldg(p.q + static_cast<long long>(q) * p.d + i);
    __syncthreads();
    const float rqv = __ldg(p.rq + q);
    const bool vec = (p.d & 3) == 0;
    for (int i = tid; i < n_c; i += 256) {
        long long row;
        bool dup = false;
        if (i >= n_list) row = p.special_rows[i - n_list];
        else if (p.implicit_stride > 0) row = static_cast<long long>(i) * p.implicit_stride;
        else row = p.cand[p.offsets[q] + i];
        const float rx = __ldg(p.rdb + row);
        if (p.use_special && i < n_list && (!(rx > 0.0f) || !(rx < 3.0e38f))) dup = true;   // a special row: scored once, through the special list
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
#pragma unroll
    for (int j = 0; j < E; ++j) L[j] = 0ull;
    unsigned long long kth = 0ull;
    for (int base = 0; base < n_c; base += 32) {
        const unsigned long long c = base + lane < n_c ? keys[base + lane] : 0ull;
        unsigned hit = __ballot_sync(0xffffffffu, c > kth);
        while (hit) {
            const int t = __ffs(hit) - 1;
            hit &= hit - 1;
            const unsigned long long cc = __shfl_sync(0xffffffffu, c, t);
            if (cc > kth) {                                       // warp-uniform (kth may have risen since the ballot)
                scan::list_insert<E>(L, cc, lane);
                kth = scan::list_kth<E>(L, p.k);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int t = lane * E + j;
        if (t < p.k) p.keys_out[static_cast<long long>(q) * p.k + t] = L[j];
    }
    if (p.thr_out) {
        const uint32_t hi = static_cast<uint32_t>(kth >> 32);
        const float tau = scan::score_unkey32(hi);
        const bool usable = hi != 0u && fabsf(tau) < 3.0e38f;     // else: list not full, or its k-th entry NaN / infinite
        // Predict the next level's candidate count from THIS level's exact scores: every row within 2*eps of tau stands
        // for next_ratio rows that will pass the filter.  Scores packed closer than the filter's resolution (e.g. recovered
        // vectors of an untrained R: all cosines within 1e-5 of 1) cannot be pruned by any approximation -- say so now,
        // before the expensive levels run, and let the fmaf-chain kernels answer.
        const uint32_t lo_key = usable ? scan::score_key32(tau - 2.0f * p.eps) : 0u;
        int near = 0;
        for (int i = lane; i < n_c; i += 32) near += (static_cast<uint32_t>(keys[i] >> 32) >= lo_key && keys[i] != 0ull) ? 1 : 0;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) near += __shfl_xor_sync(0xffffffffu, near, off);
        if (lane == 0) {
            p.thr_out[q] = usable ? tau - p.eps : __uint_as_float(0xff800000u);
            if (!usable || static_cast<float>(near) * p.next_ratio > 0.5f * static_cast<float>(p.cap)) atomicOr(p.flags, FLAG_OVERFLOW);
        }
    }
}

}  // namespace stc
}  // namespace ganrev
