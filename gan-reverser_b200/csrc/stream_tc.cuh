// stream_tc.cuh -- the HBM-bound database kernels (a handful of needles / k <= 32 centroids) as ONE warp-specialised
// TMA -> tcgen05(tf32) -> filter pipeline:
//   MODE 0  cosine top-k for <= 32 needles                         (apply_r.lua:267-282)
//   MODE 1  kmeans labels + int64 fixed-point centroid sums        (unsup.kmeans at apply_r.lua:198)
//   MODE 2  cosine-min cluster assignment                          (apply_r.lua:206-218)
//
// The exactness contract of scan.cuh stands (every returned score / label is decided by one sequential fp32 fmaf chain
// over d and the reference's comparator); what changes is who pays for the chains.  At k = 20, d = 100 a row carries
// 4000 FLOPs per 400 bytes: exact chains for every (row, centroid) pair need more fp32 issue than an SM has while HBM
// delivers the row (stream_kernel / rtile_kernel: 0.19-0.47 of HBM).  Here the fp32 rows are never touched by CUDA
// cores unless they matter:
//   * warp 0 (one thread) streams the database through a ring of 16 KB slots with TMA: box = [128 rows][32 fp32
//     columns], 128B-swizzled, i.e. already the K-major operand layout of a kind::tf32 UMMA.  Columns past d and rows
//     past N are zero-filled by TMA.  HBM traffic = the rows, once.
//   * warp 1 (one thread) issues tcgen05.mma.kind::tf32 (M = 128 rows, N = NQP columns, K = 8) straight on those
//     slots.  The tensor core uses the top 19 bits of each fp32 value: |x - x'| <= 2^-10 |x| whether it truncates or
//     rounds.  The needles / centroids are rounded to tf32 on chip (nearest-even, exactly representable, so the
//     hardware conversion is the identity): |c - c'| <= 2^-11 |c|.  By Cauchy-Schwarz the approximate dot product is
//     within (2^-10 + 2^-11) |x||c| of the real one; tensor-core accumulation (products of two 11-bit significands
//     are exact in fp32) and the fmaf chain's own rounding add at most d 2^-21 |x||c|:
//         |approximate - chain| <= tfs_eps(d) |x||c|,  tfs_eps(d) = 2^-10 + 2^-11 + 2^-12 (slack) + d 2^-20
//     (ganrev_debug_tfs_stats reports the largest observed ratio; tests/test_gpu_stream_tc.py asserts it stays below 1).
//   * epilogue warps (thread = row = TMEM lane) read the approximate scores from double-buffered TMEM accumulators:
//       MODE 0: a (row, needle) pair is a candidate iff its approximate cosine is not below [the best k-th score any
//               block has published so far] minus the bound; candidates get the exact chain, the 64-bit total-order
//               key and a place in the block's sorted lists (scan.cuh list_insert).  A block's k-th best is a lower
//               bound of the final k-th best, so sharing it through one atomicMax word per needle keeps the total
//               number of chains near k ln(N/k) instead of that per block.  The filter can pass extra pairs, never
//               drop one.
//       MODE 1/2: every centroid whose approximate objective is within twice the bound of the best is a candidate.
//               One candidate: the label is proven.  Several (a few % of the rows): the (row, centroid) pairs of the
//               warp are spread over its lanes, each lane runs ONE exact chain over the fp32 tile still in shared
//               memory, and the row's owner picks the winner with the reference's comparator (value, then lowest
//               index).  Non-finite rows / centroids and overflowing pair lists go to a global list that
//               label_exact_list_kernel (label_tc.cuh) resolves with every chain.  MODE 2 also runs the winner's chain
//               for every row (the cosine is an output).
//   * MODE 1, last 8 warps: counting sort of the tile's rows by label, then thread (g, c4) adds its labels' rows from
//     the shared-memory tile as int64 run sums (one owner per (label, column): no atomics); the slots go back to the
//     producer when the sums are done.  Integer sums are associative: any split gives identical centroids.
// Results are bit-identical to stream_kernel / rtile_kernel / the oracle (tests/test_gpu_stream_tc.py runs both).
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"
#include "scan.cuh"

namespace ganrev {
namespace tfs {

using namespace tc;

constexpr int kRows = 128;                      // rows per tile = MMA M
constexpr int kBoxCols = 32;                    // fp32 columns per TMA box = 128 bytes = one swizzle span
constexpr int kSlotBytes = kRows * 128;         // 16 KB
constexpr int kMaxSlots = 14;
constexpr int kAcc = 4;                         // TMEM accumulator stages: the MMA warp runs up to four tiles ahead of the epilogue
constexpr int kGroupThreads = 128;              // one epilogue group: 4 warps = the 4 TMEM lane quarters
constexpr int kSumThreads = 512;                // 16 warps: halves the sums latency per tile, which (with 2.5 tiles of ring slots) sets the tile period
constexpr int kAmbWarpBuf = 32;                 // rows for the global list staged per epilogue warp
constexpr int kPairCap = 32;                    // (row, centroid) pairs per warp and tile
constexpr int kBarBytes = 512;
constexpr int kLPT = 2;                         // MODE 1: labels per sums thread (host-checked: nq <= kLPT * (kSumThreads / (d/4)))
constexpr int kQCap = 1024;                     // MODE 0: queued (row, needle) candidates per block
constexpr int kQTrig = 192;                     //         a batch of chains runs once this many are waiting
constexpr int kRxDepth = 4;                     // tiles of row norms in flight per epilogue thread

struct TfsParams {
    scan::ScanParams s;
    long long n_tiles;
    int nbox;                                   // ceil(d / 32)
    int nslots;                                 // ring depth; MODE 1/2: >= nbox + 1 (a tile stays resident until its rows were consumed)
    unsigned* amb_rows;                         // MODE 1/2: rows that need every chain
    unsigned* amb_count;
    unsigned* gthr;                             // MODE 0: [32] best published k-th score per needle (score_key32 order), zeroed before the launch
    unsigned long long* stats;                  // [4] pairs resolved in-kernel / candidates, rows listed, max error ratio (float bits), unused
    int* err_flag;
    long long* trace;                           // optional clock64 timeline of CTA 0 (ganrev_debug_trace_arm("tfs")): [8 roles][256 events]
    int cen_global;                             // MODE 1/2: the chains read the centroids from global memory (shared memory is short)
    int dbg;                                    // 1: every row to the global list; 2: skip the sums (timing); 4: measure the error ratio (MODE 2)
};

// bound of |approximate - exact chain| relative to |x||c| (header comment)
__host__ __device__ __forceinline__ float tfs_eps(int d) { return 9.765625e-4f + 4.8828125e-4f + 2.44140625e-4f + static_cast<float>(d) * 9.5367431640625e-7f; }

template <int MODE> __host__ __device__ constexpr int epi_groups() { return MODE == 0 ? 1 : 2; }
template <int MODE> __host__ __device__ constexpr int threads_of() { return 64 + kGroupThreads * epi_groups<MODE>() + (MODE == 1 ? kSumThreads : 0); }

// kind::tf32 instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N.
template <int N> __device__ __forceinline__ constexpr uint32_t make_idesc_tf32() {
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// named barrier with an OR reduction of a predicate over `threads` threads
__device__ __forceinline__ bool named_bar_or(int id, int threads, bool pred) {
    uint32_t out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 q, %3, 0;\n\t"
        "bar.red.or.pred p, %1, %2, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(out) : "r"(id), "r"(threads), "r"(static_cast<uint32_t>(pred)) : "memory");
    return out != 0;
}
// mbarrier wait that SUSPENDS the thread in hardware (try_wait with a suspend-time hint) instead of spinning: a spinning
// try_wait is a shared-memory-pipe instruction every few cycles, and with a dozen waiting roles per SM the polls starved the
// LDS / SHFL / vote traffic of the working warps (measured: every phase of the sums group ran 3-4x slower than its instruction count).
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_q(uint32_t bar, uint32_t parity, int* err_flag, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = clock64();
    while (!mbar_try_wait_hint(bar, parity, 20000u)) {
        if (clock64() - t0 > kSpinLimitCycles) {
            if (err_flag) atomicExch(err_flag, code);
            __threadfence_system();
            __trap();
        }
    }
}
// Every lane waits (one warp-wide try_wait).  NOT "lane 0 polls, the others __syncwarp()": after that divergent loop the hardware kept
// lane 0 and lanes 1-31 as two groups, and every following vote / shuffle took the compiler's BRA.DIV slow path (WARPSYNC.COLLECTIVE),
// ~200 cycles per ballot (measured with the in-kernel timeline: 4800 cycles for a 20-iteration ballot loop).
__device__ __forceinline__ void warp_mbar_wait(uint32_t bar, uint32_t parity, int lane, int* err_flag, int code) {
    (void)lane;
    mbar_wait_q(bar, parity, err_flag, code);
    __syncwarp();
}
// byte offset of the 16-byte piece holding columns [c, c+4) of tile row r inside a 128B-swizzled box
__device__ __forceinline__ uint32_t swz_off(int r, int c_in_box) {
    return static_cast<uint32_t>(r * 128 + ((((c_in_box >> 2) ^ (r & 7)) << 4)));
}
// nearest-even rounding to tf32 (10 explicit mantissa bits); NaN / inf pass through
__device__ __forceinline__ uint32_t tf32_rne(float v) {
    const uint32_t u = __float_as_uint(v);
    if ((u & 0x7f800000u) == 0x7f800000u) return u;
    return (u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
}
__host__ __device__ __forceinline__ int cen_stride(int d) { return 4 * ((d >> 2) | 1); }   // odd number of 16-byte pieces: rows of different centroids fall into different banks

// exact sequential fmaf chain of tile row r (fp32, in the slot ring starting at slot0) against centroid row cr (16-byte aligned).
// Full boxes run as two unrolled halves: eight 16-byte loads in flight, then 16 dependent FMAs (the chain's latency is the floor).
__device__ __forceinline__ float chain_smem(const uint8_t* smem, int slot0, int nslots, int d, int r, const float* cr) {
    float acc = 0.0f;
    int sl = slot0;
    const int rs = r & 7;
    const int nfull = d >> 5, tailp = (d & 31) >> 2;
    const float4* c4 = reinterpret_cast<const float4*>(cr);
    for (int b = 0; b < nfull; ++b) {
        const uint8_t* xrow = smem + sl * kSlotBytes + r * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float4 xv[4], cv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                xv[e] = *reinterpret_cast<const float4*>(xrow + (((h * 4 + e) ^ rs) << 4));
                cv[e] = c4[h * 4 + e];
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                acc = __fmaf_rn(cv[e].x, xv[e].x, acc);
                acc = __fmaf_rn(cv[e].y, xv[e].y, acc);
                acc = __fmaf_rn(cv[e].z, xv[e].z, acc);
                acc = __fmaf_rn(cv[e].w, xv[e].w, acc);
            }
        }
        c4 += 8;
        if (++sl == nslots) sl = 0;
    }
    if (tailp) {
        const uint8_t* xrow = smem + sl * kSlotBytes + r * 128;
        for (int pc = 0; pc < tailp; ++pc) {
            const float4 xv = *reinterpret_cast<const float4*>(xrow + ((pc ^ rs) << 4));
            const float4 cv = c4[pc];
            acc = __fmaf_rn(cv.x, xv.x, acc);
            acc = __fmaf_rn(cv.y, xv.y, acc);
            acc = __fmaf_rn(cv.z, xv.z, acc);
            acc = __fmaf_rn(cv.w, xv.w, acc);
        }
    }
    return acc;
}
// the same chain with the row in global memory (MODE 0 candidates: the ring slot is long gone)
__device__ __forceinline__ float chain_gmem(const float* __restrict__ xr, int d, const float* cr) {
    float acc = 0.0f;
    const float4* x4 = reinterpret_cast<const float4*>(xr);
    const float4* c4 = reinterpret_cast<const float4*>(cr);
    const int n4 = d >> 2;
    int i = 0;
    for (; i + 4 <= n4; i += 4) {
        float4 xv[4], cv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { xv[e] = __ldg(x4 + i + e); cv[e] = c4[i + e]; }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            acc = __fmaf_rn(cv[e].x, xv[e].x, acc);
            acc = __fmaf_rn(cv[e].y, xv[e].y, acc);
            acc = __fmaf_rn(cv[e].z, xv[e].z, acc);
            acc = __fmaf_rn(cv[e].w, xv[e].w, acc);
        }
    }
    for (; i < n4; ++i) {
        const float4 xv = __ldg(x4 + i), cv = c4[i];
        acc = __fmaf_rn(cv.x, xv.x, acc);
        acc = __fmaf_rn(cv.y, xv.y, acc);
        acc = __fmaf_rn(cv.z, xv.z, acc);
        acc = __fmaf_rn(cv.w, xv.w, acc);
    }
    return acc;
}
// all-ones if a >= b or unordered / if a <= b (ordered): mask-building without predicate juggling
__device__ __forceinline__ unsigned fset_geu(float a, float b) { unsigned r; asm("set.geu.u32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned fset_le(float a, float b) { unsigned r; asm("set.le.u32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned fset_ge(float a, float b) { unsigned r; asm("set.ge.u32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float rsqrt_approx(float a) { float r; asm("rsqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }

#ifdef TFS_EXP_NOCONV
#define TFS_FIX(x, fx) static_cast<long long>(__float_as_int(x))
#else
#define TFS_FIX(x, fx) scan::fix64(x, fx)
#endif
#define TFS_TR(role, e) do { if (tp.trace != nullptr && blockIdx.x == 0 && (e) < 256) tp.trace[(role) * 256 + (e)] = clock64(); } while (0)

template <int MODE, int NQP, int E>
__global__ void __launch_bounds__(threads_of<MODE>(), 1)
tfs_kernel(const __grid_constant__ CUtensorMap tmX, const TfsParams tp) {
    static_assert(kAcc == 4, "stage index and parity arithmetic");
    static_assert(NQP == 16 || NQP == 32, "needle / centroid columns (UMMA N: multiples of 16 at M = 128)");
    constexpr int N = NQP;
    constexpr int K2 = 32 * E;                                 // MODE 0 list length
    constexpr int kTmemCols = kAcc * N < 32 ? 32 : kAcc * N;
    constexpr int G = epi_groups<MODE>();
    const scan::ScanParams& p = tp.s;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int d = p.d, nbox = tp.nbox, nslots = tp.nslots;
    constexpr int bblk = N * 128;                              // one box of the B operand
    const uint32_t sB = smem_base + nslots * kSlotBytes;
    uint8_t* Bp = smem + nslots * kSlotBytes;
    uint8_t* tail = Bp + nbox * bblk;
    const uint32_t tail_u32 = sB + nbox * bblk;
    // barriers
    const uint32_t bar_full = tail_u32, bar_empty = tail_u32 + 8 * kMaxSlots;
    const uint32_t bar_accf = tail_u32 + 16 * kMaxSlots, bar_acce = bar_accf + 8 * kAcc;
    const uint32_t bar_labf = bar_acce + 8 * kAcc, bar_labe = bar_labf + 16;
    const uint32_t tmem_slot = bar_labe + 16;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(tail + (tmem_slot - tail_u32));
    float* colA = reinterpret_cast<float*>(tail + kBarBytes);  // [32] 0.5|c|^2, +inf past nq (MODE 1) / sqrt(rq), 0 past nq (MODE 0, 2)
    float* colB = colA + 32;                                   // [32] rq (MODE 0, 2); [0] = max |c|^2 (MODE 1)
    float* colP = colB + 32;                                   // [32] MODE 2: 0 for a real centroid, +inf for a padded column
    float* tqs = colP + 32;                                    // [32] MODE 0: the needle's threshold in dot-product space (x |x|), NaN = none yet
    int* flagsS = reinterpret_cast<int*>(tqs + 32);            // [4]: [0] a needle / centroid whose norm is not in [1e-15, 1e15]
    float* rxring = reinterpret_cast<float*>(flagsS + 4);      // [G][kRxDepth][128]
    uint8_t* mode_mem = reinterpret_cast<uint8_t*>(rxring + G * kRxDepth * kRows);
    const int cstride = tp.cen_global ? d : cen_stride(d);
    // MODE 0
    unsigned long long* lists = reinterpret_cast<unsigned long long*>(mode_mem);           // [NQP][K2]
    unsigned long long* cand = lists + NQP * K2;                                            // [NQP][CAP]
    unsigned long long* tau = cand + NQP * scan::CAP;                                       // [NQP]
    int* ccount = reinterpret_cast<int*>(tau + NQP);                                        // [NQP]
    int* qn = ccount + NQP;                                                                 // [4] queued candidates
    unsigned* queue = reinterpret_cast<unsigned*>(qn + 4);                                  // [kQCap] (tile of this block << 12 | row in tile << 5 | needle)
    float* cen0 = reinterpret_cast<float*>(queue + kQCap);                                  // [NQP][cstride] fp32 needles for the chains
    // MODE 1 / 2
    unsigned* ambw = reinterpret_cast<unsigned*>(mode_mem);    // [8 warps][kAmbWarpBuf] rows for the global list
    unsigned* pairs = ambw + 8 * kAmbWarpBuf;                  // [8 warps][kPairCap] (owner lane | centroid << 5)
    float* pvals = reinterpret_cast<float*>(pairs + 8 * kPairCap);   // [8 warps][kPairCap] exact objective of the pair
    float* cen12 = pvals + 8 * kPairCap;                       // [NQP][cstride] fp32 centroids for the chains (or in global memory)
    int* bcnt = reinterpret_cast<int*>(cen12 + (tp.cen_global ? 0 : ((p.nq + 3) & ~3) * cstride));   // MODE 1: [2][32] rows per label of the tile
    uint16_t* bucket = reinterpret_cast<uint16_t*>(bcnt + 64); // [2][32][128] the tile's rows grouped by label, as byte offsets r*128 | (r&7)<<4 into a box
    float* cen = MODE == 0 ? cen0 : cen12;
    const float* cenp = tp.cen_global ? p.q : cen;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int nthreads = threads_of<MODE>();
    if (tid == 0) {
        for (int s = 0; s < nslots; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int a = 0; a < kAcc; ++a) { mbar_init(bar_accf + 8 * a, 1); mbar_init(bar_acce + 8 * a, 4); }
        for (int t = 0; t < 2; ++t) { mbar_init(bar_labf + 8 * t, 4); mbar_init(bar_labe + 8 * t, 1); }
        fence_barrier_init();
        flagsS[0] = 0;
        prefetch_tmap(&tmX);
    }
    if (warp == 0) tmem_alloc<kTmemCols>(tmem_slot);
    // ---- needle / centroid operand, K-major 128B-swizzled, rounded to tf32; rows >= nq and columns >= d are zero
    for (int i = tid; i < nbox * bblk / 16; i += nthreads) reinterpret_cast<uint4*>(Bp)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    {
        const int d4 = d >> 2;
        for (int g = tid; g < p.nq * d4; g += nthreads) {
            const int j = g / d4, c = (g - j * d4) * 4;
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.q + static_cast<long long>(j) * d) + (c >> 2));
            *reinterpret_cast<uint4*>(Bp + (c >> 5) * bblk + swz_off(j, c & 31)) = make_uint4(tf32_rne(v.x), tf32_rne(v.y), tf32_rne(v.z), tf32_rne(v.w));
            if (!tp.cen_global) *reinterpret_cast<float4*>(cen + j * cstride + c) = v;
        }
    }
    if (tid < 32) {
        const float inf = __uint_as_float(0x7f800000u);
        if (MODE == 1) {
            const float h = tid < p.nq ? __ldg(p.c2 + tid) : inf;      // padded column: objective -inf
            colA[tid] = h;
            float m = tid < p.nq ? 2.0f * h : 0.0f;
            if (m != m) m = inf;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            colB[tid] = m;
        } else {
            const float r = tid < p.nq ? __ldg(p.rq + tid) : 0.0f;
            colA[tid] = tid < p.nq ? __fsqrt_rn(r) : 0.0f;
            colB[tid] = r;
            colP[tid] = tid < p.nq ? 0.0f : inf;
            tqs[tid] = __uint_as_float(0x7fc00000u);
            if (tid < p.nq && (!(r >= 1.0e-30f) || !(r < 1.0e30f))) atomicOr(&flagsS[0], 1);
        }
    }
    if (MODE == 0) {
        for (int i = tid; i < NQP * K2; i += nthreads) lists[i] = 0ull;
        for (int i = tid; i < NQP; i += nthreads) { tau[i] = 0ull; ccount[i] = 0; }
        for (int i = tid; i < kQCap; i += nthreads) queue[i] = 0xFFFFFFFFu;
        if (tid < 4) qn[tid] = 0;
    }
    if (MODE == 1) {
        if (tid < 64) bcnt[tid] = 0;
    }
    fence_proxy_async_smem();                                  // generic-proxy writes of B -> visible to the tensor core
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const bool qbad = MODE != 1 && flagsS[0] != 0;
    const float eps_d = tfs_eps(d);
    const long long first = blockIdx.x, step = gridDim.x;
    const unsigned valid_mask = p.nq >= 32 ? 0xFFFFFFFFu : ((1u << p.nq) - 1u);

    if (warp == 0) {
        // ================================================================ TMA producer
        if (elect_one_sync()) {
            int slot = 0;
            uint32_t ph = 0;
            for (long long tile = first; tile < tp.n_tiles; tile += step) {
                const int row0 = static_cast<int>(tile * kRows);
                for (int b = 0; b < nbox; ++b) {
                    mbar_wait_q(bar_empty + 8 * slot, ph ^ 1u, tp.err_flag, 401);
                    mbar_expect_tx(bar_full + 8 * slot, kSlotBytes);
                    tma_load_2d(smem_base + slot * kSlotBytes, &tmX, bar_full + 8 * slot, b * kBoxCols, row0);
                    if (++slot == nslots) { slot = 0; ph ^= 1u; }
                }
                TFS_TR(0, static_cast<int>((tile - first) / step));
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        if (elect_one_sync()) {
            constexpr uint32_t idesc = make_idesc_tf32<N>();
            int slot = 0;
            uint32_t ph = 0;
            int it = 0;
            for (long long tile = first; tile < tp.n_tiles; tile += step, ++it) {
                const int a = it & (kAcc - 1);
                mbar_wait_q(bar_acce + 8 * a, ((static_cast<uint32_t>(it) >> 2) & 1u) ^ 1u, tp.err_flag, 402);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + a * N;
                for (int b = 0; b < nbox; ++b) {
                    mbar_wait_q(bar_full + 8 * slot, ph, tp.err_flag, 403);
                    tcgen05_fence_after();
                    const int steps = min(4, (d - b * kBoxCols + 7) >> 3);
                    const uint64_t ad = make_smem_desc(smem_base + slot * kSlotBytes), bd = make_smem_desc(sB + b * bblk);
                    for (int k = 0; k < steps; ++k) umma_tf32(tmem_d, ad + 2u * k, bd + 2u * k, idesc, (b | k) ? 1u : 0u);
                    if (MODE == 0) umma_commit(bar_empty + 8 * slot);     // the filter never reads the fp32 tile again
                    if (++slot == nslots) { slot = 0; ph ^= 1u; }
                }
                umma_commit(bar_accf + 8 * a);
                TFS_TR(1, it);
            }
        }
        __syncwarp();
    } else if (warp < 2 + 4 * G) {
        // ================================================================ epilogue: thread = row; group g takes tiles it = g, g + G, ...
        // Everything here is latency-bound single-warp code (one or two warps per scheduler): instruction count is what matters.
        const int q4 = warp & 3;                               // TMEM lane quarter of this warp
        const int ew = warp - 2;                               // epilogue warp 0 .. 4G-1
        const int grp = ew >> 2;
        const int rit = q4 * 32 + lane;                        // row in tile
        const int gtid = (ew & 3) * 32 + lane;
        const int bar_id = 1 + 2 * grp;                        // named barrier of this group (2 = the sums group)
        unsigned long long n_stat0 = 0, n_stat1 = 0;
        float max_ratio = 0.0f;
        unsigned amb_n = 0;                                    // warp-uniform: staged rows of this warp for the global list
        unsigned* my_amb = ambw + ew * kAmbWarpBuf;
        unsigned* my_pairs = pairs + ew * kPairCap;
        float* my_pvals = pvals + ew * kPairCap;
        auto amb_flush = [&]() {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(tp.amb_count, amb_n);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (unsigned i = lane; i < amb_n; i += 32) tp.amb_rows[base + i] = my_amb[i];
            __syncwarp();
            amb_n = 0;
        };
        int slot0 = 0;                                         // first ring slot of this group's current tile
        {
            int s = grp * nbox;
            while (s >= nslots) s -= nslots;
            slot0 = s;
        }
        unsigned gkey = 0u, gk_next = 0u;                      // MODE 0, warp 0, lane j: best published k-th score key of needle j
        // 1/(|x|^2 + 1e-12) of this thread's row arrives through a private cp.async ring kRxDepth tiles ahead
        float* my_rx = rxring + (grp * kRxDepth) * kRows + rit;
        auto rx_issue = [&](int n) {                           // the group's n-th tile -> ring entry n % kRxDepth
            const long long t = first + (static_cast<long long>(grp) + static_cast<long long>(n) * G) * step;
            const long long r = t * kRows + rit;
            if (t < tp.n_tiles && r < p.n_rows) scan::cp_async4(my_rx + (n & (kRxDepth - 1)) * kRows, p.rdb + r);
            scan::cp_async_commit();                           // (possibly empty) group: keeps the wait counts uniform
        };
#pragma unroll 1
        for (int n = 0; n < kRxDepth; ++n) rx_issue(n);

        // ---- MODE 0: queued candidates -> exact chains (rows from global memory) -> the block's sorted lists
        auto process_queue = [&]() {
            const int nq_ = min(*reinterpret_cast<volatile int*>(qn), kQCap);
            for (int base = 0; base < nq_; base += kGroupThreads) {
                const int idx = base + gtid;
                unsigned e = 0xFFFFFFFFu;
                if (idx < nq_) { e = queue[idx]; queue[idx] = 0xFFFFFFFFu; }
                unsigned long long key = 0ull;
                int j = 0;
                bool pend = false;
                if (e != 0xFFFFFFFFu) {
                    j = e & 31;
                    const long long row = (first + static_cast<long long>(e >> 12) * step) * kRows + ((e >> 5) & 127);
                    const float acc = chain_gmem(p.db + row * d, d, cenp + j * cstride);
                    key = scan::make_key(scan::cos_from(acc, colB[j], __ldg(p.rdb + row)), static_cast<uint32_t>(row));
                    pend = true;
                }
                while (named_bar_or(bar_id, kGroupThreads, pend)) {
                    if (pend) {
                        if (key > tau[j]) {
                            const int sl = atomicAdd(&ccount[j], 1);
                            if (sl < scan::CAP) { cand[j * scan::CAP + sl] = key; pend = false; }
                        } else {
                            pend = false;
                        }
                    }
                    named_bar_sync(bar_id, kGroupThreads);
                    for (int jj = ew; jj < NQP; jj += 4) {
                        const int n = min(ccount[jj], scan::CAP);
                        if (n > 0) {
                            unsigned long long L[E];
#pragma unroll
                            for (int t = 0; t < E; ++t) L[t] = lists[jj * K2 + lane * E + t];
                            for (int t = 0; t < n; ++t) scan::list_insert<E>(L, cand[jj * scan::CAP + t], lane);
#pragma unroll
                            for (int t = 0; t < E; ++t) lists[jj * K2 + lane * E + t] = L[t];
                            const unsigned long long kth = scan::list_kth<E>(L, p.k);
                            __syncwarp();
                            if (lane == 0) {
                                tau[jj] = kth;
                                ccount[jj] = 0;
                                if (kth != 0ull) atomicMax(tp.gthr + jj, static_cast<unsigned>(kth >> 32));   // publish: a lower bound of the final k-th best
                            }
                        }
                    }
                }
            }
            named_bar_sync(bar_id, kGroupThreads);
            if (gtid == 0) *reinterpret_cast<volatile int*>(qn) = 0;
            named_bar_sync(bar_id, kGroupThreads);
        };

        int nloc = 0;
        for (int it = grp; first + static_cast<long long>(it) * step < tp.n_tiles; it += G, ++nloc) {
            const long long tile = first + static_cast<long long>(it) * step;
            const int a = it & (kAcc - 1);                     // accumulator stage
            const int t2 = it & 1;                             // label buffer (MODE 1) = group
            const uint32_t par = (static_cast<uint32_t>(it) >> 2) & 1u, par2 = (static_cast<uint32_t>(it) >> 1) & 1u;
            const long long row = tile * kRows + rit;
            const bool live = row < p.n_rows;
            scan::cp_async_wait<kRxDepth - 1>();
            const float rx = live ? my_rx[(nloc & (kRxDepth - 1)) * kRows] : 1.0f;
            rx_issue(nloc + kRxDepth);
            if (MODE == 0 && ew == 0) {
                // lane j refreshes needle j's threshold: max(this block's k-th best, the best published one) - bound, divided by
                // sqrt(rq_j): a pair passes iff dot >= tqs[j] * |x|
                if ((nloc & 3) == 0) {
                    gkey = max(gkey, gk_next);
                    if (lane < p.nq) gk_next = *reinterpret_cast<volatile unsigned*>(tp.gthr + lane);   // consumed four tiles later
                }
                if (lane < p.nq) {
                    const unsigned long long tl = tau[lane];
                    const unsigned kk = max(tl == 0ull ? 0u : static_cast<unsigned>(tl >> 32), gkey);
                    if (kk != 0u) tqs[lane] = __fdividef(scan::score_unkey32(kk) - (eps_d * 1.001f + 4.0e-6f), colA[lane]);
                }
            }
            warp_mbar_wait(bar_accf + 8 * a, par, lane, tp.err_flag, 404);
            if (gtid == 0) TFS_TR(2 + grp, 2 * nloc);
            tcgen05_fence_after();
            float dot[NQP];
            {
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + a * N;
                uint32_t r0[32];
                if constexpr (NQP == 32) tmem_ld32(taddr, r0); else tmem_ld16(taddr, r0);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < NQP; ++j) dot[j] = __uint_as_float(r0[j]);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acce + 8 * a);      // accumulator stage back to the MMA warp
            if (MODE != 0 && grp == 0 && gtid == 0) TFS_TR(6, 8 * nloc + 0);
            // |x| <= 1e15 (and every |c| <= 1e15, qbad / the margin test): no product or partial sum can overflow in either arithmetic
            const bool rx_ok = (rx >= 1.0e-30f) && (rx < 3.0e38f);

            if constexpr (MODE == 0) {
                // ---- candidate filter: dot_j >= tqs_j * |x|  (NaN thresholds and NaN scores pass)
                const float xw = rsqrt_approx(rx);
                unsigned m = 0u;
#pragma unroll
                for (int q = 0; q < NQP / 4; ++q) {
                    if (4 * q >= p.nq) break;                  // (uniform) padded column groups
                    const float4 t = *reinterpret_cast<const float4*>(tqs + 4 * q);
                    m |= fset_geu(dot[4 * q + 0], t.x * xw) & (1u << (4 * q + 0));
                    m |= fset_geu(dot[4 * q + 1], t.y * xw) & (1u << (4 * q + 1));
                    m |= fset_geu(dot[4 * q + 2], t.z * xw) & (1u << (4 * q + 2));
                    m |= fset_geu(dot[4 * q + 3], t.w * xw) & (1u << (4 * q + 3));
                }
                if (!rx_ok || qbad) m = 0xFFFFFFFFu;
                m &= valid_mask;
                if (!live) m = 0u;
                n_stat0 += __popc(m);
                // queue the pairs; the chains run in batches (a round of chains + list merges per tile would stall the stream)
                for (;;) {
                    const int cnt = __popc(m);
                    int pos = 0;
                    if (cnt) pos = atomicAdd(qn, cnt);
                    const bool over = cnt && pos + cnt > kQCap;
                    if (cnt && !over) {
                        unsigned e = m;
                        while (e) { const int j = __ffs(e) - 1; e &= e - 1u; queue[pos++] = (static_cast<unsigned>(nloc) << 12) | (static_cast<unsigned>(rit) << 5) | j; }
                        m = 0u;
                    }
                    if (!named_bar_or(bar_id, kGroupThreads, over || (cnt && pos >= kQTrig))) break;
                    process_queue();
                }
            } else {
                // ---- candidates of the row: every centroid whose approximate objective is within twice the bound of the best
                float best, lim;
                bool fin;
                unsigned m = 0u;
                float nanacc = 0.0f;                           // becomes NaN iff an approximate score is NaN / inf
                if constexpr (MODE == 1) {
                    // objective c.x - 0.5|c|^2, larger wins; both compared values carry an error of at most eps |x| max|c|
                    // (+ the rounding of the fp32 subtraction)
                    best = __uint_as_float(0xff800000u);
#pragma unroll
                    for (int q = 0; q < NQP / 4; ++q) {
                        if (4 * q >= p.nq) break;              // (uniform) padded column groups
                        const float4 c2v = *reinterpret_cast<const float4*>(colA + 4 * q);
                        const float cc[4] = {c2v.x, c2v.y, c2v.z, c2v.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            nanacc = __fmaf_rn(dot[4 * q + e], 0.0f, nanacc);
                            dot[4 * q + e] = dot[4 * q + e] - cc[e];
                            best = fmaxf(best, dot[4 * q + e]);
                        }
                    }
                    const float cmax2 = colB[0];
                    const float xn = __fsqrt_rn(__fdividef(1.0f, rx)), cm = __fsqrt_rn(cmax2);
                    const float margin = 2.0f * (eps_d * xn * cm * 1.001f + 2.4e-7f * (xn * cm + 0.5f * cmax2)) + 1.0e-30f;
                    fin = (nanacc == 0.0f) && (fabsf(best) < 3.0e38f) && rx_ok && (cmax2 < 1.0e30f) && (margin < 3.0e38f);
                    lim = best - margin;
#pragma unroll
                    for (int q = 0; q < NQP / 4; ++q) {
                        if (4 * q >= p.nq) break;
#pragma unroll
                        for (int e = 0; e < 4; ++e) m |= fset_ge(dot[4 * q + e], lim) & (1u << (4 * q + e));
                    }
                } else {
                    // objective cos = dot sqrt(rx) sqrt(rq), smaller wins; compared in dot * sqrt(rq) space, the bound scaled by |x|
                    best = __uint_as_float(0x7f800000u);
#pragma unroll
                    for (int q = 0; q < NQP / 4; ++q) {
                        if (4 * q >= p.nq) break;              // (uniform) padded column groups
                        const float4 av = *reinterpret_cast<const float4*>(colA + 4 * q);
                        const float4 pv = *reinterpret_cast<const float4*>(colP + 4 * q);
                        const float aa[4] = {av.x, av.y, av.z, av.w}, pp[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            nanacc = __fmaf_rn(dot[4 * q + e], 0.0f, nanacc);
                            dot[4 * q + e] = __fmaf_rn(dot[4 * q + e], aa[e], pp[e]);
                            best = fminf(best, dot[4 * q + e]);
                        }
                    }
                    const float margin = 2.0f * (eps_d * 1.001f + 4.0e-6f) * rsqrt_approx(rx) * 1.0001f;
                    fin = (nanacc == 0.0f) && (fabsf(best) < 3.0e38f) && rx_ok && !qbad;
                    lim = best + margin;
#pragma unroll
                    for (int q = 0; q < NQP / 4; ++q) {
                        if (4 * q >= p.nq) break;
#pragma unroll
                        for (int e = 0; e < 4; ++e) m |= fset_le(dot[4 * q + e], lim) & (1u << (4 * q + e));
                    }
                }
                m &= valid_mask;
                if (!fin || (tp.dbg & 1)) m = 0u;              // (dbg 1: test hook, every row through the global list)
                if (!live) m = 1u;                             // dead rows: nothing to decide
                const int nc = __popc(m);
                if (MODE != 0 && grp == 0 && gtid == 0) TFS_TR(6, 8 * nloc + 1);
                // pairs of this warp: every candidate of a row with several
                const unsigned extra = nc > 1 ? m : 0u;
                const int cnt = nc > 1 ? nc : 0;
                int off = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, off, o); if (lane >= o) off += v; }
                const int total = __shfl_sync(0xffffffffu, off, 31);
                off -= cnt;
                const bool fits = off + cnt <= kPairCap;
                const bool listed = live && (m == 0u || !fits);       // -> global list: every chain, the reference's full comparator
                if (cnt && fits) {
                    unsigned e = extra;
                    int o2 = off;
                    while (e) { const int j = __ffs(e) - 1; e &= e - 1u; my_pairs[o2++] = static_cast<unsigned>(lane) | (static_cast<unsigned>(j) << 5); }
                }
                const int npairs = min(kPairCap, total);       // (ranges that did not fit are not written: stale entries are harmless)
                __syncwarp();
                if (MODE != 0 && grp == 0 && gtid == 0) TFS_TR(6, 8 * nloc + 2);
                scan::Best win;
                win.j = __ffs(m) - 1; win.v = 0.0f;
                if constexpr (MODE == 2) {
                    // the winner's cosine is an output: rows with a single candidate run its chain themselves
                    if (live && nc == 1) {
                        const float acc = chain_smem(smem, slot0, nslots, d, rit, cenp + win.j * cstride);
                        win.v = scan::cos_from(acc, rx, colB[win.j]);
                        if (tp.dbg & 4) {                      // measurement only: observed |approximate - chain| over the bound
                            const float approx = dot[0] * 0.0f + best * __fsqrt_rn(rx);
                            max_ratio = fmaxf(max_ratio, fabsf(approx - win.v) / (eps_d + 1.0e-6f));
                        }
                    }
                }
                for (int base = 0; base < npairs; base += 32) {
                    const int idx = base + lane;
                    const bool has = idx < npairs;
                    const unsigned pr = has ? my_pairs[idx] : 0u;
                    const int ol = pr & 31, j = (pr >> 5) & 31;
                    const float rxo = __shfl_sync(0xffffffffu, rx, ol);
                    if (has) {
                        const float acc = chain_smem(smem, slot0, nslots, d, q4 * 32 + ol, cenp + j * cstride);
                        my_pvals[idx] = MODE == 1 ? __fsub_rn(acc, colA[j]) : scan::cos_from(acc, rxo, colB[j]);
                    }
                }
                __syncwarp();
                if (MODE != 0 && grp == 0 && gtid == 0) TFS_TR(6, 8 * nloc + 3);
                if (cnt && fits) {
                    // owner: the reference's comparator over its candidates, ascending index
                    unsigned e = extra;
                    int o2 = off;
                    bool have = false;
                    while (e) {
                        const int j = __ffs(e) - 1;
                        e &= e - 1u;
                        scan::Best c;
                        c.j = j; c.v = my_pvals[o2++];
                        if (!have) { win = c; have = true; }
                        else if (scan::better<MODE>(c, win)) win = c;
                    }
                    n_stat0 += cnt;
                }
                if (live) {
                    if (listed) {
                        ++n_stat1;
                    } else {
                        p.labels[row] = win.j;
                        if (MODE == 2) p.cosv[row] = win.v;
                    }
                }
                // rows for the global list: staged per warp, flushed in bursts
                const unsigned am = __ballot_sync(0xffffffffu, listed);
                if (am) {
                    if (amb_n + __popc(am) > static_cast<unsigned>(kAmbWarpBuf)) amb_flush();
                    if (listed) my_amb[amb_n + __popc(am & ((1u << lane) - 1u))] = static_cast<unsigned>(row);
                    amb_n += __popc(am);
                    __syncwarp();
                }
                if constexpr (MODE == 2) {
                    named_bar_sync(bar_id, kGroupThreads);     // every row of the tile has been read
                    if (gtid == 0) {
                        int sl = slot0;
                        for (int b = 0; b < nbox; ++b) { mbar_arrive(bar_empty + 8 * sl); if (++sl == nslots) sl = 0; }
                    }
                } else {
                    // the row joins its label's bucket (order inside a bucket is irrelevant: integer sums); the sums group releases the slots
                    if (grp == 0 && gtid == 0) TFS_TR(6, 8 * nloc + 4);
                    warp_mbar_wait(bar_labe + 8 * t2, par2 ^ 1u, lane, tp.err_flag, 405);
                    if (grp == 0 && gtid == 0) TFS_TR(6, 8 * nloc + 5);
                    if (live && !listed) {
                        const int pos = atomicAdd(bcnt + t2 * 32 + win.j, 1);
                        bucket[(t2 * 32 + win.j) * kRows + pos] = static_cast<uint16_t>((rit << 7) | ((rit & 7) << 4));
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_labf + 8 * t2);
                }
                slot0 += G * nbox;
                while (slot0 >= nslots) slot0 -= nslots;
            }
            if (gtid == 0) TFS_TR(2 + grp, 2 * nloc + 1);
        }
        if constexpr (MODE == 0) {
            named_bar_sync(bar_id, kGroupThreads);
            if (*reinterpret_cast<volatile int*>(qn) > 0) process_queue();
            named_bar_sync(bar_id, kGroupThreads);
            for (int i = gtid; i < NQP * p.k; i += kGroupThreads) {
                const int j = i / p.k, t = i - j * p.k;
                if (j < p.nq) p.partial[(static_cast<long long>(blockIdx.x) * p.nq + j) * p.k + t] = lists[j * K2 + t];
            }
        } else {
            if (amb_n) amb_flush();
        }
        if (tp.stats) {
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                n_stat0 += __shfl_xor_sync(0xffffffffu, n_stat0, o);
                n_stat1 += __shfl_xor_sync(0xffffffffu, n_stat1, o);
                max_ratio = fmaxf(max_ratio, __shfl_xor_sync(0xffffffffu, max_ratio, o));
            }
            if (lane == 0) {
                if (n_stat0) atomicAdd(tp.stats + 0, n_stat0);
                if (n_stat1) atomicAdd(tp.stats + 1, n_stat1);
                if (max_ratio > 0.0f) atomicMax(tp.stats + 2, static_cast<unsigned long long>(__float_as_uint(max_ratio)));
            }
        }
    } else {
        // ================================================================ MODE 1: centroid sums of the labelled rows
        if constexpr (MODE == 1) {
            const scan::FixScale fx = scan::make_fix_scale(p.sc);
            const int stid = tid - (64 + kGroupThreads * G);
            const int d4 = d >> 2;
            const int groups = max(1, kSumThreads / d4);
            const int g = stid / d4, c = (stid - g * d4) * 4;
            const int cbox = c >> 5;
            const unsigned cp16 = static_cast<unsigned>(((c & 31) >> 2) << 4);
            // thread (g, c4) owns labels g, g + groups, ... (at most kLPT, host-checked) x 4 columns: the int64 sums live in
            // registers for the whole kernel
            constexpr int LPT = E == 4 ? 2 * kLPT : kLPT;      // (MODE 1 reuses the list-length parameter: E = 4 = wide rows / many labels per thread)
            long long acc[LPT][4];
            unsigned long long cntl[LPT];
#pragma unroll
            for (int u = 0; u < LPT; ++u) { acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0; cntl[u] = 0ull; }
            int slot0 = 0;
            int it = 0;
            for (long long tile = first; tile < tp.n_tiles; tile += step, ++it) {
                const int t2 = it & 1;
                warp_mbar_wait(bar_labf + 8 * t2, (static_cast<uint32_t>(it) >> 1) & 1u, lane, tp.err_flag, 406);
                if (stid == 0) TFS_TR(4, 2 * it);
                if (g < groups && !(tp.dbg & 2)) {
                    int sl = slot0 + cbox;
                    if (sl >= nslots) sl -= nslots;
                    const uint8_t* xcol = smem + sl * kSlotBytes;
#pragma unroll
                    for (int u = 0; u < LPT; ++u) {
                        const int L = g + u * groups;
                        if (L < p.nq) {
                            const int n = bcnt[t2 * 32 + L];
                            const uint16_t* bk = bucket + (t2 * 32 + L) * kRows;
                            int pos = 0;
                            for (; pos + 4 <= n; pos += 4) {
                                const uint2 r4 = *reinterpret_cast<const uint2*>(bk + pos);   // four row offsets
                                float4 x4[4];
                                x4[0] = *reinterpret_cast<const float4*>(xcol + ((r4.x & 0xffffu) ^ cp16));
                                x4[1] = *reinterpret_cast<const float4*>(xcol + ((r4.x >> 16) ^ cp16));
                                x4[2] = *reinterpret_cast<const float4*>(xcol + ((r4.y & 0xffffu) ^ cp16));
                                x4[3] = *reinterpret_cast<const float4*>(xcol + ((r4.y >> 16) ^ cp16));
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    acc[u][0] += TFS_FIX(x4[e].x, fx); acc[u][1] += TFS_FIX(x4[e].y, fx);
                                    acc[u][2] += TFS_FIX(x4[e].z, fx); acc[u][3] += TFS_FIX(x4[e].w, fx);
                                }
                            }
                            for (; pos < n; ++pos) {
                                const float4 x4 = *reinterpret_cast<const float4*>(xcol + (static_cast<unsigned>(bk[pos]) ^ cp16));
                                acc[u][0] += TFS_FIX(x4.x, fx); acc[u][1] += TFS_FIX(x4.y, fx);
                                acc[u][2] += TFS_FIX(x4.z, fx); acc[u][3] += TFS_FIX(x4.w, fx);
                            }
                            cntl[u] += static_cast<unsigned long long>(n);
                        }
                    }
                }
                named_bar_sync(2, kSumThreads);                // the tile's rows and buckets have been read
                if (stid < 32) {
                    bcnt[t2 * 32 + stid] = 0;
                    __syncwarp();
                    if (stid == 0) {
                        int sl = slot0;
                        for (int b = 0; b < nbox; ++b) { mbar_arrive(bar_empty + 8 * sl); if (++sl == nslots) sl = 0; }
                        mbar_arrive(bar_labe + 8 * t2);
                        TFS_TR(4, 2 * it + 1);
                    }
                }
                slot0 += nbox;
                if (slot0 >= nslots) slot0 -= nslots;
            }
            if (g < groups) {
#pragma unroll
                for (int u = 0; u < LPT; ++u) {
                    const int L = g + u * groups;
                    if (L < p.nq) {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (acc[u][e] != 0) atomicAdd(&p.acc[static_cast<long long>(L) * d + c + e], static_cast<unsigned long long>(acc[u][e]));
                        if (c == 0 && cntl[u] != 0ull) atomicAdd(&p.cnt[L], cntl[u]);
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
}

// shared memory of one launch (bytes, before the ring and the 1 KB alignment slack)
inline size_t tfs_fixed_bytes(int mode, int NQP, int K2, int nq, int d, bool cen_global) {
    const int nbox = (d + kBoxCols - 1) / kBoxCols;
    size_t b = static_cast<size_t>(nbox) * NQP * 128 + kBarBytes + 128 * 4 + 16 + 64 + static_cast<size_t>(mode == 0 ? 1 : 2) * kRxDepth * kRows * 4;
    if (mode == 0) return b + static_cast<size_t>(NQP) * (K2 + scan::CAP + 1) * 8 + NQP * 4 + 16 + kQCap * 4 + (cen_global ? 0 : static_cast<size_t>((nq + 3) & ~3) * cen_stride(d) * 4);
    b += 8 * kAmbWarpBuf * 4 + 8 * kPairCap * 8 + (cen_global ? 0 : static_cast<size_t>((nq + 3) & ~3) * cen_stride(d) * 4);
    if (mode == 1) b += 64 * 4 + 2 * 32 * kRows * 2;
    return b;
}

}  // namespace tfs
}  // namespace ganrev
