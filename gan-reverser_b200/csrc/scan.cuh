// scan.cuh -- database kernels over the recovered noise vectors: cosine top-k search
// (apply_r.lua:265-282), kmeans assignment + centroid sums (unsup.kmeans at apply_r.lua:198),
// cosine-min cluster assignment (apply_r.lua:206-218), per-cluster members + mean image
// (apply_r.lua:222-243).
//
// Exactness contract (mirrored by oracle/ganrev_oracle.c): every (query,row)
// dot product is ONE thread's sequential fp32 fmaf chain over d in index order, norms the
// same, IEEE div/sqrt, no fast-math; selections use a total order (score, lowest id), so the
// result does not depend on tiling, split count or GPU count.  Centroid sums are int64
// fixed-point (associative), so any partition of the rows gives identical centroids.
//
// Tiling: a 256-thread block holds a [16*TQ queries] x [128 rows] score tile in registers
// (TQ x 8 per thread); rows and queries are staged through shared memory in 32-wide d-chunks
// with coalesced global loads.
#pragma once
#include "common.cuh"

namespace ganrev {
namespace scan {

constexpr int kThreads = 256;
constexpr int RT = 128;        // rows per tile
constexpr int DK = 32;         // d-chunk staged in smem
constexpr int XS = RT + 4;     // smem stride of one d-row of the row tile (keeps float4 alignment)
constexpr int CAP = 32;        // candidate buffer entries per query

// ---- total-order keys: larger key = better ("score desc, NaN last, -0 == +0, lowest id")
__device__ __forceinline__ uint32_t score_key32(float s) {
    if (s != s) return 0u;
    s = __fadd_rn(s, 0.0f);   // -0 -> +0
    const uint32_t b = __float_as_uint(s);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float score_unkey32(uint32_t k) {
    if (k == 0u) return __uint_as_float(0x7fc00000u);
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
__device__ __forceinline__ unsigned long long make_key(float s, uint32_t id) {
    return (static_cast<unsigned long long>(score_key32(s)) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - id);
}

// Sorted (descending) list of 32*E keys held by one warp, blocked layout idx = lane*E + j.
template <int E>
__device__ __forceinline__ void list_insert(unsigned long long (&L)[E], unsigned long long c, int lane) {
    int pos = 0;
#pragma unroll
    for (int j = 0; j < E; ++j) pos += (L[j] > c) ? 1 : 0;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) pos += __shfl_xor_sync(0xffffffffu, pos, off);
    const unsigned long long carry = __shfl_up_sync(0xffffffffu, L[E - 1], 1);
    if (pos >= 32 * E) return;
#pragma unroll
    for (int j = E - 1; j >= 0; --j) {
        const int g = lane * E + j;
        const unsigned long long prev = (j > 0) ? L[j - 1] : carry;
        if (g > pos) L[j] = prev;
        else if (g == pos) L[j] = c;
    }
}
template <int E>
__device__ __forceinline__ unsigned long long list_kth(const unsigned long long (&L)[E], int k) {
    unsigned long long mine = 0ull;
#pragma unroll
    for (int j = 0; j < E; ++j)
        if (j == ((k - 1) % E)) mine = L[j];
    return __shfl_sync(0xffffffffu, mine, (k - 1) / E);
}

struct ScanParams {
    const float* db;        // [n_rows][d]
    const float* rdb;       // [n_rows] 1/(|x|^2 + 1e-12)
    long long n_rows;
    int d;
    const float* q;         // [nq][d] queries or centroids
    const float* rq;        // [nq] 1/(|q|^2 + 1e-12)
    const float* c2;        // [nq] 0.5*|c|^2 (kmeans)
    int nq;
    // search
    int k;
    unsigned long long* partial;   // [splits][nq][k]
    long long rows_per_split;
    // kmeans / cosmin
    int* labels;
    float* cosv;
    unsigned long long* acc;       // [nq][d] int64 fixed-point sums
    unsigned long long* cnt;       // [nq]
    double sc;
    int smem_acc;
};

// Stage one d-chunk of the row tile and of the query tile (transposed: [i][row]).
template <int TQ>
__device__ __forceinline__ void stage_chunk(const ScanParams& p, float* xs, float* qs, long long row0, long long row_end,
                                            int qbase, int i0) {
    constexpr int QT = 16 * TQ;
    const int tid = threadIdx.x;
#pragma unroll
    for (int t = 0; t < RT * DK / kThreads; ++t) {
        const int e = tid + kThreads * t;
        const int r = e >> 5, c = e & 31;
        const long long row = row0 + r;
        float v = 0.0f;
        if (row < row_end && i0 + c < p.d) v = __ldg(p.db + row * p.d + i0 + c);
        xs[c * XS + r] = v;
    }
#pragma unroll
    for (int t = 0; t < QT * DK / kThreads; ++t) {
        const int e = tid + kThreads * t;
        const int r = e >> 5, c = e & 31;
        float v = 0.0f;
        if (qbase + r < p.nq && i0 + c < p.d) v = __ldg(p.q + static_cast<long long>(qbase + r) * p.d + i0 + c);
        qs[c * QT + r] = v;
    }
}

// acc[v][u] = fmaf chain over the whole d for query tq*TQ+v and row r(u).
template <int TQ>
__device__ __forceinline__ void dot_tile(const ScanParams& p, float* xs, float* qs, long long row0, long long row_end,
                                         int qbase, float (&acc)[TQ][8]) {
    constexpr int QT = 16 * TQ;
    const int tq = threadIdx.x >> 4, tr = threadIdx.x & 15;
#pragma unroll
    for (int v = 0; v < TQ; ++v)
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[v][u] = 0.0f;
    for (int i0 = 0; i0 < p.d; i0 += DK) {
        __syncthreads();
        stage_chunk<TQ>(p, xs, qs, row0, row_end, qbase, i0);
        __syncthreads();
        const int kk = min(DK, p.d - i0);
        for (int i = 0; i < kk; ++i) {
            float qv[TQ];
            if (TQ == 4) {
                const float4 t4 = *reinterpret_cast<const float4*>(qs + i * QT + tq * 4);
                qv[0] = t4.x; qv[1 % TQ] = t4.y; qv[2 % TQ] = t4.z; qv[3 % TQ] = t4.w;
            } else {
#pragma unroll
                for (int v = 0; v < TQ; ++v) qv[v] = qs[i * QT + tq * TQ + v];
            }
            const float4 xa = *reinterpret_cast<const float4*>(xs + i * XS + tr * 4);
            const float4 xb = *reinterpret_cast<const float4*>(xs + i * XS + 64 + tr * 4);
            const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
            for (int v = 0; v < TQ; ++v)
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[v][u] = __fmaf_rn(qv[v], xv[u], acc[v][u]);
        }
    }
}
__device__ __forceinline__ int row_of(int tr, int u) { return tr * 4 + (u & 3) + ((u >> 2) << 6); }
__device__ __forceinline__ float cos_from(float dot, float ra, float rb) {
    return __fmul_rn(dot, __fsqrt_rn(__fmul_rn(ra, rb)));
}

// ------------------------------------------------------------------ search
// grid = (splits, ceil(nq / (16*TQ))).  K2 = 32*E >= k.
template <int TQ, int E>
__global__ void __launch_bounds__(kThreads)
search_kernel(const ScanParams p) {
    constexpr int QT = 16 * TQ;
    constexpr int K2 = 32 * E;
    extern __shared__ __align__(16) uint8_t sm[];
    float* xs = reinterpret_cast<float*>(sm);
    float* qs = xs + DK * XS;
    unsigned long long* lists = reinterpret_cast<unsigned long long*>(qs + DK * QT);
    unsigned long long* cand = lists + QT * K2;
    unsigned long long* tau = cand + QT * CAP;
    int* ccount = reinterpret_cast<int*>(tau + QT);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tq = tid >> 4, tr = tid & 15;
    const int qbase = blockIdx.y * QT;
    for (int i = tid; i < QT * K2; i += kThreads) lists[i] = 0ull;
    for (int i = tid; i < QT; i += kThreads) { tau[i] = 0ull; ccount[i] = 0; }
    __syncthreads();

    const long long r_begin = static_cast<long long>(blockIdx.x) * p.rows_per_split;
    const long long r_end = min(p.n_rows, r_begin + p.rows_per_split);
    for (long long row0 = r_begin; row0 < r_end; row0 += RT) {
        float acc[TQ][8];
        dot_tile<TQ>(p, xs, qs, row0, r_end, qbase, acc);
        // scores in place
        float rxv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long row = row0 + row_of(tr, u);
            rxv[u] = row < r_end ? __ldg(p.rdb + row) : 0.0f;
        }
        unsigned pend = 0u;
#pragma unroll
        for (int v = 0; v < TQ; ++v) {
            const int ql = tq * TQ + v;
            const bool qok = qbase + ql < p.nq;
            const float rqv = qok ? __ldg(p.rq + qbase + ql) : 0.0f;
            const unsigned long long t = tau[ql];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const long long row = row0 + row_of(tr, u);
                acc[v][u] = cos_from(acc[v][u], rqv, rxv[u]);
                if (qok && row < r_end && make_key(acc[v][u], static_cast<uint32_t>(row)) > t) pend |= 1u << (v * 8 + u);
            }
        }
        while (__syncthreads_or(pend != 0u)) {
            // append candidates that still beat the current k-th best
#pragma unroll
            for (int v = 0; v < TQ; ++v) {
                const int ql = tq * TQ + v;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const unsigned bit = 1u << (v * 8 + u);
                    if (pend & bit) {
                        const unsigned long long key = make_key(acc[v][u], static_cast<uint32_t>(row0 + row_of(tr, u)));
                        if (key > tau[ql]) {
                            const int slot = atomicAdd(&ccount[ql], 1);
                            if (slot < CAP) { cand[ql * CAP + slot] = key; pend &= ~bit; }
                        } else {
                            pend &= ~bit;
                        }
                    }
                }
            }
            __syncthreads();
            // merge: one warp per query, sequential insertion into the sorted list
            for (int ql = warp; ql < QT; ql += kThreads / 32) {
                const int n = min(ccount[ql], CAP);
                if (n > 0) {
                    unsigned long long L[E];
#pragma unroll
                    for (int j = 0; j < E; ++j) L[j] = lists[ql * K2 + lane * E + j];
                    for (int t = 0; t < n; ++t) list_insert<E>(L, cand[ql * CAP + t], lane);
#pragma unroll
                    for (int j = 0; j < E; ++j) lists[ql * K2 + lane * E + j] = L[j];
                    const unsigned long long kth = list_kth<E>(L, p.k);
                    __syncwarp();
                    if (lane == 0) { tau[ql] = kth; ccount[ql] = 0; }
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < QT * p.k; i += kThreads) {
        const int ql = i / p.k, t = i - ql * p.k;
        if (qbase + ql < p.nq)
            p.partial[(static_cast<long long>(blockIdx.x) * p.nq + qbase + ql) * p.k + t] = lists[ql * K2 + t];
    }
}

// ------------------------------------------------------------------ search, d <= 128, many queries
// Same contract and result as search_kernel<4, E>, restructured for throughput: the block's 64
// queries are staged ONCE, each 128-row tile is staged whole (one barrier pair per tile instead
// of one per 32 columns), and candidates are pre-filtered with an approximate score
// (sqrt.approx, no IEEE sqrt) against the current k-th best minus a 2^-18 relative margin; only
// survivors pay for the exact score and the 64-bit key.  The pre-filter can only pass extra
// candidates, never drop a true one, so the result is bit-identical.
template <int E>
__global__ void __launch_bounds__(kThreads)
search_kernel_wide(const ScanParams p) {
    constexpr int TQ = 4, QT = 64, K2 = 32 * E;
    extern __shared__ __align__(16) uint8_t sm[];
    const int d = p.d;
    float* xs = reinterpret_cast<float*>(sm);                 // [d][XS]
    float* qs = xs + d * XS;                                  // [d][QT]
    unsigned long long* lists = reinterpret_cast<unsigned long long*>(qs + d * QT);
    unsigned long long* cand = lists + QT * K2;
    unsigned long long* tau = cand + QT * CAP;
    int* ccount = reinterpret_cast<int*>(tau + QT);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tq = tid >> 4, tr = tid & 15;
    const int qbase = blockIdx.y * QT;
    for (int i = tid; i < QT * K2; i += kThreads) lists[i] = 0ull;
    for (int i = tid; i < QT; i += kThreads) { tau[i] = 0ull; ccount[i] = 0; }
    // queries: each warp stages 8 of the 64 query rows, lanes run along d (coalesced)
    for (int r = warp; r < QT; r += kThreads / 32)
        for (int c = lane; c < d; c += 32)
            qs[c * QT + r] = (qbase + r < p.nq) ? __ldg(p.q + static_cast<long long>(qbase + r) * d + c) : 0.0f;
    float rqv[TQ];
    bool qok[TQ];
#pragma unroll
    for (int v = 0; v < TQ; ++v) {
        qok[v] = qbase + tq * TQ + v < p.nq;
        rqv[v] = qok[v] ? __ldg(p.rq + qbase + tq * TQ + v) : 0.0f;
    }
    __syncthreads();

    const long long r_begin = static_cast<long long>(blockIdx.x) * p.rows_per_split;
    const long long r_end = min(p.n_rows, r_begin + p.rows_per_split);
    // Register-staged double buffering: the next tile's 128 x d floats are loaded (coalesced,
    // 16 rows per warp, lanes along d) while the current tile is being multiplied.
    float pre[16][4];
    auto prefetch = [&](long long row0) {
#pragma unroll
        for (int a = 0; a < 16; ++a) {
            const long long row = row0 + warp + 8 * a;
            const bool ok = row < r_end;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int c = lane + 32 * b;
                pre[a][b] = (ok && c < d) ? __ldg(p.db + row * d + c) : 0.0f;
            }
        }
    };
    prefetch(r_begin);
    for (long long row0 = r_begin; row0 < r_end; row0 += RT) {
        __syncthreads();                                      // previous tile fully consumed
#pragma unroll
        for (int a = 0; a < 16; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int c = lane + 32 * b;
                if (c < d) xs[c * XS + warp + 8 * a] = pre[a][b];
            }
        __syncthreads();
        if (row0 + RT < r_end) prefetch(row0 + RT);
        float acc[TQ][8];
#pragma unroll
        for (int v = 0; v < TQ; ++v)
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[v][u] = 0.0f;
#pragma unroll 4
        for (int i = 0; i < d; ++i) {
            const float4 t4 = *reinterpret_cast<const float4*>(qs + i * QT + tq * 4);
            const float4 xa = *reinterpret_cast<const float4*>(xs + i * XS + tr * 4);
            const float4 xb = *reinterpret_cast<const float4*>(xs + i * XS + 64 + tr * 4);
            const float qv[4] = {t4.x, t4.y, t4.z, t4.w};
            const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
            for (int v = 0; v < TQ; ++v)
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[v][u] = __fmaf_rn(qv[v], xv[u], acc[v][u]);
        }
        float rxv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long row = row0 + row_of(tr, u);
            rxv[u] = row < r_end ? __ldg(p.rdb + row) : 0.0f;
        }
        unsigned pend = 0u;
#pragma unroll
        for (int v = 0; v < TQ; ++v) {
            const unsigned long long t = tau[tq * TQ + v];
            const uint32_t thi = static_cast<uint32_t>(t >> 32);
            const float thr = score_unkey32(thi);
            const bool pass_all = thi == 0u || !(fabsf(thr) < 3.0e38f);   // list not full, k-th entry NaN, or infinite
            const float thr_adj = thr - fabsf(thr) * 3.814697265625e-06f;   // 2^-18 relative margin
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float rw = rqv[v] * rxv[u];
                float sq;
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(rw));
                const float sa = acc[v][u] * sq;
                const bool maybe = pass_all || sa >= thr_adj || rw < 1e-30f;
                if (maybe && qok[v] && row0 + row_of(tr, u) < r_end) pend |= 1u << (v * 8 + u);
            }
        }
        while (__syncthreads_or(pend != 0u)) {
#pragma unroll
            for (int v = 0; v < TQ; ++v) {
                const int ql = tq * TQ + v;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const unsigned bit = 1u << (v * 8 + u);
                    if (pend & bit) {
                        const float s = cos_from(acc[v][u], rqv[v], rxv[u]);          // exact score
                        const unsigned long long key = make_key(s, static_cast<uint32_t>(row0 + row_of(tr, u)));
                        if (key > tau[ql]) {
                            const int slot = atomicAdd(&ccount[ql], 1);
                            if (slot < CAP) { cand[ql * CAP + slot] = key; pend &= ~bit; }
                        } else {
                            pend &= ~bit;
                        }
                    }
                }
            }
            __syncthreads();
            for (int ql = warp; ql < QT; ql += kThreads / 32) {
                const int n = min(ccount[ql], CAP);
                if (n > 0) {
                    unsigned long long L[E];
#pragma unroll
                    for (int j = 0; j < E; ++j) L[j] = lists[ql * K2 + lane * E + j];
                    for (int t = 0; t < n; ++t) list_insert<E>(L, cand[ql * CAP + t], lane);
#pragma unroll
                    for (int j = 0; j < E; ++j) lists[ql * K2 + lane * E + j] = L[j];
                    const unsigned long long kth = list_kth<E>(L, p.k);
                    __syncwarp();
                    if (lane == 0) { tau[ql] = kth; ccount[ql] = 0; }
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < QT * p.k; i += kThreads) {
        const int ql = i / p.k, t = i - ql * p.k;
        if (qbase + ql < p.nq)
            p.partial[(static_cast<long long>(blockIdx.x) * p.nq + qbase + ql) * p.k + t] = lists[ql * K2 + t];
    }
}

// ------------------------------------------------------------------ wide search, d % 4 == 0
// Same contract and candidate machinery as search_kernel_wide, different data movement: the
// row tile and the queries stay ROW-major in shared memory (stride S floats, S % 32 == 4, so
// the 16 lanes of a half-warp reading 16 consecutive rows with 16-byte loads hit 32 distinct
// banks), tiles arrive by 16-byte cp.async (the register-staged prefetch of the older kernel was
// serialised by the compiler into load->store pairs: 40% of its stall samples), and the inner
// loop walks d four elements at a time.  ONE tile buffer per block keeps shared memory at
// ~110 KB so two blocks share an SM (16 warps): the other block's math hides this block's load.  Each (query, row)
// accumulator still sees fmaf(q_i, x_i, acc) for i = 0 .. d-1 in order: bit-identical.
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
    const int n = valid ? 16 : 0;                             // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__host__ __device__ __forceinline__ int wide4_stride(int d) { return d + ((4 - (d % 32) + 32) % 32); }   // smallest S >= d, S % 32 == 4

template <int E>
__global__ void __launch_bounds__(kThreads, 2)
search_kernel_wide4(const ScanParams p) {
    constexpr int TQ = 4, QT = 64, K2 = 32 * E;
    extern __shared__ __align__(16) uint8_t sm[];
    const int d = p.d, S = wide4_stride(d), d4 = d >> 2;
    const unsigned d4magic = 0xFFFFFFFFu / static_cast<unsigned>(d4) + 1u;   // ceil(2^32 / d4)
    float* xs = reinterpret_cast<float*>(sm);                 // [RT][S]
    float* qs = xs + RT * S;                                  // [QT][S]
    unsigned long long* lists = reinterpret_cast<unsigned long long*>(qs + QT * S);
    unsigned long long* cand = lists + QT * K2;
    unsigned long long* tau = cand + QT * CAP;
    int* ccount = reinterpret_cast<int*>(tau + QT);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tq = tid >> 4, tr = tid & 15;
    auto row_at = [&](int u) { return u * 16 + tr; };         // this thread's 8 rows of the tile
    const int qbase = blockIdx.y * QT;
    for (int i = tid; i < QT * K2; i += kThreads) lists[i] = 0ull;
    for (int i = tid; i < QT; i += kThreads) { tau[i] = 0ull; ccount[i] = 0; }
    for (int g = tid; g < QT * d4; g += kThreads) {           // queries, row-major
        const int r = g / d4, c4 = g - r * d4;
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (qbase + r < p.nq) v = __ldg(reinterpret_cast<const float4*>(p.q + static_cast<long long>(qbase + r) * d) + c4);
        *reinterpret_cast<float4*>(qs + r * S + 4 * c4) = v;
    }
    float rqv[TQ];
    bool qok[TQ];
#pragma unroll
    for (int v = 0; v < TQ; ++v) {
        qok[v] = qbase + tq * TQ + v < p.nq;
        rqv[v] = qok[v] ? __ldg(p.rq + qbase + tq * TQ + v) : 0.0f;
    }

    const long long r_begin = static_cast<long long>(blockIdx.x) * p.rows_per_split;
    const long long r_end = min(p.n_rows, r_begin + p.rows_per_split);
    auto issue = [&](long long row0, float* dst) {            // one row tile -> smem (rows past the end: zeros)
        for (int g = tid; g < RT * d4; g += kThreads) {
            const int r = d4 == 1 ? g : static_cast<int>(__umulhi(static_cast<unsigned>(g), d4magic)), c4 = g - r * d4;   // exact floor(g / d4) for g < 2^30
            const long long row = row0 + r;
            const bool ok = row < r_end;
            cp_async16_zfill(dst + r * S + 4 * c4, p.db + (ok ? row : r_begin) * d + 4 * c4, ok);
        }
        cp_async_commit_group();
    };
    for (long long row0 = r_begin; row0 < r_end; row0 += RT) {
        __syncthreads();                                      // everyone is done with the previous tile (and the set-up above)
        issue(row0, xs);
        cp_async_wait_all();
        __syncthreads();
        float acc[TQ][8];
#pragma unroll
        for (int v = 0; v < TQ; ++v)
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[v][u] = 0.0f;
#pragma unroll 2
        for (int i = 0; i < d; i += 4) {
            float4 qv[TQ], xv[8];
#pragma unroll
            for (int v = 0; v < TQ; ++v) qv[v] = *reinterpret_cast<const float4*>(qs + (tq * TQ + v) * S + i);
#pragma unroll
            for (int u = 0; u < 8; ++u) xv[u] = *reinterpret_cast<const float4*>(xs + row_at(u) * S + i);
#pragma unroll
            for (int v = 0; v < TQ; ++v)
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float a = acc[v][u];
                    a = __fmaf_rn(qv[v].x, xv[u].x, a);
                    a = __fmaf_rn(qv[v].y, xv[u].y, a);
                    a = __fmaf_rn(qv[v].z, xv[u].z, a);
                    a = __fmaf_rn(qv[v].w, xv[u].w, a);
                    acc[v][u] = a;
                }
        }
        float rxv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long row = row0 + row_at(u);
            rxv[u] = row < r_end ? __ldg(p.rdb + row) : 0.0f;
        }
        unsigned pend = 0u;
#pragma unroll
        for (int v = 0; v < TQ; ++v) {
            const unsigned long long t = tau[tq * TQ + v];
            const uint32_t thi = static_cast<uint32_t>(t >> 32);
            const float thr = score_unkey32(thi);
            const bool pass_all = thi == 0u || !(fabsf(thr) < 3.0e38f);   // list not full, k-th entry NaN, or infinite
            const float thr_adj = thr - fabsf(thr) * 3.814697265625e-06f;   // 2^-18 relative margin
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float rw = rqv[v] * rxv[u];
                float sq;
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(rw));
                const float sa = acc[v][u] * sq;
                const bool maybe = pass_all || sa >= thr_adj || rw < 1e-30f;
                if (maybe && qok[v] && row0 + row_at(u) < r_end) pend |= 1u << (v * 8 + u);
            }
        }
        while (__syncthreads_or(pend != 0u)) {
#pragma unroll
            for (int v = 0; v < TQ; ++v) {
                const int ql = tq * TQ + v;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const unsigned bit = 1u << (v * 8 + u);
                    if (pend & bit) {
                        const float s = cos_from(acc[v][u], rqv[v], rxv[u]);          // exact score
                        const unsigned long long key = make_key(s, static_cast<uint32_t>(row0 + row_at(u)));
                        if (key > tau[ql]) {
                            const int slot = atomicAdd(&ccount[ql], 1);
                            if (slot < CAP) { cand[ql * CAP + slot] = key; pend &= ~bit; }
                        } else {
                            pend &= ~bit;
                        }
                    }
                }
            }
            __syncthreads();
            for (int ql = warp; ql < QT; ql += kThreads / 32) {
                const int n = min(ccount[ql], CAP);
                if (n > 0) {
                    unsigned long long L[E];
#pragma unroll
                    for (int j = 0; j < E; ++j) L[j] = lists[ql * K2 + lane * E + j];
                    for (int t = 0; t < n; ++t) list_insert<E>(L, cand[ql * CAP + t], lane);
#pragma unroll
                    for (int j = 0; j < E; ++j) lists[ql * K2 + lane * E + j] = L[j];
                    const unsigned long long kth = list_kth<E>(L, p.k);
                    __syncwarp();
                    if (lane == 0) { tau[ql] = kth; ccount[ql] = 0; }
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < QT * p.k; i += kThreads) {
        const int ql = i / p.k, t = i - ql * p.k;
        if (qbase + ql < p.nq)
            p.partial[(static_cast<long long>(blockIdx.x) * p.nq + qbase + ql) * p.k + t] = lists[ql * K2 + t];
    }
}

// database rows -> query matrix (search by example row, apply_r.lua:268-272); rows this shard
// does not own stay zero so an integer max-allreduce assembles them exactly across ranks.
__global__ void gather_rows_kernel(const float* __restrict__ db, long long n_rows, int d, long long row_offset,
                                   const long long* __restrict__ rows, int Q, float* __restrict__ out) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(Q) * d) return;
    const int q = static_cast<int>(idx / d), c = static_cast<int>(idx - static_cast<long long>(q) * d);
    const long long r = rows[q] - row_offset;
    out[idx] = (r >= 0 && r < n_rows) ? db[r * d + c] : 0.0f;
}

// ------------------------------------------------------------------ label selection rules (kmeans / cosine-min)
struct Best {
    float v;
    int j;
};
// MODE 1 (unsup.kmeans, TH max scan "!(v <= best)", break on NaN): first NaN wins, else the
// largest value, lowest index on ties.
// MODE 2 (apply_r.lua:206-218 "dist < minDist"): a NaN at j == 0 sticks, otherwise NaNs never
// win; smallest value, lowest index on ties.
template <int MODE>
__device__ __forceinline__ bool better(const Best a, const Best b) {   // is a strictly preferable to b?
    if (b.j < 0) return a.j >= 0;
    if (a.j < 0) return false;
    const bool an = a.v != a.v, bn = b.v != b.v;
    if (MODE == 1) {
        if (an || bn) return an && (!bn || a.j < b.j);
        return a.v > b.v || (a.v == b.v && a.j < b.j);
    } else {
        const bool a0 = an && a.j == 0, b0 = bn && b.j == 0;
        if (a0 || b0) return a0;
        if (an || bn) return !an && bn ? true : (an && bn ? a.j < b.j : false);
        return a.v < b.v || (a.v == b.v && a.j < b.j);
    }
}

// kmeans fixed point: llrint(x * 2^shift).  With 2^shift >= 1 the product is exact in fp32 as well (a pure exponent
// change that cannot overflow: |x| * 2^shift < 2^63 by the choice of shift), so one fp32 multiply + one F2I replaces
// F2D + DMUL + D2I -- the conversion unit (16 lanes/clk/SM) is what bounds the centroid-update pass.
struct FixScale {
    double sc;
    float scf;
    bool f32;
};
__device__ __forceinline__ FixScale make_fix_scale(double sc) {
    FixScale f;
    f.sc = sc; f.scf = static_cast<float>(sc);
    f.f32 = static_cast<double>(f.scf) == sc && f.scf >= 1.0f;
    return f;
}
__device__ __forceinline__ long long fix64(float x, const FixScale& f) {
    return f.f32 ? __float2ll_rn(__fmul_rn(x, f.scf)) : __double2ll_rn(static_cast<double>(x) * f.sc);
}

// ------------------------------------------------------------------ streaming kernels, nq <= 32
// The HBM-bound regime (a handful of needles, k = 20 centroids): ONE THREAD PER ROW keeps all nq
// accumulators in registers -- still one sequential fmaf chain per (query,row) pair -- while the
// rows stream through a cp.async double-buffered shared-memory tile (rows of dc_pad floats,
// dc_pad/4 odd so the per-thread 128-bit reads are bank-conflict free) and the queries are
// broadcast from shared memory.  Algorithmic traffic = 4*N*d bytes, read exactly once.
//   MODE 0  cosine top-k      (apply_r.lua:267-282)
//   MODE 1  kmeans label + int64 fixed-point centroid sums (unsup.kmeans, apply_r.lua:198), d <= 128
//   MODE 2  cosine-min assignment (apply_r.lua:206-218)
struct StreamParams {
    ScanParams s;
    int dc, dc_pad, n_chunks;   // columns per staged chunk (multiple of 4), padded row stride, chunks per row
    unsigned c4_magic;          // ceil(2^32 / (dc/4))
    long long tiles_per_block;  // contiguous row tiles per block
    int groups;                 // MODE 1: accumulator copies (thread groups) in shared memory
};
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int SR = 128;        // rows per streamed tile
constexpr int SSTAGES = 3;     // cp.async ring depth

// 256 threads: thread t owns row (t & 127) and the query half (t >> 7), i.e. NQ/2 accumulators.
template <int NQ, int MODE, int E>
__global__ void __launch_bounds__(kThreads)
stream_kernel(const StreamParams sp) {
    const ScanParams& p = sp.s;
    const FixScale fx = make_fix_scale(MODE == 1 ? p.sc : 1.0);
    constexpr int K2 = 32 * E;
    constexpr int NH = NQ / 2;                                // accumulators per thread
    static_assert(NQ == 4 || NQ % 8 == 0, "NQ/2 is 2 (four needles: the BASELINE config-0 shape, no padded FMAs) or a multiple of 4");
    extern __shared__ __align__(16) uint8_t sm[];
    const int d = p.d, dc = sp.dc, dcp = sp.dc_pad;
    float* stage0 = reinterpret_cast<float*>(sm);             // [SSTAGES][SR][dc_pad]
    float* qs = stage0 + SSTAGES * SR * dcp;                  // [d][NQ]
    uint8_t* tail = reinterpret_cast<uint8_t*>(qs + static_cast<size_t>(d) * NQ);
    // MODE 0
    unsigned long long* lists = reinterpret_cast<unsigned long long*>(tail);
    unsigned long long* cand = lists + NQ * K2;
    unsigned long long* tau = cand + NQ * CAP;
    int* ccount = reinterpret_cast<int*>(tau + NQ);
    // MODE 1 / 2
    unsigned long long* sacc = reinterpret_cast<unsigned long long*>(tail);      // [nq*d + nq]  (MODE 1)
    float* hv = reinterpret_cast<float*>(sacc + (MODE == 1 ? static_cast<size_t>(p.nq) * d + p.nq : 0));   // [SR] half-1 best value
    int* hj = reinterpret_cast<int*>(hv + SR);                // [SR] half-1 best index
    int* slab = hj + SR;                                      // [SR] final labels of the tile
    int* perm = slab + SR;                                    // [SR] rows of the tile grouped by label (MODE 1)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rloc = tid & (SR - 1), qh = tid >> 7;           // row in tile, query half
    const int jbase = qh * NH;
    for (int i = tid; i < d * NQ; i += kThreads) {
        const int c = i / NQ, j = i - c * NQ;
        qs[i] = j < p.nq ? __ldg(p.q + static_cast<long long>(j) * d + c) : 0.0f;
    }
    if (MODE == 0) {
        for (int i = tid; i < NQ * K2; i += kThreads) lists[i] = 0ull;
        for (int i = tid; i < NQ; i += kThreads) { tau[i] = 0ull; ccount[i] = 0; }
    }
    if (MODE == 1)
        for (int i = tid; i < p.nq * d + p.nq; i += kThreads) sacc[i] = 0ull;
    float aux[NH];                                            // rq (MODE 0/2) or c2 (MODE 1) of this thread's queries
#pragma unroll
    for (int j = 0; j < NH; ++j) aux[j] = jbase + j < p.nq ? __ldg((MODE == 1 ? p.c2 : p.rq) + jbase + j) : 0.0f;

    const long long tile_begin = static_cast<long long>(blockIdx.x) * sp.tiles_per_block;
    const long long n_tiles_all = (p.n_rows + SR - 1) / SR;
    const long long tile_end = min(n_tiles_all, tile_begin + sp.tiles_per_block);
    const long long n_steps = max(0ll, tile_end - tile_begin) * sp.n_chunks;

    const unsigned c4magic = sp.c4_magic;                     // ceil(2^32 / (dc/4)): exact floor division for piece indices
    auto issue = [&](long long step) {                        // cp.async one (tile, chunk) into ring slot step % SSTAGES
        if (step < n_steps) {
            const long long tile = tile_begin + step / sp.n_chunks;
            const int c0 = static_cast<int>(step % sp.n_chunks) * dc;
            const int kk4 = min(dc, d - c0) >> 2;             // 16-byte pieces per row in this chunk
            float* buf = stage0 + static_cast<int>(step % SSTAGES) * SR * dcp;
            const long long row0 = tile * SR;
            const int rows_live = static_cast<int>(min(static_cast<long long>(SR), p.n_rows - row0));
            const float* src0 = p.db + row0 * d + c0;
            if (kk4 * 4 == d && dcp == d) {                   // whole rows, unpadded: the tile is one contiguous block
                const int total = rows_live * kk4;
                for (int g = tid; g < total; g += kThreads) cp_async16(buf + 4 * g, src0 + 4 * g);
            } else {
                const int c4 = dc >> 2;
                const int total = rows_live * c4;
                for (int g = tid; g < total; g += kThreads) {
                    const int r = c4 == 1 ? g : static_cast<int>(__umulhi(static_cast<unsigned>(g), c4magic));
                    const int col = g - r * c4;
                    if (col < kk4) cp_async16(buf + r * dcp + 4 * col, src0 + static_cast<long long>(r) * d + 4 * col);
                }
            }
        }
        cp_async_commit();                                    // (possibly empty) group: keeps the wait counts uniform
    };

    issue(0);
    issue(1);
    float acc[NH];
    for (long long step = 0; step < n_steps; ++step) {
        const int chunk = static_cast<int>(step % sp.n_chunks);
        const long long tile = tile_begin + step / sp.n_chunks;
        __syncthreads();                                      // ring slot (step+2) % 3 == (step-1) % 3 is free again
        issue(step + 2);
        cp_async_wait<2>();                                   // this thread's pieces of `step` have landed
        __syncthreads();                                      // ... and everyone else's
        if (chunk == 0) {
#pragma unroll
            for (int j = 0; j < NH; ++j) acc[j] = 0.0f;
        }
        const int c0 = chunk * dc;
        const int kk = min(dc, d - c0);
        const float* tilep = stage0 + static_cast<int>(step % SSTAGES) * SR * dcp;
        const float* xr = tilep + rloc * dcp;
        // this row's 1/(|x|^2+eps) (MODE 0/2) is fetched now so its latency hides behind the dot products
        const long long row = tile * SR + rloc;
        const bool live = row < p.n_rows;
        float rx = 0.0f;
        if (MODE != 1 && chunk == sp.n_chunks - 1 && live) rx = __ldg(p.rdb + row);
        for (int i = 0; i < kk; i += 4) {
            const float4 x4 = *reinterpret_cast<const float4*>(xr + i);
            const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
            if constexpr (NH == 2) {                           // two queries per thread: 8-byte operand loads
                float2 q2[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) q2[e] = *reinterpret_cast<const float2*>(qs + (c0 + i + e) * NQ + jbase);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    acc[0] = __fmaf_rn(q2[e].x, xv[e], acc[0]);
                    acc[1] = __fmaf_rn(q2[e].y, xv[e], acc[1]);
                }
            } else {
                float4 qv[4][(NH + 3) / 4];                   // all operand loads first, then 4*NH independent-ish FMAs
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int v = 0; v < NH / 4; ++v) qv[e][v] = *reinterpret_cast<const float4*>(qs + (c0 + i + e) * NQ + jbase + 4 * v);
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int v = 0; v < NH / 4; ++v) {
                        acc[4 * v + 0] = __fmaf_rn(qv[e][v].x, xv[e], acc[4 * v + 0]);
                        acc[4 * v + 1] = __fmaf_rn(qv[e][v].y, xv[e], acc[4 * v + 1]);
                        acc[4 * v + 2] = __fmaf_rn(qv[e][v].z, xv[e], acc[4 * v + 2]);
                        acc[4 * v + 3] = __fmaf_rn(qv[e][v].w, xv[e], acc[4 * v + 3]);
                    }
            }
        }
        if (chunk != sp.n_chunks - 1) continue;
        // ---------------------------------------------------------------- row finished
        if (MODE == 0) {
            unsigned pend = 0u;
#pragma unroll
            for (int j = 0; j < NH; ++j) {
                acc[j] = cos_from(acc[j], aux[j], rx);        // exact score, in place
                if (live && jbase + j < p.nq && make_key(acc[j], static_cast<uint32_t>(row)) > tau[jbase + j]) pend |= 1u << j;
            }
            while (__syncthreads_or(pend != 0u)) {
#pragma unroll
                for (int j = 0; j < NH; ++j) {
                    if (pend & (1u << j)) {
                        const unsigned long long key = make_key(acc[j], static_cast<uint32_t>(row));
                        if (key > tau[jbase + j]) {
                            const int slot = atomicAdd(&ccount[jbase + j], 1);
                            if (slot < CAP) { cand[(jbase + j) * CAP + slot] = key; pend &= ~(1u << j); }
                        } else {
                            pend &= ~(1u << j);
                        }
                    }
                }
                __syncthreads();
                for (int j = warp; j < NQ; j += kThreads / 32) {
                    const int n = min(ccount[j], CAP);
                    if (n > 0) {
                        unsigned long long L[E];
#pragma unroll
                        for (int t = 0; t < E; ++t) L[t] = lists[j * K2 + lane * E + t];
                        for (int t = 0; t < n; ++t) list_insert<E>(L, cand[j * CAP + t], lane);
#pragma unroll
                        for (int t = 0; t < E; ++t) lists[j * K2 + lane * E + t] = L[t];
                        const unsigned long long kth = list_kth<E>(L, p.k);
                        __syncwarp();
                        if (lane == 0) { tau[j] = kth; ccount[j] = 0; }
                    }
                }
            }
        } else {
            // this thread's half of the centroids, scanned in index order; the two halves are then
            // combined with the same total order, which equals the reference's sequential scan
            Best b;
            b.v = 0.0f; b.j = -1;
#pragma unroll
            for (int j = 0; j < NH; ++j) {
                if (jbase + j < p.nq) {
                    Best c;
                    c.j = jbase + j;
                    c.v = MODE == 1 ? __fsub_rn(acc[j], aux[j]) : cos_from(acc[j], rx, aux[j]);
                    if (better<MODE>(c, b)) b = c;
                }
            }
            if (qh == 1) { hv[rloc] = b.v; hj[rloc] = b.j; }
            __syncthreads();
            if (qh == 0) {
                Best o;
                o.v = hv[rloc]; o.j = hj[rloc];
                if (better<MODE>(o, b)) b = o;
                slab[rloc] = live ? b.j : -1;
                if (live) {
                    p.labels[row] = b.j;
                    if (MODE == 2) p.cosv[row] = b.v;
                }
            }
            if (MODE == 1) {
                // centroid sums.  (1) counting-sort the tile's rows by label: perm[] lists the rows
                // grouped by label.  (2) thread (g, c) owns column c of accumulator copy g and walks
                // perm[g], perm[g+groups], ...: consecutive rows mostly share a label, so the int64
                // fixed-point sum of a run lives in a register and shared memory is touched once per
                // run -- no atomics, no per-row read-modify-write chain.  Integer adds are associative,
                // so neither the order nor the split can change the result.
                __syncthreads();
                if (tid < SR) {
                    const int mine = slab[tid];
                    int rank = 0;
                    for (int r = 0; r < SR; ++r) {
                        const int o = slab[r];
                        rank += (o < mine || (o == mine && r < tid)) ? 1 : 0;
                    }
                    perm[rank] = tid;                          // dead rows (label -1) sort first
                }
                __syncthreads();
                // thread (g, c4): 4 consecutive columns of a contiguous segment of the sorted rows
                const int tpg = d >> 2;                       // threads per group
                const int g = tid / tpg, c = (tid - g * tpg) * 4;
                if (g < sp.groups) {
                    const int seg = (SR + sp.groups - 1) / sp.groups;
                    const int pos_end = min(SR, (g + 1) * seg);
                    int cur = -1;
                    long long run0 = 0, run1 = 0, run2 = 0, run3 = 0, cnt = 0;
                    auto flush = [&]() {
                        if (cur >= 0) {
                            unsigned long long* a = sacc + cur * d + c;
                            atomicAdd(a + 0, static_cast<unsigned long long>(run0));
                            atomicAdd(a + 1, static_cast<unsigned long long>(run1));
                            atomicAdd(a + 2, static_cast<unsigned long long>(run2));
                            atomicAdd(a + 3, static_cast<unsigned long long>(run3));
                            if (c == 0) atomicAdd(sacc + p.nq * d + cur, static_cast<unsigned long long>(cnt));
                        }
                    };
                    for (int pos = g * seg; pos < pos_end; ++pos) {
                        const int r = perm[pos];
                        const int lab = slab[r];
                        if (lab < 0) continue;
                        if (lab != cur) { flush(); cur = lab; run0 = run1 = run2 = run3 = 0; cnt = 0; }
                        const float4 x4 = *reinterpret_cast<const float4*>(tilep + r * dcp + c);
                        run0 += fix64(x4.x, fx);
                        run1 += fix64(x4.y, fx);
                        run2 += fix64(x4.z, fx);
                        run3 += fix64(x4.w, fx);
                        ++cnt;
                    }
                    flush();
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    if (MODE == 0) {
        for (int i = tid; i < NQ * p.k; i += kThreads) {
            const int j = i / p.k, t = i - j * p.k;
            if (j < p.nq) p.partial[(static_cast<long long>(blockIdx.x) * p.nq + j) * p.k + t] = lists[j * K2 + t];
        }
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
}

// Merge [parts][nq][k] partial lists per query (one warp per query).
//   mode 0: write ids (int64, + id_offset) and scores
//   mode 1: write keys re-based to global ids (for the NCCL allgather)
// ------------------------------------------------------------------ register-tiled labelling, 9 <= nq <= 32
// kmeans (MODE 1) and cosine-min (MODE 2) at k = 20 perform 20+ FMAs per loaded float: an operand fetched
// from shared memory for EVERY FMA caps the SM at 32 FMA lanes per clock (the 128 B/clk shared->register
// path), which is where stream_kernel sits (0.7 TB/s).  Here each thread owns a 4-query x 4-row register
// tile: a warp is one group of 4 queries (its query loads are warp-uniform broadcasts), its lanes are rows
// (16-byte loads of 32 consecutive rows, stride S = 4 mod 32 floats: conflict-free), so 8 shared loads feed
// 64 FMAs.  Every (query,row) accumulator is still one sequential fmaf chain over d.  The per-row winner is
// the thread's best of its 4 queries (index order), then the NQ/4 warps' winners are merged in index order
// with the same comparator -- the sequential scan of the reference.  Row tiles arrive by 16-byte cp.async
// into ONE buffer; 2-3 blocks per SM overlap each other's loads.  MODE 1 then runs stream_kernel's
// counting-sort / int64 run-sum update on the same row-major tile.
template <int NQ, int MODE>
__global__ void __launch_bounds__(NQ * 8)
rtile_kernel(const StreamParams sp) {
    static_assert(MODE == 1 || MODE == 2, "labelling modes only");
    static_assert(NQ == 16 || NQ == 24 || NQ == 32, "NQ/4 warps, at least 128 threads");
    const ScanParams& p = sp.s;
    const FixScale fx = make_fix_scale(MODE == 1 ? p.sc : 1.0);
    constexpr int T = NQ * 8, NW = NQ / 4;
    extern __shared__ __align__(16) uint8_t sm[];
    const int d = p.d, S = wide4_stride(d), d4 = d >> 2;
    const unsigned d4magic = 0xFFFFFFFFu / static_cast<unsigned>(d4) + 1u;   // ceil(2^32 / d4)
    float* xs = reinterpret_cast<float*>(sm);                 // [SR][S]
    float* qs = xs + SR * S;                                  // [NQ][S]
    float* pv = qs + NQ * S;                                  // [SR][NW] best value per (row, warp)
    int* pj = reinterpret_cast<int*>(pv + SR * NW);           // [SR][NW] best index
    int* slab = pj + SR * NW;                                 // [SR] final labels of the tile
    int* perm = slab + SR;                                    // [SR] rows grouped by label (MODE 1)
    int* lstart = perm + SR;                                  // [NQ + 2] first sorted position of each label (MODE 1)
    unsigned long long* sacc = reinterpret_cast<unsigned long long*>(lstart + NQ + 2);   // [nq*d + nq] (MODE 1; 8-byte aligned: all counts above are even)

    const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
    for (int g = tid; g < NQ * d4; g += T) {                  // centroids / queries, row-major, zero rows past nq
        const int r = g / d4, c4 = g - r * d4;
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (r < p.nq) v = __ldg(reinterpret_cast<const float4*>(p.q + static_cast<long long>(r) * d) + c4);
        *reinterpret_cast<float4*>(qs + r * S + 4 * c4) = v;
    }
    if (MODE == 1)
        for (int i = tid; i < p.nq * d + p.nq; i += T) sacc[i] = 0ull;
    float aux[4];                                             // c2 (MODE 1) or rq (MODE 2) of this warp's 4 queries
#pragma unroll
    for (int v = 0; v < 4; ++v) aux[v] = wq * 4 + v < p.nq ? __ldg((MODE == 1 ? p.c2 : p.rq) + wq * 4 + v) : 0.0f;

    const long long tile_begin = static_cast<long long>(blockIdx.x) * sp.tiles_per_block;
    const long long n_tiles_all = (p.n_rows + SR - 1) / SR;
    const long long tile_end = min(n_tiles_all, tile_begin + sp.tiles_per_block);
    for (long long tile = tile_begin; tile < tile_end; ++tile) {
        const long long row0 = tile * SR;
        __syncthreads();                                      // previous tile fully consumed (and the set-up above)
        for (int g = tid; g < SR * d4; g += T) {
            const int r = d4 == 1 ? g : static_cast<int>(__umulhi(static_cast<unsigned>(g), d4magic)), c4 = g - r * d4;   // exact floor(g / d4) for g < 2^30
            const bool ok = row0 + r < p.n_rows;
            cp_async16_zfill(xs + r * S + 4 * c4, p.db + (ok ? row0 + r : row0) * d + 4 * c4, ok);
        }
        cp_async_commit_group();
        float rx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long row = row0 + u * 32 + lane;
            rx[u] = (MODE == 2 && row < p.n_rows) ? __ldg(p.rdb + row) : 0.0f;
        }
        cp_async_wait_all();
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[v][u] = 0.0f;
#pragma unroll 2
        for (int i = 0; i < d; i += 4) {
            float4 qv[4], xv[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) qv[v] = *reinterpret_cast<const float4*>(qs + (wq * 4 + v) * S + i);
#pragma unroll
            for (int u = 0; u < 4; ++u) xv[u] = *reinterpret_cast<const float4*>(xs + (u * 32 + lane) * S + i);
#pragma unroll
            for (int v = 0; v < 4; ++v)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float a = acc[v][u];
                    a = __fmaf_rn(qv[v].x, xv[u].x, a);
                    a = __fmaf_rn(qv[v].y, xv[u].y, a);
                    a = __fmaf_rn(qv[v].z, xv[u].z, a);
                    a = __fmaf_rn(qv[v].w, xv[u].w, a);
                    acc[v][u] = a;
                }
        }
        // this thread's winner per row among its 4 queries, in index order
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            Best b;
            b.v = 0.0f; b.j = -1;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                if (wq * 4 + v < p.nq) {
                    Best c;
                    c.j = wq * 4 + v;
                    c.v = MODE == 1 ? __fsub_rn(acc[v][u], aux[v]) : cos_from(acc[v][u], rx[u], aux[v]);
                    if (better<MODE>(c, b)) b = c;
                }
            }
            pv[(u * 32 + lane) * NW + wq] = b.v;
            pj[(u * 32 + lane) * NW + wq] = b.j;
        }
        __syncthreads();
        if (tid < SR) {                                       // merge the warps' winners in index order
            Best b;
            b.v = pv[tid * NW]; b.j = pj[tid * NW];
#pragma unroll
            for (int w = 1; w < NW; ++w) {
                Best o;
                o.v = pv[tid * NW + w]; o.j = pj[tid * NW + w];
                if (better<MODE>(o, b)) b = o;
            }
            const long long row = row0 + tid;
            const bool live = row < p.n_rows;
            slab[tid] = live ? b.j : -1;
            if (live) {
                p.labels[row] = b.j;
                if (MODE == 2) p.cosv[row] = b.v;
            }
        }
        if (MODE == 1) {
            // centroid sums.  (1) counting-sort the tile's rows by label (perm[] = rows grouped by label, lstart[L] =
            // first position of label L).  (2) thread (g, c4) owns 4 columns of the labels L = g, g + groups, ...: it
            // walks that label's rows with int64 run sums in registers and adds them to the block's accumulator with a
            // plain read-modify-write -- every (label, column) has exactly one owner, so no atomics (64-bit shared
            // atomics are CAS spin loops).  Integer adds are associative: neither order nor split changes the result.
            __syncthreads();
            if (tid < SR) {
                const int mine = slab[tid];
                int rank = 0;
                for (int r = 0; r < SR; ++r) {
                    const int o = slab[r];
                    rank += (o < mine || (o == mine && r < tid)) ? 1 : 0;
                }
                perm[rank] = tid;                              // dead rows (label -1) sort first
            }
            if (tid <= p.nq) {                                 // lstart[L] = rows with a label < L (L = nq: all rows)
                int c = 0;
                for (int r = 0; r < SR; ++r) c += slab[r] < tid ? 1 : 0;
                lstart[tid] = c;
            }
            __syncthreads();
            const int tpg = d >> 2;                           // threads per group
            const int g = tid / tpg, c = (tid - g * tpg) * 4;
            if (g < sp.groups) {
                for (int L = g; L < p.nq; L += sp.groups) {
                    const int pos0 = lstart[L], pos1 = lstart[L + 1];
                    if (pos0 == pos1) continue;
                    long long run0 = 0, run1 = 0, run2 = 0, run3 = 0;
                    int pos = pos0;
                    for (; pos + 4 <= pos1; pos += 4) {       // four rows in flight: the perm -> row -> convert chain is latency-bound
                        float4 x4[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) x4[e] = *reinterpret_cast<const float4*>(xs + perm[pos + e] * S + c);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            run0 += fix64(x4[e].x, fx);
                            run1 += fix64(x4[e].y, fx);
                            run2 += fix64(x4[e].z, fx);
                            run3 += fix64(x4[e].w, fx);
                        }
                    }
                    for (; pos < pos1; ++pos) {
                        const float4 x4 = *reinterpret_cast<const float4*>(xs + perm[pos] * S + c);
                        run0 += fix64(x4.x, fx);
                        run1 += fix64(x4.y, fx);
                        run2 += fix64(x4.z, fx);
                        run3 += fix64(x4.w, fx);
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
    if (MODE == 1) {
        const int per = p.nq * d + p.nq;
        for (int i = tid; i < per; i += T) {
            const unsigned long long v = sacc[i];
            if (v != 0ull) {
                if (i < p.nq * d) atomicAdd(&p.acc[i], v);
                else atomicAdd(&p.cnt[i - p.nq * d], v);
            }
        }
    }
}

template <int E>
__global__ void __launch_bounds__(kThreads)
merge_kernel(const unsigned long long* __restrict__ partial, int parts, int nq, int k, long long id_offset, int mode,
             long long* __restrict__ ids, float* __restrict__ scores, unsigned long long* __restrict__ keys_out) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    if (q >= nq) return;
    unsigned long long L[E];
#pragma unroll
    for (int j = 0; j < E; ++j) L[j] = 0ull;
    unsigned long long kth = 0ull;
    const int total = parts * k;
    for (int base = 0; base < total; base += 32) {
        const int e = base + lane;
        unsigned long long c = 0ull;
        if (e < total) {
            const int part = e / k, t = e - part * k;
            c = partial[(static_cast<long long>(part) * nq + q) * k + t];
        }
        for (int t = 0; t < 32; ++t) {
            const unsigned long long cc = __shfl_sync(0xffffffffu, c, t);
            if (cc > kth) {   // warp-uniform
                list_insert<E>(L, cc, lane);
                kth = list_kth<E>(L, k);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int t = lane * E + j;
        if (t < k) {
            const unsigned long long key = L[j];
            const long long o = static_cast<long long>(q) * k + t;
            if (mode == 0) {
                if (key == 0ull) { ids[o] = -1; scores[o] = 0.0f; }
                else {
                    ids[o] = id_offset + static_cast<long long>(0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull));
                    scores[o] = score_unkey32(static_cast<uint32_t>(key >> 32));
                }
            } else {
                if (key == 0ull) keys_out[o] = 0ull;
                else {
                    const uint32_t lid = 0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull);
                    const uint32_t gid = static_cast<uint32_t>(id_offset) + lid;
                    keys_out[o] = (key & 0xFFFFFFFF00000000ull) | static_cast<unsigned long long>(0xFFFFFFFFu - gid);
                }
            }
        }
    }
}

// ------------------------------------------------------------------ kmeans label / cosine-min assign (SGEMM-tile path)
template <int TQ, int MODE>
__global__ void __launch_bounds__(kThreads)
assign_kernel(const ScanParams p, const long long n_tiles) {
    const FixScale fx = make_fix_scale(MODE == 1 ? p.sc : 1.0);
    constexpr int QT = 16 * TQ;
    extern __shared__ __align__(16) uint8_t sm[];
    float* xs = reinterpret_cast<float*>(sm);
    float* qs = xs + DK * XS;
    float* redv = qs + DK * QT;                         // [16][RT]
    int* redj = reinterpret_cast<int*>(redv + 16 * RT);  // [16][RT]
    int* slab = redj + 16 * RT;                          // [RT]
    unsigned long long* sacc = reinterpret_cast<unsigned long long*>(slab + RT);   // [nq*d] + [nq] when smem_acc
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tq = tid >> 4, tr = tid & 15;
    const int nacc = p.nq * p.d;
    if (MODE == 1 && p.smem_acc) {
        for (int i = tid; i < nacc + p.nq; i += kThreads) sacc[i] = 0ull;
    }
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * RT;
        Best best[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { best[u].v = 0.0f; best[u].j = -1; }
        float rxv[8];
        if (MODE == 2) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const long long row = row0 + row_of(tr, u);
                rxv[u] = row < p.n_rows ? __ldg(p.rdb + row) : 0.0f;
            }
        }
        for (int qbase = 0; qbase < p.nq; qbase += QT) {
            float acc[TQ][8];
            dot_tile<TQ>(p, xs, qs, row0, p.n_rows, qbase, acc);
#pragma unroll
            for (int v = 0; v < TQ; ++v) {
                const int jg = qbase + tq * TQ + v;
                if (jg < p.nq) {
                    const float aux = MODE == 1 ? __ldg(p.c2 + jg) : __ldg(p.rq + jg);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        Best c;
                        c.j = jg;
                        c.v = MODE == 1 ? __fsub_rn(acc[v][u], aux) : cos_from(acc[v][u], rxv[u], aux);
                        if (better<MODE>(c, best[u])) best[u] = c;
                    }
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            redv[tq * RT + row_of(tr, u)] = best[u].v;
            redj[tq * RT + row_of(tr, u)] = best[u].j;
        }
        __syncthreads();
        if (tid < RT) {
            Best b;
            b.v = redv[tid]; b.j = redj[tid];
            for (int t = 1; t < 16; ++t) {
                Best c;
                c.v = redv[t * RT + tid]; c.j = redj[t * RT + tid];
                if (better<MODE>(c, b)) b = c;
            }
            slab[tid] = b.j;
            const long long row = row0 + tid;
            if (row < p.n_rows) {
                p.labels[row] = b.j;
                if (MODE == 2) p.cosv[row] = b.v;
            }
        }
        __syncthreads();
        if (MODE == 1) {
            const int rows_here = static_cast<int>(min(static_cast<long long>(RT), p.n_rows - row0));
            for (int rr = warp; rr < rows_here; rr += kThreads / 32) {
                const int j = slab[rr];
                const float* xrow = p.db + (row0 + rr) * p.d;
                for (int c = lane; c < p.d; c += 32) {
                    const long long qv = fix64(__ldg(xrow + c), fx);
                    if (p.smem_acc) atomicAdd(&sacc[j * p.d + c], static_cast<unsigned long long>(qv));
                    else atomicAdd(&p.acc[static_cast<long long>(j) * p.d + c], static_cast<unsigned long long>(qv));
                }
                if (lane == 0) {
                    if (p.smem_acc) atomicAdd(&sacc[nacc + j], 1ull);
                    else atomicAdd(&p.cnt[j], 1ull);
                }
            }
        }
    }
    if (MODE == 1 && p.smem_acc) {
        __syncthreads();
        for (int i = tid; i < nacc; i += kThreads)
            if (sacc[i] != 0ull) atomicAdd(&p.acc[i], sacc[i]);
        for (int i = tid; i < p.nq; i += kThreads)
            if (sacc[nacc + i] != 0ull) atomicAdd(&p.cnt[i], sacc[nacc + i]);
    }
}

// centroid = sum / count for non-empty clusters (empty keep the old centroid); add counts to the
// running totals; clear the accumulators for the next iteration.
__global__ void kmeans_finalize_kernel(float* __restrict__ cen, unsigned long long* __restrict__ acc,
                                       unsigned long long* __restrict__ cnt, unsigned long long* __restrict__ total,
                                       int k, int d, double sc) {
    const int j = blockIdx.x;
    const long long c = static_cast<long long>(cnt[j]);
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        if (c != 0) {
            const long long a = static_cast<long long>(acc[static_cast<long long>(j) * d + i]);
            cen[static_cast<long long>(j) * d + i] = static_cast<float>(static_cast<double>(a) / (static_cast<double>(c) * sc));
        }
        acc[static_cast<long long>(j) * d + i] = 0ull;
    }
    __syncthreads();
    if (threadIdx.x == 0) { total[j] += static_cast<unsigned long long>(c); cnt[j] = 0ull; }
}
__global__ void counts_to_float_kernel(const unsigned long long* __restrict__ total, float* __restrict__ out, int k) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < k) out[j] = static_cast<float>(static_cast<long long>(total[j]));
}

// ------------------------------------------------------------------ per-vector prep
// rn[r] = 1/(|x_r|^2 + 1e-12) (fmaf chain), optional half-sq (0.5*|x|^2), optional max|x| (as uint bits).
__global__ void vec_prep_kernel(const float* __restrict__ x, long long n, int d, float* __restrict__ rn,
                                float* __restrict__ halfsq, unsigned int* __restrict__ maxabs_bits) {
    const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    float mx = 0.0f;
    if (r < n) {
        const float* xr = x + r * d;
        float s = 0.0f;
        for (int i = 0; i < d; ++i) {
            const float v = xr[i];
            s = __fmaf_rn(v, v, s);
            const float a = fabsf(v);
            if (!(a <= mx)) mx = a;
        }
        if (rn) rn[r] = __fdiv_rn(1.0f, __fadd_rn(s, 1e-12f));
        if (halfsq) halfsq[r] = __fmul_rn(0.5f, s);
    }
    if (maxabs_bits) {
        unsigned int b = __float_as_uint(mx);   // non-negative floats and NaN order as unsigned ints
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, off));
        if ((threadIdx.x & 31) == 0 && b != 0u) atomicMax(maxabs_bits, b);
    }
}
// cosineSimilarity(v1, v2), apply_r.lua:396-400
__global__ void cosine_pair_kernel(const float* __restrict__ a, const float* __restrict__ b, int d, float* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float dot = 0.0f, na = 0.0f, nb = 0.0f;
    for (int i = 0; i < d; ++i) {
        dot = __fmaf_rn(a[i], b[i], dot);
        na = __fmaf_rn(a[i], a[i], na);
        nb = __fmaf_rn(b[i], b[i], nb);
    }
    const float ra = __fdiv_rn(1.0f, __fadd_rn(na, 1e-12f)), rb = __fdiv_rn(1.0f, __fadd_rn(nb, 1e-12f));
    *out = cos_from(dot, ra, rb);
}

// ------------------------------------------------------------------ cluster members + mean image
// One warp per cluster scans all rows; keeps the best m by (cos desc, id asc).  K2 = 128.
// keys_out (row-sharded database): the kept keys re-based to global ids, for the allgather + merge_kernel; raw_counts: unclipped.
__global__ void __launch_bounds__(32)
cluster_members_kernel(const int* __restrict__ cluster, const float* __restrict__ cosv, long long n, int m, long long id_offset,
                       long long* __restrict__ member_ids, int* __restrict__ member_counts,
                       unsigned long long* __restrict__ keys_out, unsigned long long* __restrict__ raw_counts) {
    constexpr int E = 4;
    const int j = blockIdx.x, lane = threadIdx.x;
    unsigned long long L[E];
#pragma unroll
    for (int t = 0; t < E; ++t) L[t] = 0ull;
    unsigned long long kth = 0ull;
    long long count = 0;
    for (long long base = 0; base < n; base += 32) {
        const long long i = base + lane;
        unsigned long long c = 0ull;
        if (i < n && cluster[i] == j) c = make_key(cosv[i], static_cast<uint32_t>(i));
        unsigned hit = __ballot_sync(0xffffffffu, c != 0ull);
        count += __popc(hit);
        while (hit) {
            const int t = __ffs(hit) - 1;
            hit &= hit - 1;
            const unsigned long long cc = __shfl_sync(0xffffffffu, c, t);
            if (cc > kth) {
                list_insert<E>(L, cc, lane);
                kth = list_kth<E>(L, m);
            }
        }
    }
    const int keep = static_cast<int>(min(count, static_cast<long long>(m)));
    if (lane == 0) {
        member_counts[j] = keep;
        if (raw_counts) raw_counts[j] = static_cast<unsigned long long>(count);
    }
#pragma unroll
    for (int t = 0; t < E; ++t) {
        const int r = lane * E + t;
        if (r < m) {
            const unsigned long long key = L[t];
            const uint32_t lid = 0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull);
            member_ids[static_cast<long long>(j) * m + r] = (r < keep) ? id_offset + static_cast<long long>(lid) : -1;
            if (keys_out)
                keys_out[static_cast<long long>(j) * m + r] =
                    (r < keep) ? ((key & 0xFFFFFFFF00000000ull) | static_cast<unsigned long long>(0xFFFFFFFFu - (static_cast<uint32_t>(id_offset) + lid))) : 0ull;
        }
    }
}
// keep = min(global count, m) after the count allreduce
__global__ void cluster_keep_kernel(const unsigned long long* __restrict__ raw_counts, int k, int m, int* __restrict__ member_counts) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < k) member_counts[j] = static_cast<int>(min(raw_counts[j], static_cast<unsigned long long>(m)));
}
// face = zeros; face:add(img) in member order; face:div(count)   (apply_r.lua:236-242)
__global__ void cluster_mean_kernel(const float* __restrict__ images, int px, const long long* __restrict__ member_ids,
                                    const int* __restrict__ member_counts, int m, float* __restrict__ mean) {
    const int j = blockIdx.y;
    const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (pidx >= px) return;
    const int keep = member_counts[j];
    float s = 0.0f;
    for (int r = 0; r < keep; ++r) s = __fadd_rn(s, images[member_ids[static_cast<long long>(j) * m + r] * px + pidx]);
    mean[static_cast<long long>(j) * px + pidx] = __fdiv_rn(s, static_cast<float>(keep));
}
// Row-sharded images: each rank copies the member images it owns into staged[(j - j0)*m + r][px] (zero bits elsewhere; one
// integer max-allreduce assembles them exactly), then every rank averages the staged rows in member order.
__global__ void cluster_stage_kernel(const float* __restrict__ images, long long n_local, long long id_offset, int px,
                                     const long long* __restrict__ member_ids, int m, int j0, int nj, float* __restrict__ staged) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(nj) * m * px;
    if (idx >= total) return;
    const long long slot = idx / px;
    const int pidx = static_cast<int>(idx - slot * px);
    const long long gid = member_ids[static_cast<long long>(j0) * m + slot];
    const long long lid = gid - id_offset;
    staged[idx] = (gid >= 0 && lid >= 0 && lid < n_local) ? images[lid * px + pidx] : 0.0f;
}
__global__ void cluster_mean_staged_kernel(const float* __restrict__ staged, int px, const int* __restrict__ member_counts, int m, int j0,
                                           float* __restrict__ mean) {
    const int jj = blockIdx.y;
    const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (pidx >= px) return;
    const int keep = member_counts[j0 + jj];
    float s = 0.0f;
    for (int r = 0; r < keep; ++r) s = __fadd_rn(s, staged[(static_cast<long long>(jj) * m + r) * px + pidx]);
    mean[static_cast<long long>(j0 + jj) * px + pidx] = __fdiv_rn(s, static_cast<float>(keep));
}

}  // namespace scan
}  // namespace ganrev
