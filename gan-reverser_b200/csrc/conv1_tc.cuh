// conv1_tc.cuh -- R conv1 (C -> 64, 3x3/pad1, folded BN, ELU; models.lua:399-411) on tcgen05.
//
// The layer has K = 9*C taps per pixel: far too thin for the TMA/implicit-GEMM kernel (its A
// operand would need 64-channel pixels), and on CUDA cores its 576*C FMAs per pixel are bound by
// shared-memory weight broadcasts.  Here the threads build the im2col tile themselves:
//   * thread t of a CTA owns pixel t of a 128-pixel tile: it gathers the 9*C taps from the fp32
//     NCHW image (explicit dropout mask fused, models.lua:399-406), splits every tap into
//     bf16 hi + bf16 lo (x = hi + lo to 2^-16, so the INPUT is not quantised to bf16) and writes
//     its K-major row straight into the 128B-swizzled UMMA layout (row = 128 B, 16-byte chunk c
//     of row r at position c ^ (r & 7));
//   * one thread issues KP/16 tcgen05.mma (M=128, N=64) against the resident weight tile
//     [64][KP] = [w*bnscale | w*bnscale | 0] and commits to an mbarrier;
//   * every warp reads its TMEM lane quarter (one pixel per lane, 64 channels), adds the folded
//     BN shift, applies ELU, packs bf16 and stores the warp's 32 pixels = 4 KB contiguous NHWC
//     through an XOR-swizzled staging buffer (the tile's own A rows, free once the MMA retired).
// Each CTA is software-pipelined over its tiles (two A buffers, two 64-column accumulators: tile
// i+1's MMA is in flight and tile i+2's taps are being fetched while tile i's epilogue runs) and
// four CTAs share an SM.  Nothing here is GEMM-heavy: the point is to take the 576*C FMAs per
// pixel off the FP32 pipe.  What is left is bound by the conversion/transcendental unit (16
// lanes/clk/SM): 64 ex2 (ELU) + 48 bf16 packs per pixel, ~900 cycles per 128-pixel tile.
#pragma once
#include "conv_tc.cuh"

namespace ganrev {
namespace tc {

template <int CIN> struct Conv1Cfg {
    static constexpr int K9 = CIN * 9;
    static constexpr int KP = CIN == 1 ? 32 : 64;          // 2*K9 (hi | lo) padded to a multiple of 16
    static constexpr int kSteps = KP / 16;
    static constexpr int kChunks = KP / 8;                 // 16-byte chunks per row actually used
    static constexpr int kABytes = 128 * 128;              // 128 rows x 128 B
    static constexpr int kBBytes = 64 * 128;
    static constexpr int kSmemBytes = 2 * kABytes + kBBytes + 64 * 4 + 32 + 1024;   // two A tiles, + shift, barriers/slot, alignment slack
};

// wB: bf16 [64][KP] K-major (BN scale folded, hi and lo halves identical); shift: [64].
template <int CIN>
__global__ void __launch_bounds__(128)
r_conv1_tc_kernel(const float* __restrict__ img, const uint8_t* __restrict__ mask, const bf16* __restrict__ wB,
                  const float* __restrict__ shift, bf16* __restrict__ out, int H, int W, int lgW, int lgHW,
                  long long npix_total, int n_tiles, int* err_flag) {
    using C = Conv1Cfg<CIN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    uint4* sA = reinterpret_cast<uint4*>(smem);                         // 2 x [128 rows][8 chunks]
    uint4* sB = reinterpret_cast<uint4*>(smem + 2 * C::kABytes);        // [64 rows][8 chunks]
    float* s_shift = reinterpret_cast<float*>(smem + 2 * C::kABytes + C::kBBytes);
    const uint32_t bar = smem_base + 2 * C::kABytes + C::kBBytes + 64 * 4;   // bar, bar + 8: MMA of the even / odd tiles retired
    const uint32_t tmem_slot = bar + 16;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + 2 * C::kABytes + C::kBBytes + 64 * 4 + 16);

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (t == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<128>(tmem_slot);               // two 64-column accumulators
    if (t < 64) {                                                       // weight row t -> swizzled K-major row
        s_shift[t] = shift[t];
        const uint4* wrow = reinterpret_cast<const uint4*>(wB + static_cast<size_t>(t) * C::KP);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) sB[t * 8 + (c ^ (t & 7))] = wrow[c];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    constexpr uint32_t idesc = make_idesc<64, 128>();
    const uint64_t bdesc = make_smem_desc(smem_base + 2 * C::kABytes);
    const int HW = 1 << lgHW;                                           // H, W are powers of two (check_geom)

    // Raw taps of pixel t of a tile (0 outside the image) and their dropout-mask bytes.  Nothing here
    // CONSUMES a loaded value: the loads stay in flight across the epilogue of the current tile and
    // are first touched when the next tile's row is built.
    // (Row / column validity is decided once per pixel and every tap is ONE predicated load at a warp-uniform offset from the
    // pixel's own address: the first version spent ~22 integer instructions per tap on bounds tests and 64-bit addresses.)
    const bool has_mask = mask != nullptr;
    auto gather = [&](const int tile, float (&x)[C::K9], uint32_t (&mk)[C::K9]) {
        const long long pix = static_cast<long long>(tile) * 128 + t;
        const bool live = pix < npix_total;
        const long long n = pix >> lgHW;
        const int rem = static_cast<int>(pix) & (HW - 1);
        const int h = rem >> lgW, w = rem & (W - 1);
        const bool rv[3] = {live && h > 0, live, live && h < H - 1};
        const bool cv[3] = {w > 0, true, w < W - 1};
        const long long base = n * CIN * static_cast<long long>(HW) + rem;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
            const float* pc = img + base + static_cast<long long>(ci) * HW;
            const uint8_t* pm = has_mask ? mask + base + static_cast<long long>(ci) * HW : nullptr;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const bool ok = rv[ky] && cv[kx];
                    const int off = (ky - 1) * W + (kx - 1);
                    x[(ci * 3 + ky) * 3 + kx] = ok ? __ldg(pc + off) : 0.0f;
                    mk[(ci * 3 + ky) * 3 + kx] = (ok && has_mask) ? static_cast<uint32_t>(__ldg(pm + off)) : 1u;
                }
        }
    };
    float x[C::K9];
    uint32_t mk[C::K9];
    if (static_cast<int>(blockIdx.x) < n_tiles) gather(blockIdx.x, x, mk);

    // Build tile's A rows from the prefetched taps into buffer `b`, then one elected thread issues its MMAs
    // into accumulator `b`; their completion arrives on bar + 8*b.
    auto build_and_issue = [&](const int b) {
        uint4* A = sA + b * (C::kABytes / 16);
        // hi = the tap truncated to its top 16 bits (a bf16 value, exactly), lo = tap - hi (exact in fp32,
        // truncated to bf16 when packed): tap = hi + lo to 2^-16.  Integer ops + one pack per pair -- the
        // conversion unit is this kernel's scarcest pipe (64 ex2 per pixel already go through it).
        float f[C::KP];
#pragma unroll
        for (int k = 2 * C::K9; k < C::KP; ++k) f[k] = 0.0f;
#pragma unroll
        for (int k = 0; k < C::K9; ++k) {
            const float xv = mk[k] != 0u ? x[k] : 0.0f;               // v1 dropout: x*mask, no rescale (models.lua:399-406)
            const float hi = __uint_as_float(__float_as_uint(xv) & 0xFFFF0000u);
            f[k] = hi;
            f[C::K9 + k] = xv - hi;
        }
        // Packing by byte permute (integer pipe), not by cvt (the conversion unit is the bottleneck here): the hi halves ARE bf16 values;
        // the lo halves are truncated instead of rounded (2^-8 of lo = 2^-16 of the tap, the same order as the split itself).
        auto hi2 = [](float a, float b) { return __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x7632); };
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) {
            uint4 pk;
            pk.x = hi2(f[8 * c + 0], f[8 * c + 1]);
            pk.y = hi2(f[8 * c + 2], f[8 * c + 3]);
            pk.z = hi2(f[8 * c + 4], f[8 * c + 5]);
            pk.w = hi2(f[8 * c + 6], f[8 * c + 7]);
            A[t * 8 + (c ^ (t & 7))] = pk;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        tcgen05_fence_before();                                        // (and this thread's TMEM reads of the tile before last are done)
        __syncthreads();
        if (warp == 0 && elect_one_sync()) {
            tcgen05_fence_after();
            const uint64_t adesc = make_smem_desc(smem_base + b * C::kABytes);
#pragma unroll
            for (int k = 0; k < C::kSteps; ++k) umma_bf16(tmem_base + b * 64, adesc + 2u * k, bdesc + 2u * k, idesc, k > 0 ? 1u : 0u);
            umma_commit(bar + 8 * b);
        }
        __syncwarp();
    };
    // Software pipeline: while tile i's epilogue runs, tile i+1's MMA is already in flight (second A buffer,
    // second accumulator) and tile i+2's taps are being fetched.
    const int first = blockIdx.x, step = gridDim.x;
    if (first < n_tiles) {
        build_and_issue(0);
        if (first + step < n_tiles) gather(first + step, x, mk);
    }
    int it = 0;
    for (int tile = first; tile < n_tiles; tile += step, ++it) {
        const int b = it & 1;
        if (tile + step < n_tiles) {
            build_and_issue(b ^ 1);
            if (tile + 2 * step < n_tiles) gather(tile + 2 * step, x, mk);
        }
        mbar_wait(bar + 8 * b, static_cast<uint32_t>(it >> 1) & 1u, err_flag, 106);
        tcgen05_fence_after();
        // ---- epilogue: lane = pixel, 64 channels
        uint32_t r0[32], r1[32];
        const uint32_t tq = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + b * 64;
        tmem_ld32(tq, r0);
        tmem_ld32(tq + 32, r1);
        tmem_ld_wait();
        // staging buffer: this warp's own 32 rows (4 KB) of THIS tile's A buffer, free now that its MMA has retired
        uint4* wstage = sA + b * (C::kABytes / 16) + warp * 256;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int ch = 8 * j + e;
                const float a = __uint_as_float(ch < 32 ? r0[ch & 31] : r1[ch & 31]);
                v[e] = elu_fast(a + s_shift[ch]);
            }
            uint4 pk;
            pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
            pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
            wstage[lane * 8 + (j ^ (lane & 7))] = pk;
        }
        __syncwarp();
        const long long warp_pix0 = static_cast<long long>(tile) * 128 + warp * 32;
        uint4* o = reinterpret_cast<uint4*>(out + warp_pix0 * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = i * 32 + lane;                              // 16-byte element of the warp's 4 KB block
            const int px = e >> 3, c = e & 7;
            if (warp_pix0 + px < npix_total) __stcg(o + e, wstage[px * 8 + (c ^ (px & 7))]);
        }
        __syncwarp();                                                  // stores have read the staging rows before the next gather overwrites them
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<128>(tmem_base); }
}

}  // namespace tc
}  // namespace ganrev
