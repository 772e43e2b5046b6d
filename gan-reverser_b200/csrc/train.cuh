// train.cuh -- one optimisation step of R (train_r.lua:138-170): R_default (models.lua:389-464) in TRAINING
// mode -- batch-norm batch statistics, nn.Dropout / nn.SpatialDropout with the caller's masks -- forward and backward, fp32
// throughout (the reference trains in FloatTensor arithmetic), NCHW like the reference.  At the reference's batch size (32 faces,
// 5.6 GMAC forward, 17 GMAC per step) the step is a chain of ~80 small launches, so these are plain CUDA-core kernels written
// for clarity and determinism (fixed reduction orders, no float atomics): the tensor-core convolution kernels of conv_tc.cuh fold
// eval-mode batch norm into their weights and cannot serve a pass that needs the batch statistics of the raw convolution output.
//
// Upstream semantics restated here (torch/nn, not vendored in the reference):
//   nn.SpatialConvolution 3x3, stride 1, pad 1;  nn.(Spatial)BatchNormalization: eps 1e-5, momentum 0.1, normalises with the
//   biased batch variance, running_var takes the unbiased one;  nn.ELU alpha = 1;  nn.Dropout(p) v2: y = x * mask / (1 - p);
//   nn.Dropout(0.5, true) (the fixer's input layer, v1): y = x * mask;  nn.SpatialDropout(0.25): one Bernoulli(0.75) draw per
//   (sample, channel), no rescaling in training;  nn.SpatialMaxPooling(2,2): first maximum in row-major window order;
//   nn.MSECriterion: mean over all elements;  optim.adam: step = lr sqrt(1 - b2^t) / (1 - b1^t), x -= step * m / (sqrt(v) + eps).
#pragma once
#include "common.cuh"

namespace ganrev {
namespace trn {

constexpr float kBnEps = 1.0e-5f;
constexpr float kBnMomentum = 0.1f;

// ---------------------------------------------------------------- 3x3 convolution, stride 1, pad 1 (forward and backward-data)
// out[n][o][h][w] = bias[o] + sum_i sum_t in[n][i][h+ty-1][w+tx-1] * Wt(o, i, t).  TR = false: Wt = w[(o*CI + i)*9 + t] (forward);
// TR = true: Wt = w[(i*CO + o)*9 + (8 - t)] with CO = channels of THIS kernel's output (backward-data: in = dY, out = dX).
// A thread owns 4 consecutive pixels of a row x 8 output channels: per input channel 3 rows of (one aligned 16-byte load + two edge
// floats) and 18 16-byte broadcast loads of the staged weights [ci][tap][8] feed 288 FMAs (the first version -- one pixel per
// thread, one scalar shared load per FMA -- was shared-load bound).  Each output still adds its products in (ci, tap) order.
// W is a multiple of 4 (H, W are powers of two >= 8 here).
constexpr int kCvO = 8;
template <bool TR>
__global__ void __launch_bounds__(128)
conv3x3_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
               int NB, int CI, int CO, int H, int W) {
    extern __shared__ float4 ws4[];                                // [CI][9][8]
    float* ws = reinterpret_cast<float*>(ws4);
    const int o0 = blockIdx.y * kCvO;
    for (int i = threadIdx.x; i < kCvO * CI * 9; i += blockDim.x) {
        const int oo = i & 7, r = i >> 3, ci = r / 9, t = r - ci * 9;
        const int o = o0 + oo;
        float v = 0.0f;
        if (o < CO) v = TR ? __ldg(w + (static_cast<size_t>(ci) * CO + o) * 9 + (8 - t)) : __ldg(w + (static_cast<size_t>(o) * CI + ci) * 9 + t);
        ws[i] = v;
    }
    __syncthreads();
    // blockIdx.z covers `ipb` images when an image has fewer quads than the block has threads (the 16x16 / 8x8 layers: more warps
    // per staged weight tile), else one image whose quads are split over blockIdx.x
    const int HW = H * W, qpi = HW >> 2, ipb = max(1, static_cast<int>(blockDim.x) / qpi);
    const int g = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.z * ipb + (ipb > 1 ? g / qpi : 0), pix = (ipb > 1 ? g % qpi : g) * 4;
    if (pix >= HW || n >= NB) return;
    const int h = pix / W, x0 = pix - h * W;
    float acc[kCvO][4];
#pragma unroll
    for (int oo = 0; oo < kCvO; ++oo) {
        const float b = (bias != nullptr && o0 + oo < CO) ? __ldg(bias + o0 + oo) : 0.0f;
        acc[oo][0] = acc[oo][1] = acc[oo][2] = acc[oo][3] = b;
    }
    const float* ip = in + static_cast<size_t>(n) * CI * HW + pix;
    // rows h-1..h+1, columns x0-1..x0+4 of one input channel; the NEXT channel's window is requested before the current one's
    // 288 FMAs so that its L1 / L2 latency hides behind them
    auto window = [&](const float* cp, float (&v)[3][6]) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int hh = h + r - 1;
            if (hh >= 0 && hh < H) {
                const float* rp = cp + (r - 1) * W;
                const float4 m = __ldg(reinterpret_cast<const float4*>(rp));
                v[r][0] = x0 > 0 ? __ldg(rp - 1) : 0.0f;
                v[r][1] = m.x; v[r][2] = m.y; v[r][3] = m.z; v[r][4] = m.w;
                v[r][5] = x0 + 4 < W ? __ldg(rp + 4) : 0.0f;
            } else {
#pragma unroll
                for (int c = 0; c < 6; ++c) v[r][c] = 0.0f;
            }
        }
    };
    float vn[3][6];
    window(ip, vn);
    for (int ci = 0; ci < CI; ++ci) {
        float v[3][6];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) v[r][c] = vn[r][c];
        ip += HW;
        if (ci + 1 < CI) window(ip, vn);
        const float4* wr = ws4 + ci * 18;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float4 wa = wr[2 * t], wb = wr[2 * t + 1];
            const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int oo = 0; oo < kCvO; ++oo)
#pragma unroll
                for (int px = 0; px < 4; ++px) acc[oo][px] = fmaf(v[t / 3][t % 3 + px], wv[oo], acc[oo][px]);
        }
    }
#pragma unroll
    for (int oo = 0; oo < kCvO; ++oo)
        if (o0 + oo < CO)
            *reinterpret_cast<float4*>(out + (static_cast<size_t>(n) * CO + o0 + oo) * HW + pix) = make_float4(acc[oo][0], acc[oo][1], acc[oo][2], acc[oo][3]);
}

// dW[o][i][t] = sum_{n,h,w} dY[n][o][h][w] * X[n][i][h+ty-1][w+tx-1].  Block = 16 output x 8 input channels x 9 taps over one slice
// (blockIdx.z of gridDim.z) of the images; the partial sums of a slice go to dw_part[z] and wgrad_sum_kernel adds the slices in
// index order (deterministic, and enough blocks to fill the machine at batch 32).  Per chunk of <= 256 pixels (whole rows of one
// image) the block stages dY[16][chunk] and X[8][rows + 2][W + 8] (zero halo, rows 16-byte aligned) in shared memory -- by cp.async,
// double-buffered, so the next chunk's loads run under this chunk's FMAs; thread
// (pixel lane 0..15, o-group of 4, i-pair) walks the chunk's pixel quads: 4 + 18 shared loads feed 288 FMAs into its 4 x 2 x 9
// register tile (the first version -- positions strided over threads, every operand from L1/L2 -- was 46-60 % of the step).  The 16
// pixel lanes of a tile are the lanes of a half-warp: fixed-order xor-shuffle reduction at the end.
constexpr int kWgO = 16, kWgI = 8, kWgChunk = 256;
__device__ __forceinline__ void wg_cp16(void* smem_dst, const void* gsrc, bool valid) {   // 16-byte cp.async, zero-filled when !valid
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc), "r"(n) : "memory");
}
__global__ void __launch_bounds__(256)
conv3x3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw_part, int B, int CI, int CO, int H, int W) {
    extern __shared__ float4 wg_smem4[];
    const int HW = H * W, chunk = min(kWgChunk, HW), rows = chunk / W, XS = W + 8, xplane = (rows + 2) * XS;
    const int stage_floats = kWgO * kWgChunk + kWgI * xplane;      // one stage: dY [16][chunk] then X [8][rows + 2][XS] (pixel (r, c) at [r + 1][c + 4])
    float* smem = reinterpret_cast<float*>(wg_smem4);              // two stages: the next chunk arrives by cp.async while this one is consumed
    const int o0 = blockIdx.x * kWgO, i0 = blockIdx.y * kWgI;
    const int per = (B + gridDim.z - 1) / gridDim.z, n_begin = blockIdx.z * per, n_end = min(B, n_begin + per);
    const int tid = threadIdx.x, pl = tid & 15, og = (tid >> 4) & 3, ip = tid >> 6;     // pixel lane, o-group (4 channels), i-pair
    float s[4][2][9];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int t = 0; t < 9; ++t) s[a][b][t] = 0.0f;
    const int chunks_per_img = HW / chunk, qpr = W >> 2, nquads = chunk >> 2, xq = XS >> 2;   // xq 16-byte groups per staged row: [halo][W/4][halo]
    const int n_chunks = max(0, n_end - n_begin) * chunks_per_img;
    auto issue = [&](int j) {                                       // chunk j of this block -> stage j & 1
        const int n = n_begin + j / chunks_per_img, h0 = (j % chunks_per_img) * rows;
        float* sdy = smem + (j & 1) * stage_floats;
        float* sx = sdy + kWgO * kWgChunk;
        for (int e = tid; e < kWgO * nquads; e += 256) {
            const int o = e / nquads, q = e - o * nquads;
            const bool ok = o0 + o < CO;
            wg_cp16(sdy + o * kWgChunk + 4 * q, dy + (static_cast<size_t>(n) * CO + (ok ? o0 + o : 0)) * HW + h0 * W + 4 * q, ok);
        }
        for (int e = tid; e < kWgI * (rows + 2) * xq; e += 256) {
            const int i = e / ((rows + 2) * xq), r2 = e - i * (rows + 2) * xq, r = r2 / xq, g = r2 - r * xq;
            const int hh = h0 + r - 1;
            const bool ok = i0 + i < CI && hh >= 0 && hh < H && g >= 1 && g <= qpr;
            wg_cp16(sx + i * xplane + r * XS + 4 * g, x + (static_cast<size_t>(n) * CI + (ok ? i0 + i : 0)) * HW + (ok ? hh * W + 4 * (g - 1) : 0), ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (n_chunks > 0) issue(0);
    for (int j = 0; j < n_chunks; ++j) {
        if (j + 1 < n_chunks) { issue(j + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                           // chunk j has landed for every thread
        const float* sdy = smem + (j & 1) * stage_floats;
        const float* sx = sdy + kWgO * kWgChunk;
        for (int q = pl; q < nquads; q += 16) {
            const int r = q / qpr, c4 = (q - r * qpr) << 2;        // quad = pixels (r, c4 .. c4+3) of the chunk
            float4 g4[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) g4[a] = reinterpret_cast<const float4*>(sdy + (og * 4 + a) * kWgChunk)[q];
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const float* xp = sx + (ip * 2 + b) * xplane + r * XS + c4 + 3;      // -> pixel (r - 1, c4 - 1)
#pragma unroll
                for (int ty = 0; ty < 3; ++ty) {
                    const float* rp = xp + ty * XS;
                    const float4 m = *reinterpret_cast<const float4*>(rp + 1);
                    const float v[6] = {rp[0], m.x, m.y, m.z, m.w, rp[5]};
#pragma unroll
                    for (int tx = 0; tx < 3; ++tx)
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            float acc = s[a][b][ty * 3 + tx];
                            acc = fmaf(g4[a].x, v[tx], acc);
                            acc = fmaf(g4[a].y, v[tx + 1], acc);
                            acc = fmaf(g4[a].z, v[tx + 2], acc);
                            acc = fmaf(g4[a].w, v[tx + 3], acc);
                            s[a][b][ty * 3 + tx] = acc;
                        }
                }
            }
        }
        __syncthreads();                                           // stage j & 1 may be refilled (by issue(j + 2) in the next iteration)
    }
    float* dw = dw_part + static_cast<size_t>(blockIdx.z) * CO * CI * 9;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                float v = s[a][b][t];
#pragma unroll
                for (int off = 8; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                const int o = o0 + og * 4 + a, i = i0 + ip * 2 + b;
                if (pl == 0 && o < CO && i < CI) dw[(static_cast<size_t>(o) * CI + i) * 9 + t] = v;
            }
}
inline size_t wgrad_smem_bytes(int H, int W) {
    const int chunk = H * W < kWgChunk ? H * W : kWgChunk, rows = chunk / W;
    return 2 * sizeof(float) * (static_cast<size_t>(kWgO) * kWgChunk + static_cast<size_t>(kWgI) * (rows + 2) * (W + 8));   // two stages
}

__global__ void wgrad_sum_kernel(const float* __restrict__ part, float* __restrict__ dw, int n, int slices) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float v = 0.0f;
    for (int z = 0; z < slices; ++z) v += part[static_cast<size_t>(z) * n + e];
    dw[e] = v;
}

// ---------------------------------------------------------------- per-channel reductions over (n, h, w): one block per channel
// mode 0: sum(a)  (bias gradient);  mode 1: batch mean and 1/sqrt(var + eps) of a, running statistics updated;
// mode 2: sum(a) and sum(a * xhat) with xhat = (x - mean) * invstd  (batch-norm backward: dbeta, dgamma)
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) t += red[k];
    return t;
}
// Blocks of up to 1024 threads (kRedThreads at the launch sites); 16-byte loads when HW % 4 == 0.  visit(v) sees every element of
// channel c once, in a fixed per-thread order.
constexpr int kRedThreads = 1024;
template <typename F>
__device__ __forceinline__ void for_channel(const float* __restrict__ a, int c, int B, int C, int HW, F visit) {
    if ((HW & 3) == 0) {
        const int q = HW >> 2;
        for (int e = threadIdx.x; e < B * q; e += blockDim.x) {
            const int n = e / q, p = e - n * q;
            const size_t idx = (static_cast<size_t>(n) * C + c) * HW + 4 * p;
            const float4 v = __ldg(reinterpret_cast<const float4*>(a + idx));
            visit(idx, v.x); visit(idx + 1, v.y); visit(idx + 2, v.z); visit(idx + 3, v.w);
        }
    } else {
        for (int e = threadIdx.x; e < B * HW; e += blockDim.x) {
            const int n = e / HW, p = e - n * HW;
            const size_t idx = (static_cast<size_t>(n) * C + c) * HW + p;
            visit(idx, __ldg(a + idx));
        }
    }
}
__global__ void __launch_bounds__(kRedThreads)
channel_sum_kernel(const float* __restrict__ a, float* __restrict__ out, int B, int C, int HW) {
    __shared__ double red[32];
    const int c = blockIdx.x;
    double s = 0.0;
    for_channel(a, c, B, C, HW, [&](size_t, float v) { s += static_cast<double>(v); });
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[c] = static_cast<float>(s);
}
__global__ void __launch_bounds__(kRedThreads)
bn_stats_kernel(const float* __restrict__ x, float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ run_mean, float* __restrict__ run_var,
                int B, int C, int HW) {
    __shared__ double red[32];
    const int c = blockIdx.x;
    const double M = static_cast<double>(B) * HW;
    double s = 0.0;
    for_channel(x, c, B, C, HW, [&](size_t, float v) { s += static_cast<double>(v); });
    const double mu = block_sum(s, red) / M;
    __syncthreads();                                               // red[] is reused
    double q = 0.0;
    for_channel(x, c, B, C, HW, [&](size_t, float v) { const double dlt = static_cast<double>(v) - mu; q += dlt * dlt; });
    q = block_sum(q, red);
    if (threadIdx.x == 0) {
        const double var = q / M;
        mean[c] = static_cast<float>(mu);
        invstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kBnEps)));
        run_mean[c] = (1.0f - kBnMomentum) * run_mean[c] + kBnMomentum * static_cast<float>(mu);
        run_var[c] = (1.0f - kBnMomentum) * run_var[c] + kBnMomentum * static_cast<float>(M > 1.0 ? q / (M - 1.0) : var);
    }
}
__global__ void __launch_bounds__(kRedThreads)
bn_bwd_reduce_kernel(const float* __restrict__ dz, const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int C, int HW) {
    __shared__ double red[32];
    const int c = blockIdx.x;
    const float mu = mean[c], is = invstd[c];
    double s = 0.0, sx = 0.0;
    if ((HW & 3) == 0) {
        const int q = HW >> 2;
        for (int e = threadIdx.x; e < B * q; e += blockDim.x) {
            const int n = e / q, p = e - n * q;
            const size_t idx = (static_cast<size_t>(n) * C + c) * HW + 4 * p;
            const float4 g = __ldg(reinterpret_cast<const float4*>(dz + idx)), v = __ldg(reinterpret_cast<const float4*>(x + idx));
            s += static_cast<double>(g.x); sx += static_cast<double>(g.x * ((v.x - mu) * is));
            s += static_cast<double>(g.y); sx += static_cast<double>(g.y * ((v.y - mu) * is));
            s += static_cast<double>(g.z); sx += static_cast<double>(g.z * ((v.z - mu) * is));
            s += static_cast<double>(g.w); sx += static_cast<double>(g.w * ((v.w - mu) * is));
        }
    } else {
        for (int e = threadIdx.x; e < B * HW; e += blockDim.x) {
            const int n = e / HW, p = e - n * HW;
            const size_t idx = (static_cast<size_t>(n) * C + c) * HW + p;
            const float g = __ldg(dz + idx);
            s += static_cast<double>(g);
            sx += static_cast<double>(g * ((__ldg(x + idx) - mu) * is));
        }
    }
    s = block_sum(s, red);
    __syncthreads();
    sx = block_sum(sx, red);
    if (threadIdx.x == 0) { dbeta[c] = static_cast<float>(s); dgamma[c] = static_cast<float>(sx); }
}

// ---------------------------------------------------------------- elementwise passes (index e over [B][C][HW])
// forward: act = ELU(gamma * (x - mean) * invstd + beta); out = act * mask * scale (mask == nullptr: out = act).
// The mask is indexed per element, or per (n, c) when `spatial` (nn.SpatialDropout).
__global__ void bn_elu_drop_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const uint8_t* __restrict__ mask, int spatial, float scale, float* __restrict__ act,
                                   float* __restrict__ out, long long total, int C, int HW) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c = static_cast<int>((e / HW) % C);
    const float z = fmaf(gamma[c] * invstd[c], x[e] - mean[c], beta[c]);
    const float a = z > 0.0f ? z : expm1f(z);
    act[e] = a;
    if (out) out[e] = mask ? (mask[spatial ? e / HW : e] ? a * scale : 0.0f) : a;
}
// backward through dropout and ELU: dz = dout * mask * scale * (act > 0 ? 1 : act + 1)
__global__ void drop_elu_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ act, const uint8_t* __restrict__ mask, int spatial, float scale,
                                    float* __restrict__ dz, long long total, int HW) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= total) return;
    float g = dout[e];
    if (mask) g = mask[spatial ? e / HW : e] ? g * scale : 0.0f;
    const float a = act[e];
    dz[e] = a > 0.0f ? g : g * (a + 1.0f);
}
// batch-norm backward (input gradient): dx = gamma * invstd * (dz - dbeta / M - xhat * dgamma / M)
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const float* __restrict__ dgamma, const float* __restrict__ dbeta, float* __restrict__ dx,
                                    long long total, int C, int HW, float invM) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c = static_cast<int>((e / HW) % C);
    const float xh = (x[e] - mean[c]) * invstd[c];
    dx[e] = gamma[c] * invstd[c] * (dz[e] - dbeta[c] * invM - xh * dgamma[c] * invM);
}
// v1 dropout on the input image (the fixer's nn.Dropout(0.5, true)): y = x * mask
__global__ void mask_mul_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, float* __restrict__ y, long long total) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e < total) y[e] = mask[e] ? x[e] : 0.0f;
}
// nn.Dropout (v2) on its own (conv3's dropout sits behind the pooling): y = x * mask * scale; the backward pass is the same map
__global__ void drop_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, float scale, float* __restrict__ y, long long total) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e < total) y[e] = mask[e] ? x[e] * scale : 0.0f;
}
// 2x2 max pooling, stride 2: first maximum in window order (0,0),(0,1),(1,0),(1,1); `arg` keeps its position for the backward pass
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ arg, long long total_out, int H, int W) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= total_out) return;
    const int Wo = W / 2, Ho = H / 2;
    const int wo = static_cast<int>(e % Wo), ho = static_cast<int>((e / Wo) % Ho);
    const long long nc = e / (static_cast<long long>(Wo) * Ho);
    const float* p = x + nc * H * W + static_cast<long long>(2 * ho) * W + 2 * wo;
    float best = p[0];
    int a = 0;
    if (p[1] > best) { best = p[1]; a = 1; }
    if (p[W] > best) { best = p[W]; a = 2; }
    if (p[W + 1] > best) { best = p[W + 1]; a = 3; }
    y[e] = best;
    arg[e] = static_cast<uint8_t>(a);
}
__global__ void maxpool_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ arg, float* __restrict__ dx, long long total_out, int H, int W) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= total_out) return;
    const int Wo = W / 2, Ho = H / 2;
    const int wo = static_cast<int>(e % Wo), ho = static_cast<int>((e / Wo) % Ho);
    const long long nc = e / (static_cast<long long>(Wo) * Ho);
    float* p = dx + nc * H * W + static_cast<long long>(2 * ho) * W + 2 * wo;
    const int a = arg[e];
    const float g = dy[e];
    p[0] = a == 0 ? g : 0.0f; p[1] = a == 1 ? g : 0.0f; p[W] = a == 2 ? g : 0.0f; p[W + 1] = a == 3 ? g : 0.0f;
}

// ---------------------------------------------------------------- Linear: y[b][o] = bias[o] + sum_k x[b][k] w[o][k]
__global__ void __launch_bounds__(256)
linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y, int B, int K, int O) {
    const int gw = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= B * O) return;
    const int b = gw / O, o = gw - b * O;
    const float* xr = x + static_cast<size_t>(b) * K;
    const float* wr = w + static_cast<size_t>(o) * K;
    float s = 0.0f;
    for (int k = lane; k < K; k += 32) s = fmaf(__ldg(xr + k), __ldg(wr + k), s);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) y[gw] = s + __ldg(bias + o);
}
// dx[b][k] = sum_o dy[b][o] w[o][k].  Block = 64 columns k x 32 batch rows (blockIdx.y) x 4 slices of the o range: thread
// (k, slice) keeps the 32 rows' sums in registers, reads each weight ONCE (coalesced over k) and the dy values of a 32 x 32 chunk as
// 16-byte broadcasts from shared memory; the four slices are added in index order through shared memory.  (The first version -- a
// thread per (b, k) -- read the 16 MB weight matrix of the 8192 -> 512 layer once per batch row.)
constexpr int kLbK = 64, kLbS = 4, kLbOC = 32, kLbB = 32;
__global__ void __launch_bounds__(kLbK * kLbS)
linear_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int B, int K, int O) {
    __shared__ float4 sm4[kLbS * kLbB * kLbK / 4];                 // 32 KB: the dy chunks [slice][o][32 rows] (16 KB), then the slices' sums
    float* sm = reinterpret_cast<float*>(sm4);
    const int kk = threadIdx.x & (kLbK - 1), sl = threadIdx.x / kLbK;
    const int k = blockIdx.x * kLbK + kk, b0 = blockIdx.y * kLbB;
    const int per = (O + kLbS - 1) / kLbS, o_begin = sl * per, o_end = min(O, o_begin + per);
    float acc[kLbB];
#pragma unroll
    for (int r = 0; r < kLbB; ++r) acc[r] = 0.0f;
    float* sdy = sm + sl * kLbOC * kLbB;
    for (int oc = 0; oc < per; oc += kLbOC) {
        __syncthreads();
        for (int e = kk; e < kLbOC * kLbB; e += kLbK) {            // this slice's chunk: o fastest in the global reads
            const int oo = e & (kLbOC - 1), bb = e / kLbOC, o = o_begin + oc + oo;
            sdy[oo * kLbB + bb] = (o < o_end && b0 + bb < B) ? __ldg(dy + static_cast<size_t>(b0 + bb) * O + o) : 0.0f;
        }
        __syncthreads();
        for (int oo = 0; oo < kLbOC; ++oo) {
            const int o = o_begin + oc + oo;
            if (o >= o_end) break;
            const float wv = k < K ? __ldg(w + static_cast<size_t>(o) * K + k) : 0.0f;
            const float4* row = reinterpret_cast<const float4*>(sdy + oo * kLbB);
#pragma unroll
            for (int r4 = 0; r4 < kLbB / 4; ++r4) {
                const float4 g = row[r4];
                acc[4 * r4 + 0] = fmaf(g.x, wv, acc[4 * r4 + 0]); acc[4 * r4 + 1] = fmaf(g.y, wv, acc[4 * r4 + 1]);
                acc[4 * r4 + 2] = fmaf(g.z, wv, acc[4 * r4 + 2]); acc[4 * r4 + 3] = fmaf(g.w, wv, acc[4 * r4 + 3]);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kLbB; ++r) sm[(sl * kLbB + r) * kLbK + kk] = acc[r];
    __syncthreads();
    for (int r = sl; r < kLbB; r += kLbS) {                         // slice sl finishes rows sl, sl + 4, ...
        float v = 0.0f;
#pragma unroll
        for (int z = 0; z < kLbS; ++z) v += sm[(z * kLbB + r) * kLbK + kk];
        if (k < K && b0 + r < B) dx[static_cast<size_t>(b0 + r) * K + k] = v;
    }
}
// dw[o][k] = sum_b dy[b][o] x[b][k];  db[o] = sum_b dy[b][o]  (k == 0 threads)
__global__ void linear_bwd_w_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw, float* __restrict__ db, int B, int K, int O) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<long long>(O) * K) return;
    const int o = static_cast<int>(e / K), k = static_cast<int>(e - static_cast<long long>(o) * K);
    float s = 0.0f, sb = 0.0f;
    for (int b = 0; b < B; ++b) { const float g = __ldg(dy + static_cast<size_t>(b) * O + o); s = fmaf(g, __ldg(x + static_cast<size_t>(b) * K + k), s); sb += g; }
    dw[e] = s;
    if (k == 0) db[o] = sb;
}
__global__ void tanh_fwd_kernel(float* __restrict__ y, long long total) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e < total) y[e] = tanhf(y[e]);
}

// ---------------------------------------------------------------- criterion, penalties, Adam
// nn.MSECriterion: loss = mean((pred - target)^2), dpred = 2 (pred - target) / n (and through Tanh when the net ends in one);
// one block, fixed-order reduction in double.
__global__ void __launch_bounds__(256)
mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, float* __restrict__ dpred, double* __restrict__ loss, int n, int tanh_out) {
    __shared__ double red[8];
    double s = 0.0;
    for (int e = threadIdx.x; e < n; e += 256) {
        const float df = pred[e] - target[e];
        s += static_cast<double>(df) * static_cast<double>(df);
        float g = 2.0f * df / static_cast<float>(n);
        if (tanh_out) g *= 1.0f - pred[e] * pred[e];
        dpred[e] = g;
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) loss[0] = s / static_cast<double>(n);
}
// L1 / L2 penalty sums over the parameters (loss value only): partial[block] = sum |p|, partial2[block] = sum p^2
__global__ void __launch_bounds__(256)
penalty_kernel(const float* __restrict__ p, const uint8_t* __restrict__ is_param, long long n, double* __restrict__ partial) {
    __shared__ double red[8];
    double a = 0.0, q = 0.0;
    for (long long e = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; e < n; e += static_cast<long long>(gridDim.x) * 256)
        if (is_param[e]) { const double v = p[e]; a += fabs(v); q += v * v; }
    a = block_sum(a, red);
    q = block_sum(q, red);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = a; partial[2 * blockIdx.x + 1] = q; }
}
struct AdamHyper { float lr, beta1, beta2, eps, l1, l2, clamp, step_size; };
// grad += l1 sign(p) + l2 p; clamp; Adam moments and update (train_r.lua:150-166, optim.adam)
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, const uint8_t* __restrict__ is_param,
                            long long n, AdamHyper h) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n || !is_param[e]) return;
    const float x = p[e];
    float gr = g[e];
    if (h.l1 != 0.0f || h.l2 != 0.0f) gr += h.l1 * (x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f)) + h.l2 * x;
    if (h.clamp != 0.0f) gr = fminf(fmaxf(gr, -h.clamp), h.clamp);
    g[e] = gr;
    const float mm = h.beta1 * m[e] + (1.0f - h.beta1) * gr;
    const float vv = h.beta2 * v[e] + (1.0f - h.beta2) * gr * gr;
    m[e] = mm; v[e] = vv;
    p[e] = x - h.step_size * mm / (sqrtf(vv) + h.eps);
}

}  // namespace trn
}  // namespace ganrev
