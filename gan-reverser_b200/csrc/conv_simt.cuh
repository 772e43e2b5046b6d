// conv_simt.cuh -- plain CUDA-core statement of the same ConvGemm layer (one thread per
// output element, indices derived from the OUTPUT side).  It is not a fallback: it exists
// so the tcgen05 kernel can be A/B-checked on the device layer by layer
// (ganrev_set_option("conv_impl", 1)) and is far too slow for production.
#pragma once
#include "common.cuh"

namespace ganrev {

__device__ __forceinline__ float simt_conv_at(const ConvGemm& p, int n, int h, int w, int phase, int co) {
    const int Ktot = p.ngroups * p.ndy * p.Cin;
    const bf16* wrow = p.B + (static_cast<size_t>(phase) * p.cout_pad + co) * Ktot;
    float acc = 0.0f;
    for (int g = 0; g < p.ngroups; ++g)
        for (int j = 0; j < p.ndy; ++j) {
            const int hh = h + p.gdy0[phase][g] + j, ww = w + p.gdx[phase][g];
            if (hh < 0 || hh >= p.Hin || ww < 0 || ww >= p.Win) continue;   // zero padding
            const bf16* a = p.A + ((static_cast<size_t>(n) * p.Hin + hh) * p.Win + ww) * p.Cin;
            const bf16* wt = wrow + static_cast<size_t>(g * p.ndy + j) * p.Cin;
            for (int ci = 0; ci < p.Cin; ++ci) acc = fmaf(__bfloat162float(a[ci]), __bfloat162float(wt[ci]), acc);
        }
    return acc + p.shift[co];   // BN scale is folded into the weights
}

__global__ void conv_simt_kernel(const ConvGemm p, const long long total) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int co = static_cast<int>(idx % p.cout_real);
    long long t = idx / p.cout_real;
    const int ow = static_cast<int>(t % p.Wout); t /= p.Wout;
    const int oh = static_cast<int>(t % p.Hout); t /= p.Hout;
    const int n = static_cast<int>(t);
    float v;
    if (p.pool) {
        v = simt_conv_at(p, n, 2 * oh, 2 * ow, 0, co);
        v = fmaxf(v, simt_conv_at(p, n, 2 * oh, 2 * ow + 1, 0, co));
        v = fmaxf(v, simt_conv_at(p, n, 2 * oh + 1, 2 * ow, 0, co));
        v = fmaxf(v, simt_conv_at(p, n, 2 * oh + 1, 2 * ow + 1, 0, co));
    } else if (p.up == 2) {
        v = simt_conv_at(p, n, oh >> 1, ow >> 1, (oh & 1) * 2 + (ow & 1), co);
    } else {
        v = simt_conv_at(p, n, oh, ow, 0, co);
    }
    v = apply_act(v, p.act);
    const size_t off = static_cast<size_t>(n) * p.out_sN + (static_cast<size_t>(oh) * p.Wout + ow) * p.out_sP + static_cast<size_t>(co) * p.out_sC;
    if (p.out_fp32) reinterpret_cast<float*>(p.out)[off] = v;
    else reinterpret_cast<bf16*>(p.out)[off] = __float2bfloat16_rn(v);
}

}  // namespace ganrev
