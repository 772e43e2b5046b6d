// Throughput of the float -> int64 fixed-point conversion variants used by the kmeans centroid sums (per SM, per clock).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o f2i_bench f2i_bench.cu && ./f2i_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ long long fix_f2i(float x, float sc) { return __float2ll_rn(__fmul_rn(x, sc)); }
// integer-only: |x| * 2^shift as mant << (e - 150 + shift); falls back to F2I when bits would be shifted out
__device__ __forceinline__ long long fix_alu(float x, int shift, float sc) {
    const unsigned u = __float_as_uint(x);
    const int e = (u >> 23) & 0xff;
    const int sh = e - 150 + shift;
    if (sh < 0 || e == 0) return __float2ll_rn(__fmul_rn(x, sc));
    const unsigned long long m = static_cast<unsigned long long>((u & 0x7fffffu) | 0x800000u) << sh;
    const long long s = static_cast<int>(u) >> 31;
    return (static_cast<long long>(m) ^ s) - s;
}
// two 32-bit conversions: hi = rint(v * 2^-24) (|hi| < 2^31 when |v| < 2^55), lo = v - hi * 2^24 (exact), result = hi * 2^24 + rint(lo)
__device__ __forceinline__ long long fix_two32(float x, float sc) {
    const float v = __fmul_rn(x, sc);
    const int hi = __float2int_rn(v * 5.9604644775390625e-8f);
    const float lo = __fmaf_rn(static_cast<float>(hi), -16777216.0f, v);
    return (static_cast<long long>(hi) << 24) + __float2int_rn(lo);
}
template <int V>
__global__ void k(const float* __restrict__ in, long long* out, int iters, float sc, int shift) {
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = in[threadIdx.x + 32 * i];
    long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (V == 0) acc[i] += fix_f2i(x[i], sc);
            if (V == 1) acc[i] += fix_alu(x[i], shift, sc);
            if (V == 2) acc[i] += fix_two32(x[i], sc);
            if (V == 3) acc[i] += __float2int_rn(x[i] * sc * 1e-9f);
            x[i] = __fmaf_rn(x[i], 1.00001f, 1.0e-4f);
        }
    }
    long long s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* in; long long* out;
    cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 1024 * 8);
    float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = (float)(i % 97) * 0.013f - 0.6f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const int iters = 20000; const int shift = 37; const float sc = 137438953472.0f;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int threads : {64, 256, 1024}) {
        for (int v = 0; v < 4; ++v) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (v == 0) k<0><<<148, threads>>>(in, out, iters, sc, shift);
                if (v == 1) k<1><<<148, threads>>>(in, out, iters, sc, shift);
                if (v == 2) k<2><<<148, threads>>>(in, out, iters, sc, shift);
                if (v == 3) k<3><<<148, threads>>>(in, out, iters, sc, shift);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double conv = (double)iters * 8 * threads;   // per SM
            printf("threads/SM %4d variant %d (%s): %.3f ms, %.2f conversions per clock per SM at %.0f MHz nominal\n", threads, v,
                   v == 0 ? "FMUL+F2I.S64" : v == 1 ? "integer shift" : v == 2 ? "two F2I.S32" : "F2I.S32 only", ms, conv / (ms * 1e-3 * clk * 1e3), clk / 1e3);
        }
    }
    long long hsum; cudaMemcpy(&hsum, out, 8, cudaMemcpyDeviceToHost);
    printf("%lld %s\n", hsum, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
