// Throughput of float32<->float64 conversions on B200 versus integer bit-manipulation equivalents.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ double f2d_bits(float f) {   // exact for +0 and positive normal floats
    const uint32_t u = __float_as_uint(f);
    const uint32_t hi = u ? (u >> 3) + 0x38000000u : 0u;
    return __hiloint2double((int)hi, (int)(u << 29));
}
__device__ __forceinline__ float d2f_bits(double d) {   // round-to-nearest-even, +0 and doubles inside float32's normal range
    const uint32_t hi = (uint32_t)__double2hiint(d), lo = (uint32_t)__double2loint(d);
    uint32_t w = ((hi - 0x38000000u) << 3) | (lo >> 29);
    const uint32_t rest = lo << 3;
    w += (rest > 0x80000000u) || (rest == 0x80000000u && (w & 1u));
    return __uint_as_float(hi ? w : 0u);
}
template <int MODE>
__global__ void tput(float* sink, int iters, float a) {
    float x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double d0 = x0, d1 = x1, d2 = x2, d3 = x3;
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) { d0 = (double)x0 + d0; d1 = (double)x1 + d1; d2 = (double)x2 + d2; d3 = (double)x3 + d3; x0 += 1.f; x1 += 1.f; x2 += 1.f; x3 += 1.f; }
        if (MODE == 1) { x0 = __double2float_rn(d0) + x0; x1 = __double2float_rn(d1) + x1; x2 = __double2float_rn(d2) + x2; x3 = __double2float_rn(d3) + x3; d0 += 1.0; d1 += 1.0; d2 += 1.0; d3 += 1.0; }
        if (MODE == 2) { d0 = f2d_bits(x0) + d0; d1 = f2d_bits(x1) + d1; d2 = f2d_bits(x2) + d2; d3 = f2d_bits(x3) + d3; x0 += 1.f; x1 += 1.f; x2 += 1.f; x3 += 1.f; }
        if (MODE == 3) { x0 = d2f_bits(d0) + x0; x1 = d2f_bits(d1) + x1; x2 = d2f_bits(d2) + x2; x3 = d2f_bits(d3) + x3; d0 += 1.0; d1 += 1.0; d2 += 1.0; d3 += 1.0; }
        if (MODE == 4) { d0 += 1.0; d1 += 1.0; d2 += 1.0; d3 += 1.0; x0 += 1.f; x1 += 1.f; x2 += 1.f; x3 += 1.f; }   // baseline: the adds alone
    }
    if (x0 + x1 + x2 + x3 + (float)(d0 + d1 + d2 + d3) == 1.2345f) sink[0] = x0;
}
__global__ void check(int* bad) {
    // exhaustive over positive normal floats (stride 7) and a sweep of doubles
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t u = 0x00800000u + i * 7u; u < 0x7f800000u; u += gridDim.x * blockDim.x * 7u) {
        const float f = __uint_as_float(u);
        if (f2d_bits(f) != (double)f) atomicAdd(bad, 1);
        const double d = (double)f * (1.0 + (double)(u & 1023u) * 1.1102230246251565e-13) ;
        if (d < 3.4e38 && d2f_bits(d) != __double2float_rn(d)) atomicAdd(bad + 1, 1);
        const double e = (double)f + (double)__uint_as_float(u + 1) ;   // exact ties and near-ties
        const double h = e * 0.5;
        if (h < 3.4e38 && d2f_bits(h) != __double2float_rn(h)) atomicAdd(bad + 2, 1);
    }
    if (i == 0) { if (f2d_bits(0.0f) != 0.0) atomicAdd(bad, 1); if (d2f_bits(0.0) != 0.0f) atomicAdd(bad + 1, 1); }
}
template <int MODE>
static void run(const char* name, float* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tput<MODE><<<148 * 4, 256>>>(d, 1000, 1.0f); cudaDeviceSynchronize();
    cudaEventRecord(e0); tput<MODE><<<148 * 4, 256>>>(d, 20000, 1.0f); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 148.0 * 4 * 256 * 4 * 20000;
    printf("%-28s %.3f ms  -> %.1f conversions(+adds)/clk/SM\n", name, ms, ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float* d; cudaMalloc(&d, 64);
    run<4>("baseline dadd+fadd", d);
    run<0>("F2F.F64.F32 (+dadd,fadd)", d);
    run<1>("F2F.F32.F64 (+fadd,dadd)", d);
    run<2>("f2d bit ops (+dadd,fadd)", d);
    run<3>("d2f bit ops (+fadd,dadd)", d);
    int* bad; cudaMalloc(&bad, 16); cudaMemset(bad, 0, 16);
    check<<<148 * 8, 256>>>(bad); cudaDeviceSynchronize();
    int hb[4]; cudaMemcpy(hb, bad, 16, cudaMemcpyDeviceToHost);
    printf("mismatches: f2d %d, d2f %d, d2f ties %d  (%s)\n", hb[0], hb[1], hb[2], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
