// How fast does the register-resident biquad cascade (2 sections, DF2T, float64) run on B200?
// cycles per 23-sample chunk for 1..16 warps per SM, and with the chunk split into 2 independent halves (ILP 2).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double step(const double (&c)[2][6], double (&z)[2][2], double x) {
#pragma unroll
    for (int s = 0; s < 2; s++) {
        double y = __fma_rn(c[s][0], x, z[s][0]);
        z[s][0] = __fma_rn(-c[s][4], y, __fma_rn(c[s][1], x, z[s][1]));
        z[s][1] = __fma_rn(-c[s][5], y, c[s][2] * x);
        x = y;
    }
    return x;
}
template <int ILP>
__global__ void k(double* out, long long* cyc, int reps, const double* cin) {
    double c[2][6];
    for (int s = 0; s < 2; s++) for (int j = 0; j < 6; j++) c[s][j] = cin[s * 6 + j];
    double y[24];
    for (int j = 0; j < 24; j++) y[j] = 1e-3 * (threadIdx.x + j);
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
        if (ILP == 1) {
            double z[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
            for (int j = 0; j < 23; j++) y[j] = step(c, z, y[j]);
        } else {
            double za[2][2] = {{0, 0}, {0, 0}}, zb[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
            for (int j = 0; j < 12; j++) { y[j] = step(c, za, y[j]); y[12 + j] = step(c, zb, y[12 + j]); }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = (t1 - t0) / reps;
    double s = 0; for (int j = 0; j < 24; j++) s += y[j];
    if (s == 1.2345) out[0] = s;
}
int main() {
    double h[12] = {0.7726678130550604, -1.5453356261101208, 0.7726678130550604, 1, -1.6609362580597045, 0.6937015023502847,
                    1, -2, 1, 1, -1.8246287238899792, 0.8606231249922253};
    double *d, *c; long long* cy; cudaMalloc(&d, 64); cudaMalloc(&c, 96); cudaMalloc(&cy, 8);
    cudaMemcpy(c, h, 96, cudaMemcpyHostToDevice);
    for (int warps : {1, 4, 8, 16, 32}) {
        long long a, b;
        k<1><<<148, warps * 32>>>(d, cy, 2000, c); cudaDeviceSynchronize(); cudaMemcpy(&a, cy, 8, cudaMemcpyDeviceToHost);
        k<2><<<148, warps * 32>>>(d, cy, 2000, c); cudaDeviceSynchronize(); cudaMemcpy(&b, cy, 8, cudaMemcpyDeviceToHost);
        printf("warps/SM %2d: 23-sample chunk %5lld cycles (ILP1), 24-sample 2x12 %5lld cycles (ILP2); pipe-bound %d cycles\n", warps, a, b, warps * 230 * 2 / 4);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
