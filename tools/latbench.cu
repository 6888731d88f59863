// Dependent-issue latencies on this GPU (one warp): DFMA, DMUL, double max (DSETP+FSEL), 64-bit shuffle, LDS.128, rcp seed.
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double* out, long long* t, double a, double b) {
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = a;
    __syncthreads();
    double x = a + threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, b, a);
    long long t1 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x * b;
    long long t2 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = (x > a) ? x * 1.0000001 : a;   // max + something dependent
    long long t3 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1);
    long long t4 = clock64();
    int idx = threadIdx.x & 31;
#pragma unroll 16
    for (int i = 0; i < N; ++i) { double2 v = *reinterpret_cast<double2*>(&sm[(idx * 2) & 62]); idx = (int)(v.x) & 31; x += v.y; }
    long long t5 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
    long long t6 = clock64();
    float f = (float)x;
#pragma unroll 16
    for (int i = 0; i < N; ++i) f = fmaf(f, 1.0001f, 0.5f);
    long long t7 = clock64();
    out[threadIdx.x] = x + f;
    if (threadIdx.x == 0) { t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; t[3] = t4 - t3; t[4] = t5 - t4; t[5] = t6 - t5; t[6] = t7 - t6; }
}
int main() {
    double* out; long long* t; long long h[7];
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&t, 64);
    for (int warps = 1; warps <= 8; warps *= 2) {
        k<<<1, 32 * warps>>>(out, t, 1.0, 0.999);
        cudaMemcpy(h, t, 56, cudaMemcpyDeviceToHost);
        printf("warps=%d  DFMA %.1f  DMUL %.1f  dmax+mul %.1f  SHFL64 %.1f  LDS.128+cvt+add %.1f  RCP64H %.1f  FFMA %.1f  (cycles per dependent op)\n", warps,
               h[0] / (double)N, h[1] / (double)N, h[2] / (double)N, h[3] / (double)N, h[4] / (double)N, h[5] / (double)N, h[6] / (double)N);
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
