// Memory ceilings for the pruning kernel's access pattern on this GPU: 256-bit stores only,
// 256-bit load+store copy, and 2-load+1-store (the partial x partial pattern).  nvcc -O3 -arch=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ld4(const double* p, double (&v)[4]) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st4(double* p, const double (&v)[4]) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
template <int MODE, int R>
__global__ void __launch_bounds__(256) k(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ c, size_t bufStride, int nbuf) {
    const size_t base = (size_t)blockIdx.x * (256 * R) * 4 + threadIdx.x * 4;
    for (int j = 0; j < nbuf; ++j) {
        double v[R][4], w[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) { v[r][0] = j; v[r][1] = r; v[r][2] = 1; v[r][3] = 2; }
        if (MODE >= 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) ld4(a + j * bufStride + base + r * 1024, v[r]);
        }
        if (MODE >= 2) {
#pragma unroll
            for (int r = 0; r < R; ++r) { ld4(b + j * bufStride + base + r * 1024, w[r]); v[r][0] += w[r][0]; v[r][1] += w[r][1]; v[r][2] += w[r][2]; v[r][3] += w[r][3]; }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) st4(c + j * bufStride + base + r * 1024, v[r]);
    }
}
template <int MODE, int R> void run(const char* name, double* a, double* b, double* c, size_t elemsPerBuf, int nbuf) {
    int blocks = (int)(elemsPerBuf / (256 * R * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 2; ++it) k<MODE, R><<<blocks, 256>>>(a, b, c, elemsPerBuf, nbuf);
    cudaEventRecord(e0);
    for (int it = 0; it < 3; ++it) k<MODE, R><<<blocks, 256>>>(a, b, c, elemsPerBuf, nbuf);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    double bytes = (double)elemsPerBuf * 8 * nbuf * (1 + MODE);
    printf("%-28s R=%d blocks=%d  %.3f ms  %.1f GB/s\n", name, R, blocks, ms, bytes / ms / 1e6);
}
int main() {
    const size_t elems = (size_t)1 << 24;      // 128 MB per buffer (1M patterns x 4 cats x 4)
    const int nbuf = 24;                        // 3 GB per array
    double *a, *b, *c;
    cudaMalloc(&a, elems * 8 * nbuf); cudaMalloc(&b, elems * 8 * nbuf); cudaMalloc(&c, elems * 8 * nbuf);
    cudaMemset(a, 0, elems * 8 * nbuf); cudaMemset(b, 0, elems * 8 * nbuf);
    run<0, 1>("store only", a, b, c, elems, nbuf); run<0, 4>("store only", a, b, c, elems, nbuf);
    run<1, 1>("load+store", a, b, c, elems, nbuf); run<1, 4>("load+store", a, b, c, elems, nbuf);
    run<2, 1>("2 loads+store", a, b, c, elems, nbuf); run<2, 4>("2 loads+store", a, b, c, elems, nbuf);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
