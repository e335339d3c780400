// Dependent-issue latency of the instructions on the wavefront critical path (one warp).
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int K> __global__ void lat(int* out, long long* cyc, int a, int b) {
    __shared__ int sm[64];
    sm[threadIdx.x & 63] = threadIdx.x;
    __syncthreads();
    int x = threadIdx.x + a, y = b;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (K == 0) x = __viaddmin_s32(x, a, y);                       // VIADDMNMX chain (operand a)
        if (K == 1) x = __viaddmin_s32(y, a, x);                       // chain through 3rd operand
        if (K == 2) x = min(x + a, y);                                  // compiler's choice
        if (K == 3) x = x + y;                                          // IADD3
        if (K == 4) x = min(x, y + i);                                  // VIMNMX
        if (K == 5) x = __shfl_up_sync(0xffffffffu, x, 1);              // SHFL.UP
        if (K == 6) x = sm[x & 63];                                     // LDS dependent
        if (K == 7) x = __viaddmin_s16x2((unsigned)x, (unsigned)a, (unsigned)y);
        if (K == 8) { x = __shfl_up_sync(0xffffffffu, x, 1); x = (threadIdx.x == 0) ? a : x; x = __viaddmin_s32(x, a, y); }
        if (K == 9) { int p = (x != i) ? a : 0; x = x + p; }          // ISETP+SEL+IADD
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[K] = t1 - t0;
    out[threadIdx.x] = x;
}
int main() {
    int* out; long long* cyc; cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 128);
    const char* names[] = {"VIADDMNMX(a)", "VIADDMNMX(c)", "min(x+a,y)", "IADD3", "VIMNMX", "SHFL.UP", "LDS", "VIADDMNMX.S16x2", "SHFL+SEL+VIADDMNMX", "ISETP+SEL+IADD"};
#define RUN(K) lat<K><<<1, 32>>>(out, cyc, 3, 1000000); cudaDeviceSynchronize(); lat<K><<<1, 32>>>(out, cyc, 3, 1000000); cudaDeviceSynchronize(); printf("%-22s %.2f cycles/iter\n", names[K], (double)cyc[K] / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    return 0;
}
