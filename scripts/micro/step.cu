// One warp running the wavefront step loop in isolation: where do ~450 cycles/step go?
#include <cstdio>
#include <cuda_runtime.h>
template <int C, int VAR>
__global__ void stepk(int* out, long long* cyc, int steps, int insc, int delc, int subc) {
    __shared__ int hyp[4096];
    __shared__ int bnd[4096];
    const int lane = threadIdx.x;
    for (int i = lane; i < 4096; i += 32) { hyp[i] = (i * 7) & 63; bnd[i] = i; }
    __syncwarp();
    int v[C], rt[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { v[c] = lane * C + c; rt[c] = (lane * 13 + c * 5) & 63; }
    int pl = 1 << 29;
    long long t0 = clock64();
    for (int s = 1; s <= steps; ++s) {
        int hand = __shfl_up_sync(0xffffffffu, v[C - 1], 1);
        if (VAR >= 1 && lane == 0) hand = 1 << 29;
        const int diag = pl; pl = hand;
        const int i = s - lane;
        bool active = true;
        if (VAR >= 2) active = (unsigned)(i - 1) < (unsigned)steps;
        if (active) {
            const int ht = (VAR >= 3) ? hyp[(s - lane) & 4095] : (s & 63);
            int dg = diag, lf = hand;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int up = v[c];
                int sb = dg;
                if (rt[c] != ht) sb += subc;
                const int t = __viaddmin_s32(up, insc, sb);
                lf = __viaddmin_s32(lf, delc, t);
                dg = up; v[c] = lf;
            }
            if (VAR >= 4 && lane == 31) bnd[s & 4095] = v[C - 1];
            if (VAR >= 5 && lane == 31) out[64 + (s & 1023) * 32] = v[C - 1];
        }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    int acc = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) acc ^= v[c];
    out[lane] = acc + bnd[lane];
}
int main() {
    int* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 128);
    const int steps = 2000;
#define RUN(C, VAR) stepk<C, VAR><<<1, 32>>>(out, cyc, steps, 3, 3, 4); cudaDeviceSynchronize(); stepk<C, VAR><<<1, 32>>>(out, cyc, steps, 3, 3, 4); cudaDeviceSynchronize(); printf("C=%2d var=%d  %.1f cycles/step\n", C, VAR, (double)cyc[0] / steps);
    RUN(4, 0) RUN(4, 1) RUN(4, 2) RUN(4, 3) RUN(4, 4) RUN(4, 5)
    RUN(16, 0) RUN(16, 1) RUN(16, 2) RUN(16, 3) RUN(16, 4) RUN(16, 5)
    return 0;
}
