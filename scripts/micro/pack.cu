// Micro-benchmark behind the K0 (pack) design: int64 [T][N] -> int32 [N][Tp] (+ u16 [N][Tp16]).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pack scripts/micro/pack.cu && /tmp/pack
// Variants: lane-owns-row direct stores (NT=8/16/32), loads only, stores only, smem-staged
// coalesced stores, and a plain streaming copy of the same byte counts as the upper bound.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

struct Args {
    const long long* tok;
    int T, N;
    int* p32;
    int Tp;
    unsigned short* p16;
    int Tp16;
    int* sink;
};

template <int NT, int MODE>  // MODE 0 full, 1 loads only, 2 stores only, 3 no u16
__global__ void __launch_bounds__(128) direct_kernel(Args a) {
    const int lane = threadIdx.x & 31;
    const int n = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32 + lane;
    if (n >= a.N) return;
    const long long* src = a.tok + n;
    int* row32 = a.p32 + (size_t)n * a.Tp;
    unsigned short* row16 = a.p16 + (size_t)n * a.Tp16;
    int acc = 0;
    for (int t0 = 0; t0 + NT <= a.T; t0 += NT, src += (size_t)NT * a.N) {
        int v[NT];
        if (MODE != 2) {
            long long raw[NT];
#pragma unroll
            for (int u = 0; u < NT; ++u) raw[u] = __ldcs(src + (size_t)u * a.N);
#pragma unroll
            for (int u = 0; u < NT; ++u) { v[u] = (int)raw[u]; acc ^= v[u]; }
        } else {
#pragma unroll
            for (int u = 0; u < NT; ++u) v[u] = n + t0 + u;
        }
        if (MODE != 1) {
#pragma unroll
            for (int c = 0; c < NT / 4; ++c)
                *reinterpret_cast<int4*>(row32 + t0 + 4 * c) = make_int4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            if (MODE != 3) {
#pragma unroll
                for (int c = 0; c < NT / 8; ++c) {
                    const int* w = v + 8 * c;
                    *reinterpret_cast<uint4*>(row16 + t0 + 8 * c) =
                        make_uint4((w[0] & 0xffff) | (w[1] << 16), (w[2] & 0xffff) | (w[3] << 16),
                                   (w[4] & 0xffff) | (w[5] << 16), (w[6] & 0xffff) | (w[7] << 16));
                }
            }
        }
    }
    if (acc == 0x12345678) a.sink[0] = acc;
}

// smem-staged: warp owns 32 sequences x 32 positions, stores leave as full 128-byte lines
template <int MODE>  // 0 full, 3 no u16
__global__ void __launch_bounds__(128) staged_kernel(Args a) {
    __shared__ __align__(16) int tile[4][32][36];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n0 = (blockIdx.x * 4 + w) * 32;
    if (n0 >= a.N) return;
    const int n = n0 + lane;
    const long long* src = a.tok + n;
    int acc = 0;
    for (int t0 = 0; t0 + 32 <= a.T; t0 += 32, src += (size_t)32 * a.N) {
        long long raw[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) raw[u] = __ldcs(src + (size_t)u * a.N);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            *reinterpret_cast<int4*>(&tile[w][lane][4 * c]) =
                make_int4((int)raw[4 * c], (int)raw[4 * c + 1], (int)raw[4 * c + 2], (int)raw[4 * c + 3]);
            acc ^= (int)raw[4 * c];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3), c = lane & 7;
            const int4 x = *reinterpret_cast<const int4*>(&tile[w][r][4 * c]);
            *reinterpret_cast<int4*>(a.p32 + (size_t)(n0 + r) * a.Tp + t0 + 4 * c) = x;
        }
        if (MODE != 3) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = i * 8 + (lane >> 2), c = lane & 3;
                const int4 x = *reinterpret_cast<const int4*>(&tile[w][r][8 * c]);
                const int4 y = *reinterpret_cast<const int4*>(&tile[w][r][8 * c + 4]);
                *reinterpret_cast<uint4*>(a.p16 + (size_t)(n0 + r) * a.Tp16 + t0 + 8 * c) =
                    make_uint4((x.x & 0xffff) | (x.y << 16), (x.z & 0xffff) | (x.w << 16),
                               (y.x & 0xffff) | (y.y << 16), (y.z & 0xffff) | (y.w << 16));
            }
        }
        __syncwarp();
    }
    if (acc == 0x12345678) a.sink[0] = acc;
}

// CTA-staged: a CTA of 4 warps owns 32 sequences; warp w loads positions [tb+32w, tb+32w+32)
// of a 128-position chunk into a shared [32][132] tile, then whole rows leave as contiguous
// 16-byte vectors (32 adjacent rows = one contiguous block of the output).
template <int MODE, int NB>
__global__ void __launch_bounds__(128) cta_kernel(Args a) {
    __shared__ __align__(16) int tile[32][132];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n0 = blockIdx.x * 32;
    const int n = n0 + lane;
    int acc = 0;
    for (int tb = 0; tb < a.T; tb += 128) {
        const int t0 = tb + 32 * w;
        const long long* src = a.tok + (size_t)t0 * a.N + n;
#pragma unroll
        for (int h = 0; h < 32 / NB; ++h) {
            long long raw[NB];
#pragma unroll
            for (int u = 0; u < NB; ++u) raw[u] = (t0 + h * NB + u < a.T) ? __ldcs(src + (size_t)(h * NB + u) * a.N) : 0;
#pragma unroll
            for (int c = 0; c < NB / 4; ++c) {
                *reinterpret_cast<int4*>(&tile[lane][32 * w + h * NB + 4 * c]) =
                    make_int4((int)raw[4 * c], (int)raw[4 * c + 1], (int)raw[4 * c + 2], (int)raw[4 * c + 3]);
                acc ^= (int)raw[4 * c];
            }
        }
        __syncthreads();
        const int width = min(128, a.Tp - tb);  // multiple of 4
        const int width16 = min(128, a.Tp16 - tb);  // multiple of 8
        for (int r = w; r < 32; r += 4) {
            for (int c = lane; c < width / 4; c += 32)
                *reinterpret_cast<int4*>(a.p32 + (size_t)(n0 + r) * a.Tp + tb + 4 * c) =
                    *reinterpret_cast<const int4*>(&tile[r][4 * c]);
            if (MODE != 3)
                for (int c = lane; c < width16 / 8; c += 32) {
                    const int4 x = *reinterpret_cast<const int4*>(&tile[r][8 * c]);
                    const int4 y = *reinterpret_cast<const int4*>(&tile[r][8 * c + 4]);
                    *reinterpret_cast<uint4*>(a.p16 + (size_t)(n0 + r) * a.Tp16 + tb + 8 * c) =
                        make_uint4((x.x & 0xffff) | (x.y << 16), (x.z & 0xffff) | (x.w << 16),
                                   (y.x & 0xffff) | (y.y << 16), (y.z & 0xffff) | (y.w << 16));
                }
        }
        __syncthreads();
    }
    if (acc == 0x12345678) a.sink[0] = acc;
}

__global__ void copy_kernel(const int4* __restrict__ in, size_t nin, int4* __restrict__ out, size_t nout) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int4 acc = make_int4(0, 0, 0, 0);
    for (size_t i = i0; i < nin; i += stride) { const int4 x = __ldcs(in + i); acc.x ^= x.x; acc.y ^= x.y; }
    for (size_t i = i0; i < nout; i += stride) out[i] = acc;
}

int main() {
    const int T = 101, N = 131072, Tp = 104, Tp16 = 104, NBUF = 2;
    Args a[NBUF];
    for (int b = 0; b < NBUF; ++b) {
        long long* tok;
        CK(cudaMalloc(&tok, sizeof(long long) * (size_t)T * N));
        std::vector<long long> h((size_t)T * N);
        for (size_t i = 0; i < h.size(); ++i) h[i] = (long long)((i * 2654435761u) % 10000u);
        CK(cudaMemcpy(tok, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
        a[b].tok = tok; a[b].T = T; a[b].N = N; a[b].Tp = Tp; a[b].Tp16 = Tp16;
        CK(cudaMalloc(&a[b].p32, sizeof(int) * (size_t)N * Tp));
        CK(cudaMalloc(&a[b].p16, sizeof(short) * (size_t)N * Tp16));
        CK(cudaMalloc(&a[b].sink, 64));
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = (N / 32 + 3) / 4;
    auto run = [&](const char* name, auto launch) {
        for (int i = 0; i < 4; ++i) launch(a[i % NBUF]);
        cudaEventRecord(e0);
        const int reps = 40;
        for (int i = 0; i < reps; ++i) launch(a[i % NBUF]);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t e = cudaGetLastError();
        printf("%-34s %8.2f us  %s\n", name, ms * 1000 / reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    run("direct NT=8 full", [&](Args& x) { direct_kernel<8, 0><<<grid, 128>>>(x); });
    run("direct NT=16 full", [&](Args& x) { direct_kernel<16, 0><<<grid, 128>>>(x); });
    run("direct NT=32 full", [&](Args& x) { direct_kernel<32, 0><<<grid, 128>>>(x); });
    run("direct NT=16 no-u16", [&](Args& x) { direct_kernel<16, 3><<<grid, 128>>>(x); });
    run("direct NT=16 loads only", [&](Args& x) { direct_kernel<16, 1><<<grid, 128>>>(x); });
    run("direct NT=32 loads only", [&](Args& x) { direct_kernel<32, 1><<<grid, 128>>>(x); });
    run("direct NT=16 stores only", [&](Args& x) { direct_kernel<16, 2><<<grid, 128>>>(x); });
    run("direct NT=16 loads, 64-thr CTAs", [&](Args& x) { direct_kernel<16, 1><<<grid * 2, 64>>>(x); });
    run("staged NT=32 full", [&](Args& x) { staged_kernel<0><<<grid, 128>>>(x); });
    run("staged NT=32 no-u16", [&](Args& x) { staged_kernel<3><<<grid, 128>>>(x); });
    run("cta-staged NB=16 full", [&](Args& x) { cta_kernel<0, 16><<<N / 32, 128>>>(x); });
    run("cta-staged NB=32 full", [&](Args& x) { cta_kernel<0, 32><<<N / 32, 128>>>(x); });
    run("cta-staged NB=8 full", [&](Args& x) { cta_kernel<0, 8><<<N / 32, 128>>>(x); });
    run("cta-staged NB=16 no-u16", [&](Args& x) { cta_kernel<3, 16><<<N / 32, 128>>>(x); });
    const size_t nin = (size_t)T * N * 8 / 16, nout32 = (size_t)N * Tp * 4 / 16, nout = nout32 + (size_t)N * Tp16 * 2 / 16;
    run("stream copy same bytes (in+32+16)", [&](Args& x) { copy_kernel<<<148 * 8, 512>>>((const int4*)x.tok, nin, (int4*)x.p32, nout32); copy_kernel<<<148 * 8, 512>>>((const int4*)x.tok, 0, (int4*)x.p16, nout - nout32); });
    run("stream read only", [&](Args& x) { copy_kernel<<<148 * 8, 512>>>((const int4*)x.tok, nin, (int4*)x.p32, 0); });
    run("stream write only (32)", [&](Args& x) { copy_kernel<<<148 * 8, 512>>>((const int4*)x.tok, 0, (int4*)x.p32, nout32); });
    return 0;
}
