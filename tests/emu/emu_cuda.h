// emu_cuda.h -- a small SIMT emulator so the CUDA kernel sources under
// pydrobert-pytorch_b200/csrc compile with g++ and run on the CPU.
//
// TEST INFRASTRUCTURE ONLY.  It exists so that `pytest -m "not gpu"` can check the
// *logic* of the kernels (index arithmetic, wavefront skew, boundary hand-off,
// output epilogues) against the oracle in a container that has no GPU.  It is never
// built into, loaded by, or reachable from the product library; the product path has
// no CPU fallback.  Memory-model effects (races, fences) are NOT modelled: every
// CUDA thread is a cooperative fiber on one OS thread, switched only at
// synchronisation points (__syncthreads, __syncwarp, warp shuffles/votes, mbarrier
// waits, LEV_SPIN_YIELD).
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define B200LEV_EMU 1

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct longlong2 { long long x, y; };
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline longlong2 make_longlong2(long long x, long long y) { return longlong2{x, y}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
    memset(p, v, n);
    return 0;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) {
    memcpy(d, s, n);
    return 0;
}
#define cudaMemcpyDeviceToDevice 3
#define cudaFuncAttributeMaxDynamicSharedMemorySize 8
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

namespace emu {

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
    uint3 tid{0, 0, 0};
    int lin = 0, warp = 0, lane = 0;
};

struct WarpState {
    int count = 0;
    unsigned gen = 0;
    unsigned shfl_seq = 0;
    uint64_t buf[2][32];
};

struct Cta {
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    int bar_count = 0;
    unsigned bar_gen = 0;
    int live = 0;
    int nthreads = 0;
};

extern Cta g_cta;
extern Fiber* g_cur;
extern ucontext_t g_sched;
extern uint3 g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern unsigned char* g_dyn_smem;
extern std::function<void()>* g_body;
extern unsigned long long g_switches;

void yield();
void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body);

// warp-level rendezvous of the lanes in `mask`
static inline void warp_barrier(unsigned mask) {
    WarpState& w = g_cta.warps[g_cur->warp];
    int expected = __builtin_popcount(mask);
    unsigned my = w.gen;
    if (++w.count == expected) {
        w.count = 0;
        w.gen++;
    } else {
        while (w.gen == my) yield();
    }
}

template <typename T>
static inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle payload too large");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T>
static inline T from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

// exchange: every lane in mask publishes `v`; returns the value published by `src`
// (or own value if src is out of range / not participating)
template <typename T, typename SrcFn>
static inline T exchange(unsigned mask, T v, SrcFn src_of) {
    WarpState& w = g_cta.warps[g_cur->warp];
    int lane = g_cur->lane;
    // all lanes of one exchange see the same sequence number: it only advances
    // after the rendezvous below, and only lane-local copies are used afterwards
    unsigned seq = w.shfl_seq;
    int slot = seq & 1;
    w.buf[slot][lane] = to_bits(v);
    unsigned my = w.gen;
    int expected = __builtin_popcount(mask);
    if (++w.count == expected) {
        w.count = 0;
        w.shfl_seq++;
        w.gen++;
    } else {
        while (w.gen == my) yield();
    }
    int src = src_of(lane);
    if (src < 0 || src > 31 || !((mask >> src) & 1)) return v;
    return from_bits<T>(w.buf[slot][src]);
}

}  // namespace emu

#define threadIdx (emu::g_cur->tid)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define warpSize 32

static inline void __syncthreads() {
    emu::Cta& c = emu::g_cta;
    unsigned my = c.bar_gen;
    if (++c.bar_count == c.live) {
        c.bar_count = 0;
        c.bar_gen++;
    } else {
        while (c.bar_gen == my) emu::yield();
    }
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_barrier(mask); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) { emu::yield(); }

template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    return emu::exchange(mask, v, [=](int lane) { return (lane & ~(width - 1)) | (src & (width - 1)); });
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    return emu::exchange(mask, v, [=](int lane) {
        int base = lane & ~(width - 1);
        int s = lane - (int)d;
        return s < base ? lane : s;
    });
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    return emu::exchange(mask, v, [=](int lane) {
        int base = lane & ~(width - 1);
        int s = lane + (int)d;
        return s >= base + width ? lane : s;
    });
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
    return emu::exchange(mask, v, [=](int lane) {
        int s = lane ^ x;
        return (s & ~(width - 1)) == (lane & ~(width - 1)) ? s : lane;
    });
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    unsigned r = 0;
    emu::WarpState& w = emu::g_cta.warps[emu::g_cur->warp];
    unsigned seq = w.shfl_seq;
    int slot = seq & 1;
    (void)emu::exchange(mask, (int)(pred != 0), [](int lane) { return lane; });
    for (int l = 0; l < 32; ++l)
        if (((mask >> l) & 1) && (int)w.buf[slot][l]) r |= 1u << l;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline unsigned __activemask() { return 0xffffffffu; }

// ---- integer / DPX intrinsics -------------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __viaddmin_s32(int a, int b, int c) { return std::min(a + b, c); }
static inline unsigned __viaddmin_u32(unsigned a, unsigned b, unsigned c) { return std::min(a + b, c); }
static inline int __viaddmax_s32(int a, int b, int c) { return std::max(a + b, c); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) {
    shift &= 31;
    return shift ? (hi << shift) | (lo >> (32 - shift)) : hi;
}
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned sel) {
    const unsigned long long src = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned s = (sel >> (4 * i)) & 0xf;
        unsigned b = (unsigned)(src >> (8 * (s & 7))) & 0xff;
        if (s & 8) b = (b & 0x80) ? 0xff : 0;
        r |= b << (8 * i);
    }
    return r;
}
static inline int __vimin3_s32(int a, int b, int c) { return std::min(std::min(a, b), c); }
static inline int __vimax3_s32(int a, int b, int c) { return std::max(std::max(a, b), c); }
static inline unsigned __emu_pack16(int lo, int hi) {
    return ((unsigned)(uint16_t)(int16_t)lo) | (((unsigned)(uint16_t)(int16_t)hi) << 16);
}
static inline int __emu_lo16(unsigned x) { return (int16_t)(x & 0xffff); }
static inline int __emu_hi16(unsigned x) { return (int16_t)(x >> 16); }
static inline unsigned __viaddmin_s16x2(unsigned a, unsigned b, unsigned c) {
    int lo = std::min((int)(int16_t)(__emu_lo16(a) + __emu_lo16(b)), __emu_lo16(c));
    int hi = std::min((int)(int16_t)(__emu_hi16(a) + __emu_hi16(b)), __emu_hi16(c));
    return __emu_pack16(lo, hi);
}
static inline unsigned __vimin3_s16x2(unsigned a, unsigned b, unsigned c) {
    return __emu_pack16(std::min(std::min(__emu_lo16(a), __emu_lo16(b)), __emu_lo16(c)),
                        std::min(std::min(__emu_hi16(a), __emu_hi16(b)), __emu_hi16(c)));
}
static inline unsigned __vmins2(unsigned a, unsigned b) {
    return __emu_pack16(std::min(__emu_lo16(a), __emu_lo16(b)), std::min(__emu_hi16(a), __emu_hi16(b)));
}
static inline unsigned __viaddmin_u16x2(unsigned a, unsigned b, unsigned c) {  // halves add modulo 65536
    unsigned lo = std::min((a + b) & 0xffffu, c & 0xffffu), hi = std::min(((a >> 16) + (b >> 16)) & 0xffffu, c >> 16);
    return lo | (hi << 16);
}
static inline unsigned __vminu2(unsigned a, unsigned b) {
    unsigned lo = std::min(a & 0xffffu, b & 0xffffu), hi = std::min(a >> 16, b >> 16);
    return lo | (hi << 16);
}
static inline unsigned __vadd2(unsigned a, unsigned b) {
    return ((a + b) & 0xffffu) | ((((a >> 16) + (b >> 16)) & 0xffffu) << 16);
}
template <typename T>
static inline T __ldg(const T* p) { return *p; }
template <typename T>
static inline T __ldcg(const T* p) { return *p; }
template <typename T>
static inline T __ldcs(const T* p) { return *p; }
template <typename T>
static inline void __stcs(T* p, T v) { *p = v; }
template <typename T>
static inline void __stcg(T* p, T v) { *p = v; }
static inline float __int2float_rn(int x) { return (float)x; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline float __int_as_float(int x) { return emu::from_bits<float>((uint64_t)(uint32_t)x); }
static inline int __float_as_int(float x) { return (int)(uint32_t)emu::to_bits(x); }
using std::max;
using std::min;

// ---- atomics (single OS thread: plain read-modify-write) ---------------------------
template <typename T>
static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T>
static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T>
static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <typename T>
static inline T atomicMax(T* p, T v) { T o = *p; *p = std::max(o, v); return o; }
template <typename T>
static inline T atomicMin(T* p, T v) { T o = *p; *p = std::min(o, v); return o; }
template <typename T>
static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T>
static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

// ---- bf16 / fp16 storage types (conversion only) -----------------------------------
struct __nv_bfloat16 { uint16_t x; };
struct __half { uint16_t x; };
static inline float __bfloat162float(__nv_bfloat16 h) {
    uint32_t b = (uint32_t)h.x << 16;
    float f;
    memcpy(&f, &b, 4);
    return f;
}
static inline __nv_bfloat16 __float2bfloat16(float f) {  // round to nearest even
    uint32_t b;
    memcpy(&b, &f, 4);
    __nv_bfloat16 h;
    if ((b & 0x7fffffffu) > 0x7f800000u) { h.x = (uint16_t)((b >> 16) | 0x40); return h; }
    uint32_t lsb = (b >> 16) & 1;
    b += 0x7fffu + lsb;
    h.x = (uint16_t)(b >> 16);
    return h;
}
static inline float __half2float(__half h) {
    uint32_t s = (h.x >> 15) & 1, e = (h.x >> 10) & 0x1f, m = h.x & 0x3ff;
    float f;
    if (e == 0) f = ldexpf((float)m, -24);
    else if (e == 31) f = m ? NAN : INFINITY;
    else f = ldexpf((float)(m | 0x400), (int)e - 25);
    return s ? -f : f;
}
static inline __half __float2half(float f) {  // round to nearest even
    __half h;
    uint32_t b;
    memcpy(&b, &f, 4);
    uint32_t s = (b >> 16) & 0x8000u;
    float a = fabsf(f);
    if (std::isnan(f)) { h.x = (uint16_t)(s | 0x7e00); return h; }
    if (a >= 65520.0f) { h.x = (uint16_t)(s | 0x7c00); return h; }
    if (a < ldexpf(1.0f, -24) * 0.5f) { h.x = (uint16_t)s; return h; }
    int e;
    float m = frexpf(a, &e);  // a = m * 2^e, m in [0.5,1)
    int he = e + 14;          // biased exponent if normal
    uint32_t bits;
    if (he <= 0) {
        float q = nearbyintf(ldexpf(a, 24));
        bits = (uint32_t)q;
    } else {
        float q = nearbyintf(ldexpf(m, 11));  // 1024..2048
        uint32_t qi = (uint32_t)q;
        if (qi == 2048) { qi = 1024; he += 1; }
        bits = ((uint32_t)he << 10) | (qi & 0x3ff);
        if (he >= 31) bits = 0x7c00;
    }
    h.x = (uint16_t)(s | bits);
    return h;
}

// ---- launch ------------------------------------------------------------------------
template <typename... KArgs, typename... Args>
static inline void lev_launch(void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t, Args... args) {
    emu::launch(grid, block, smem, [=]() { k(KArgs(args)...); });
}
#define LEV_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::g_dyn_smem)
#define LEV_SPIN_YIELD() emu::yield()
#define LEV_OPAQUE_PTR(p) asm volatile("" : "+r"(p))
static inline void lev_prefetch_l2(const void*) {}
static inline void lev_st_f32_if(bool p, float* ptr, float v) { if (p) *ptr = v; }
static inline unsigned lev_ldg_l2(const unsigned* p) { return *p; }
template <typename T>
static inline T lev_ldg_stream(const T* p) { return *p; }
static inline void lev_cp_async16(void* smem_dst, const void* gsrc) { memcpy(smem_dst, gsrc, 16); }
static inline void lev_cp_async_commit() {}
template <int N>
static inline void lev_cp_async_wait() {}
static inline float __frcp_rn(float x) { return 1.0f / x; }
// mbarrier + bulk copy.  Barrier word: low 32 bits = completed phases, high 32 bits =
// bytes still expected in the current phase.  The copy itself happens at issue.
static inline void lev_mbar_init(unsigned long long* bar, unsigned) { *bar = 0; }
static inline void lev_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    *bar += (unsigned long long)bytes << 32;
}
static inline void lev_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes,
                                unsigned long long* bar) {
    memcpy(smem_dst, gsrc, bytes);
    *bar -= (unsigned long long)bytes << 32;
    if ((*bar >> 32) == 0) *bar += 1;  // all expected bytes landed: the phase completes
}
static inline void lev_mbar_wait(unsigned long long* bar, unsigned parity) {
    while (((*bar) & 1ull) == parity) emu::yield();
}
typedef uintptr_t lev_saddr;
static inline lev_saddr lev_saddr_of(const void* p) { return (lev_saddr)p; }
static inline int lev_lds32(lev_saddr a) { return *(const int*)a; }
static inline int lev_lds32_sync(lev_saddr a) { return *(const volatile int*)a; }
static inline void lev_sts32(lev_saddr a, int v) { *(volatile int*)a = v; }
static inline int lev_ld_volatile_shared(const int* p) { return *(const volatile int*)p; }
static inline void lev_st_volatile_shared(int* p, int v) { *(volatile int*)p = v; }
