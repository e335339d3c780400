// simt.h -- the only place that knows whether the kernels are being compiled by
// nvcc for sm_100a (the product) or by g++ against tests/emu/emu_cuda.h (kernel-logic
// tests on a GPU-less box; never shipped, never loaded by the product).
#pragma once

#ifdef B200LEV_EMU
// tests/emu/emu_cuda.h was force-included by the test build and supplies the CUDA
// vocabulary (threadIdx, __shfl_*_sync, __syncthreads, atomics, lev_launch, ...).
#else
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>

template <typename... KArgs, typename... Args>
static inline void lev_launch(void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st, Args... args) {
    k<<<grid, block, smem, st>>>(KArgs(args)...);
}
#define LEV_DYN_SMEM(type, name)                                   \
    extern __shared__ __align__(16) unsigned char name##_raw_[];   \
    type* name = reinterpret_cast<type*>(name##_raw_)
#define LEV_SPIN_YIELD() ((void)0)

// Ampere-style asynchronous global->shared copies (LDGSTS), 16 bytes per call
__device__ __forceinline__ void lev_cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void lev_cp_async_commit() {
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void lev_cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// TMA bulk copy (cp.async.bulk, SASS UBLKCP) global -> shared, completion on an mbarrier
__device__ __forceinline__ void lev_mbar_init(unsigned long long* bar, unsigned count) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void lev_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes)
                 : "memory");
}
// bytes must be a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void lev_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes,
                                             unsigned long long* bar) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
        "l"(gsrc), "r"(bytes), "r"(b)
        : "memory");
}
__device__ __forceinline__ void lev_mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LEV_MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LEV_MBAR_DONE_%=;\n"
        "bra LEV_MBAR_WAIT_%=;\n"
        "LEV_MBAR_DONE_%=:\n"
        "}" ::"r"(a),
        "r"(parity)
        : "memory");
}
// Shared-memory accesses through a 32-bit shared-space address computed ONCE: inside the
// DP step loops this keeps the generic->shared conversion (S2UR SR_CgaCtaId + ULEA on
// sm_100) off the per-step critical path.
typedef unsigned lev_saddr;
__device__ __forceinline__ lev_saddr lev_saddr_of(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ int lev_lds32(lev_saddr a) {
    // volatile + memory clobber: the same shared address holds different data for every
    // pair a CTA processes, so the load must never be commoned or hoisted across the
    // staging barrier
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ int lev_lds32_sync(lev_saddr a) {  // data another warp produces
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void lev_sts32(lev_saddr a, int v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// ask L2 for a line a later load will want (no register, no scoreboard)
__device__ __forceinline__ void lev_prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// a store under a predicate: no branch (and no convergence barrier pair) around one instruction
__device__ __forceinline__ void lev_st_f32_if(bool p, float* ptr, float v) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.s32 q, %0, 0;\n"
        "@q st.global.f32 [%1], %2;\n"
        "}" ::"r"((int)p),
        "l"(ptr), "f"(v)
        : "memory");
}
// hide a pointer's derivation from the optimiser, so that the addresses built from it stay
// "pointer + small constant x stride" (one IMAD.WIDE each) instead of being re-derived
#define LEV_OPAQUE_PTR(p) asm volatile("" : "+l"(p))
// a value other CTAs may be raising concurrently: read it at L2, never from a stale L1 line
__device__ __forceinline__ unsigned lev_ldg_l2(const unsigned* p) { return __ldcg(p); }
// read-once global data: evict-first so it does not displace what the next kernel re-reads
template <typename T>
__device__ __forceinline__ T lev_ldg_stream(const T* p) {
    return __ldcs(p);
}
template <>
__device__ __forceinline__ int64_t lev_ldg_stream<int64_t>(const int64_t* p) {
    return (int64_t)__ldcs(reinterpret_cast<const long long*>(p));
}
template <>
__device__ __forceinline__ int8_t lev_ldg_stream<int8_t>(const int8_t* p) {
    return (int8_t)__ldcs(reinterpret_cast<const signed char*>(p));
}
__device__ __forceinline__ int lev_ld_volatile_shared(const int* p) {
    return *reinterpret_cast<const volatile int*>(p);
}
__device__ __forceinline__ void lev_st_volatile_shared(int* p, int v) {
    *reinterpret_cast<volatile int*>(p) = v;
}
#endif

#define LEV_FULL_MASK 0xffffffffu
