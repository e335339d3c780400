// lev_arith.cuh -- value-type traits and the scalar epilogue shared by the DP kernels.
#pragma once
#include <type_traits>

#include "lev_common.cuh"

template <typename V>
struct LevArith;
template <>
struct LevArith<int> {
    static __device__ __forceinline__ int big() { return LEV_BIG_I32; }
    static __device__ __forceinline__ int ins(const LevParams& p) { return p.ins_i; }
    static __device__ __forceinline__ int del(const LevParams& p) { return p.del_i; }
    static __device__ __forceinline__ int sub(const LevParams& p) { return p.sub_i; }
};
template <>
struct LevArith<float> {
    static __device__ __forceinline__ float big() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float ins(const LevParams& p) { return p.ins_f; }
    static __device__ __forceinline__ float del(const LevParams& p) { return p.del_f; }
    static __device__ __forceinline__ float sub(const LevParams& p) { return p.sub_f; }
};

// SM:390-405 (final) and SM:356-378 (prefix row i): scale, normalise, empty-ref rule.
// `positive` is (hyp_len > 0) for the final value and (i > 0) for prefix row i.
__device__ __forceinline__ float lev_finalize(float val, const LevParams& p, int r,
                                              bool positive) {
    float v = val * p.mult;
    if (p.norm) v = (r == 0) ? (positive ? 1.0f : 0.0f) : v / (float)r;
    return v;
}

