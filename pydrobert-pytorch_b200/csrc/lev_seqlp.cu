// lev_seqlp.cu -- "next #1" (SURVEY 8f): sequence_log_probs, tensor path.
//
// Reference: _sequence_log_probs_tensor, src/pydrobert/torch/_decoding.py:1516-1548 ("DC"):
//   log_softmax over the class axis (DC:1529), gather the hypothesis token (DC:1546), zero
//   the steps whose token is outside [0, V) (DC:1530) or past the first eos -- the eos step
//   itself counts (DC:1531-1544) -- and sum over the step axis (DC:1548).  It is what turns
//   decoder logits into the `log_probs` argument of minimum_error_rate_loss.
//
// Pure HBM streaming: cfg2's (T=100, 512, V=10 000) bf16 logits are 1.02 GB.  One warp per
// (a, t, b) ROW of V logits, 128-bit loads, online logsumexp (one pass: running maximum and
// rescaled sum per lane, 9 exponentials per 8 logits).  Rows that are masked out are not
// read at all.  Per-row results go to an fp32 scratch and are summed over t in a fixed order
// (one warp per sequence, a fixed butterfly) by a second, tiny kernel: deterministic, no
// float atomics.  The logsumexp of every row is
// kept for the backward pass,
//   d logits[a,t,b,v] = g[a,b] * ([v == hyp] - exp(logits - lse))      (unmasked rows, else 0)
// which streams the logits once more and writes the gradient (another 1 + 1 GB at cfg2).
//
// Rounding follows torch: log_softmax returns the logits' dtype, so each step's value is
// rounded to it before the (fp32-accumulated) sum, whose result is rounded once more.
#include "lev_common.cuh"

namespace {

template <int DT> struct SlpElem;
template <> struct SlpElem<B200LEV_F32> {
    typedef float T; typedef float Acc;
    static constexpr int VEC = 4;
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
    static __device__ __forceinline__ float round(float v) { return v; }
};
template <> struct SlpElem<B200LEV_F64> {
    typedef double T; typedef double Acc;
    static constexpr int VEC = 2;
    static __device__ __forceinline__ double ld(const double* p) { return *p; }
    static __device__ __forceinline__ void st(double* p, double v) { *p = v; }
    static __device__ __forceinline__ double round(double v) { return v; }
};
template <> struct SlpElem<B200LEV_F16> {
    typedef __half T; typedef float Acc;
    static constexpr int VEC = 8;
    static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
    static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half(v); }
    static __device__ __forceinline__ float round(float v) { return __half2float(__float2half(v)); }
};
template <> struct SlpElem<B200LEV_BF16> {
    typedef __nv_bfloat16 T; typedef float Acc;
    static constexpr int VEC = 8;
    static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }
    static __device__ __forceinline__ float round(float v) { return __bfloat162float(__float2bfloat16(v)); }
};

// exp(x - m) as one FFMA + one SFU op in fp32: 2^(x log2e - m log2e), ex2.approx (2 ulp).  The
// terms that matter have arguments near 0, where the scaling's rounding error vanishes -- well
// inside the 2e-6 parity tolerance.  `ms` is the pre-scaled m (slp_scale).
#define SLP_LOG2E 1.4426950408889634f
__device__ __forceinline__ float slp_ex2(float y) {
#ifdef B200LEV_EMU
    return exp2f(y);
#else
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
    return r;
#endif
}
__device__ __forceinline__ float slp_scale(float m) { return m * SLP_LOG2E; }
__device__ __forceinline__ double slp_scale(double m) { return m; }
__device__ __forceinline__ float slp_expm(float x, float ms) { return slp_ex2(fmaf(x, SLP_LOG2E, -ms)); }
__device__ __forceinline__ double slp_expm(double x, double ms) { return exp(x - ms); }
__device__ __forceinline__ float slp_exp(float x) { return slp_ex2(x * SLP_LOG2E); }
__device__ __forceinline__ double slp_exp(double x) { return exp(x); }
__device__ __forceinline__ float slp_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double slp_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float slp_log(float x) { return logf(x); }
__device__ __forceinline__ double slp_log(double x) { return log(x); }

// 16 bytes of a row as VEC accumulation-type values (raw form: what stays in registers while
// several loads are in flight)
template <int DT>
__device__ __forceinline__ void slp_unpack(const uint4& raw,
                                           typename SlpElem<DT>::Acc (&x)[SlpElem<DT>::VEC]) {
    typedef typename SlpElem<DT>::T T;
    constexpr int VEC = SlpElem<DT>::VEC;
    T tmp[VEC];
    memcpy(tmp, &raw, 16);
#pragma unroll
    for (int k = 0; k < VEC; ++k) x[k] = SlpElem<DT>::ld(&tmp[k]);
}
#ifndef B200LEV_EMU
// bf16 -> fp32 is a 16-bit shift: one SHF (low half) or one LOP3 (high half) per value
template <>
__device__ __forceinline__ void slp_unpack<B200LEV_BF16>(const uint4& raw, float (&x)[8]) {
    const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        x[2 * k] = __uint_as_float(w[k] << 16);
        x[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
}
#endif
template <int DT>
__device__ __forceinline__ void slp_load_vec(const typename SlpElem<DT>::T* p,
                                             typename SlpElem<DT>::Acc (&x)[SlpElem<DT>::VEC]) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    slp_unpack<DT>(raw, x);
}

template <int DT>
__device__ __forceinline__ void slp_store_vec(typename SlpElem<DT>::T* p,
                                              const typename SlpElem<DT>::Acc (&x)[SlpElem<DT>::VEC]) {
    typedef typename SlpElem<DT>::T T;
    constexpr int VEC = SlpElem<DT>::VEC;
    T tmp[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) SlpElem<DT>::st(&tmp[k], x[k]);
    uint4 raw;
    memcpy(&raw, tmp, 16);
    *reinterpret_cast<uint4*>(p) = raw;
}

template <typename A>
__device__ __forceinline__ A slp_warp_sum(A v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(LEV_FULL_MASK, v, o);
    return v;
}
template <typename A>
__device__ __forceinline__ A slp_warp_max(A v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const A t = __shfl_xor_sync(LEV_FULL_MASK, v, o);
        v = t > v ? t : v;
    }
    return v;
}

}  // namespace

// steps that count for sequence (a, b): first eos index + 1 (the eos step is included), T
// when there is none (DC:1531-1532 with _lens_from_eos, _string.py:137-143)
__global__ void __launch_bounds__(256)
lev_seqlp_len_kernel(const int64_t* __restrict__ hyp, int64_t outer, int64_t T, int64_t inner,
                     int has_eos, int64_t eos, int32_t* __restrict__ len) {
    // one warp per sequence, lanes along t: the first eos is the lowest set ballot bit
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= outer * inner) return;
    const int64_t a = q / inner, b = q - a * inner;
    int64_t n = T;
    if (has_eos)
        for (int64_t t0 = 0; t0 < T; t0 += 32) {
            const int64_t t = t0 + lane;
            const bool hit = t < T && hyp[(a * T + t) * inner + b] == eos;
            const unsigned bits = __ballot_sync(LEV_FULL_MASK, hit);
            if (bits != 0) {
                n = t0 + __ffs((int)bits);
                break;
            }
        }
    if (lane == 0) len[q] = (int32_t)n;
}

// logsumexp of one row, all lanes return it
template <int DT>
__device__ __forceinline__ typename SlpElem<DT>::Acc slp_row_lse(const typename SlpElem<DT>::T* z,
                                                                 int64_t V, int lane) {
    typedef typename SlpElem<DT>::Acc A;
    constexpr int VEC = SlpElem<DT>::VEC;
    A m = -(A)INFINITY, ms = -(A)INFINITY, s = (A)0;  // ms = slp_scale(m)
    auto feed = [&](const A* x, int n) {
        A cm = x[0];
        for (int k = 1; k < n; ++k) cm = slp_max(cm, x[k]);
        if (cm > m) {  // rescale the running sum to the new maximum (exp(-inf) = 0 at first)
            s *= slp_exp(m - cm);
            m = cm;
            ms = slp_scale(cm);
        }
        if (m > -(A)INFINITY)
            for (int k = 0; k < n; ++k) s += slp_expm(x[k], ms);
    };
    const bool aligned = ((reinterpret_cast<uintptr_t>(z) & 15) == 0);
    int64_t v0 = 0;
    if (aligned) {
        const int64_t nvec = V / VEC;
        int64_t c = lane;
        // two groups of four 128-bit loads ping-pong: one group is in flight while the other
        // is being summed
        if (c + 96 < nvec) {
            uint4 ra[4], rb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) ra[u] = *reinterpret_cast<const uint4*>(z + (c + 32 * u) * VEC);
            c += 128;
            while (true) {
                const bool more_b = c + 96 < nvec;
                if (more_b) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) rb[u] = *reinterpret_cast<const uint4*>(z + (c + 32 * u) * VEC);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    A x[VEC];
                    slp_unpack<DT>(ra[u], x);
                    feed(x, VEC);
                }
                if (!more_b) break;
                c += 128;
                const bool more_a = c + 96 < nvec;
                if (more_a) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) ra[u] = *reinterpret_cast<const uint4*>(z + (c + 32 * u) * VEC);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    A x[VEC];
                    slp_unpack<DT>(rb[u], x);
                    feed(x, VEC);
                }
                if (!more_a) break;
                c += 128;
            }
        }
        for (; c < nvec; c += 32) {
            A x[VEC];
            slp_load_vec<DT>(z + c * VEC, x);
            feed(x, VEC);
        }
        v0 = nvec * VEC;
    }
    for (int64_t v = v0 + lane; v < V; v += 32) {
        const A x = (A)SlpElem<DT>::ld(z + v);
        feed(&x, 1);
    }
    // combine the lanes' (m, s)
    const A mx = slp_warp_max(m);
    A part = (m > -(A)INFINITY) ? s * slp_exp(m - mx) : (A)0;
    if (!(mx > -(A)INFINITY && mx < (A)INFINITY)) {
        // all -inf, or a +inf / NaN logit: torch's log_softmax gives NaN / -inf rows there;
        // fall back to the plain definition so those values propagate the same way
        A tot = (A)0;
        for (int64_t v = lane; v < V; v += 32) tot += slp_exp((A)SlpElem<DT>::ld(z + v) - mx);
        tot = slp_warp_sum(tot);
        return mx + slp_log(tot);
    }
    part = slp_warp_sum(part);
    return mx + slp_log(part);
}

// One warp per row (a, t, b).  row_lp = round(logits[hyp] - lse) for steps that count, else 0;
// row_lse = lse (0 for skipped rows).
template <int DT>
__global__ void __launch_bounds__(256)
lev_seqlp_row_kernel(const typename SlpElem<DT>::T* __restrict__ logits, int64_t rows, int64_t T,
                     int64_t inner, int64_t V, const int64_t* __restrict__ hyp,
                     const int32_t* __restrict__ len, typename SlpElem<DT>::Acc* __restrict__ row_lp,
                     typename SlpElem<DT>::Acc* __restrict__ row_lse) {
    typedef typename SlpElem<DT>::Acc A;
    const int lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
         row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t b = row % inner, at = row / inner;
        const int64_t t = at % T, a = at / T;
        const int64_t tok = hyp[row];
        const bool counts = tok >= 0 && tok < V && t < (int64_t)len[a * inner + b];  // DC:1530,1542-1544
        A lp = (A)0, lse = (A)0;
        if (counts) {  // warp-uniform
            const typename SlpElem<DT>::T* z = logits + row * V;
            lse = slp_row_lse<DT>(z, V, lane);
            lp = SlpElem<DT>::round((A)SlpElem<DT>::ld(z + tok) - lse);
        }
        if (lane == 0) {
            row_lp[row] = lp;
            row_lse[row] = lse;
        }
    }
}

// out[a, b] = round(sum_t row_lp[a, t, b]) in a fixed order (DC:1548)
template <int DT>
__global__ void __launch_bounds__(256)
lev_seqlp_sum_kernel(const typename SlpElem<DT>::Acc* __restrict__ row_lp, int64_t outer, int64_t T,
                     int64_t inner, typename SlpElem<DT>::T* __restrict__ out) {
    // one warp per sequence: lane l sums t = l, l + 32, ... in order, then a fixed butterfly
    typedef typename SlpElem<DT>::Acc A;
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= outer * inner) return;
    const int64_t a = q / inner, b = q - a * inner;
    A acc = (A)0;
    for (int64_t t = lane; t < T; t += 32) acc += row_lp[(a * T + t) * inner + b];
    acc = slp_warp_sum(acc);
    if (lane == 0) SlpElem<DT>::st(out + q, acc);
}

// grad_logits[row][v] = round(g[a, b] * ([v == tok] - exp(logits - lse))) on rows that count,
// zeros elsewhere.  One warp per row, 128-bit loads and stores.
template <int DT>
__global__ void __launch_bounds__(256)
lev_seqlp_bwd_kernel(const typename SlpElem<DT>::T* __restrict__ logits, int64_t rows, int64_t T,
                     int64_t inner, int64_t V, const int64_t* __restrict__ hyp,
                     const int32_t* __restrict__ len,
                     const typename SlpElem<DT>::Acc* __restrict__ row_lse,
                     const typename SlpElem<DT>::T* __restrict__ grad_out,
                     typename SlpElem<DT>::T* __restrict__ grad_logits) {
    typedef typename SlpElem<DT>::Acc A;
    constexpr int VEC = SlpElem<DT>::VEC;
    const int lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
         row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t b = row % inner, at = row / inner;
        const int64_t t = at % T, a = at / T;
        const int64_t tok = hyp[row];
        const bool counts = tok >= 0 && tok < V && t < (int64_t)len[a * inner + b];
        const typename SlpElem<DT>::T* z = logits + row * V;
        typename SlpElem<DT>::T* gz = grad_logits + row * V;
        const A g = counts ? (A)SlpElem<DT>::ld(grad_out + a * inner + b) : (A)0;
        const A lse = row_lse[row];
        const A lses = slp_scale(lse);
        const bool aligned = ((reinterpret_cast<uintptr_t>(z) & 15) == 0) &&
                             ((reinterpret_cast<uintptr_t>(gz) & 15) == 0);
        int64_t v0 = 0;
        if (aligned) {
            const int64_t nvec = V / VEC;
            if (counts) {
                int64_t c = lane;
                for (; c + 32 < nvec; c += 64) {  // two loads in flight per lane
                    A x0[VEC], x1[VEC], y[VEC];
                    slp_load_vec<DT>(z + c * VEC, x0);
                    slp_load_vec<DT>(z + (c + 32) * VEC, x1);
#pragma unroll
                    for (int k = 0; k < VEC; ++k)
                        y[k] = g * ((c * VEC + k == tok ? (A)1 : (A)0) - slp_expm(x0[k], lses));
                    slp_store_vec<DT>(gz + c * VEC, y);
#pragma unroll
                    for (int k = 0; k < VEC; ++k)
                        y[k] = g * (((c + 32) * VEC + k == tok ? (A)1 : (A)0) - slp_expm(x1[k], lses));
                    slp_store_vec<DT>(gz + (c + 32) * VEC, y);
                }
                for (; c < nvec; c += 32) {
                    A x[VEC], y[VEC];
                    slp_load_vec<DT>(z + c * VEC, x);
#pragma unroll
                    for (int k = 0; k < VEC; ++k)
                        y[k] = g * ((c * VEC + k == tok ? (A)1 : (A)0) - slp_expm(x[k], lses));
                    slp_store_vec<DT>(gz + c * VEC, y);
                }
            } else {
                A y[VEC];
#pragma unroll
                for (int k = 0; k < VEC; ++k) y[k] = (A)0;
                for (int64_t c = lane; c < nvec; c += 32) slp_store_vec<DT>(gz + c * VEC, y);
            }
            v0 = nvec * VEC;
        }
        for (int64_t v = v0 + lane; v < V; v += 32) {
            A y = (A)0;
            if (counts) y = g * ((v == tok ? (A)1 : (A)0) - slp_expm((A)SlpElem<DT>::ld(z + v), lses));
            SlpElem<DT>::st(gz + v, y);
        }
    }
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
static unsigned slp_row_grid(int64_t rows) {
    int64_t n = (rows + 7) / 8;  // 8 warps per CTA
    const int64_t cap = 148 * 16;
    return (unsigned)(n < 1 ? 1 : (n > cap ? cap : n));
}

template <int DT>
static int slp_forward(const void* logits, int64_t outer, int64_t T, int64_t inner, int64_t V,
                       const int64_t* hyp, int has_eos, int64_t eos, int32_t* len, void* row_lp,
                       void* row_lse, void* out, cudaStream_t st) {
    typedef typename SlpElem<DT>::T Tt;
    typedef typename SlpElem<DT>::Acc A;
    const int64_t seqs = outer * inner, rows = seqs * T;
    lev_launch(lev_seqlp_len_kernel, dim3((unsigned)((seqs + 7) / 8)), dim3(256), 0, st, hyp, outer, T,
               inner, has_eos, eos, len);
    if (rows > 0)
        lev_launch(lev_seqlp_row_kernel<DT>, dim3(slp_row_grid(rows)), dim3(256), 0, st, (const Tt*)logits,
                   rows, T, inner, V, hyp, (const int32_t*)len, (A*)row_lp, (A*)row_lse);
    lev_launch(lev_seqlp_sum_kernel<DT>, dim3((unsigned)((seqs + 7) / 8)), dim3(256), 0, st,
               (const A*)row_lp, outer, T, inner, (Tt*)out);
    return lev_check_cuda("lev_seqlp forward");
}

template <int DT>
static int slp_backward(const void* logits, int64_t outer, int64_t T, int64_t inner, int64_t V,
                        const int64_t* hyp, const int32_t* len, const void* row_lse,
                        const void* grad_out, void* grad_logits, cudaStream_t st) {
    typedef typename SlpElem<DT>::T Tt;
    typedef typename SlpElem<DT>::Acc A;
    const int64_t rows = outer * inner * T;
    if (rows > 0)
        lev_launch(lev_seqlp_bwd_kernel<DT>, dim3(slp_row_grid(rows)), dim3(256), 0, st, (const Tt*)logits,
                   rows, T, inner, V, hyp, len, (const A*)row_lse, (const Tt*)grad_out, (Tt*)grad_logits);
    return lev_check_cuda("lev_seqlp backward");
}

#define SLP_DT_SWITCH(dtype, CALL)                                        \
    switch (dtype) {                                                      \
        case B200LEV_F32: return CALL(B200LEV_F32);                       \
        case B200LEV_F16: return CALL(B200LEV_F16);                       \
        case B200LEV_BF16: return CALL(B200LEV_BF16);                     \
        case B200LEV_F64: return CALL(B200LEV_F64);                       \
        default:                                                          \
            lev_set_error("unsupported floating dtype code %d", (int)dtype); \
            return B200LEV_ERR_ARG;                                       \
    }

extern "C" int b200lev_seqlp_forward(const void* logits, int32_t dtype, int64_t outer, int64_t T,
                                     int64_t inner, int64_t V, const int64_t* hyp, int32_t has_eos,
                                     int64_t eos, int32_t* len, void* row_lp, void* row_lse, void* out,
                                     void* stream) {
    if (outer < 0 || T < 0 || inner < 0 || V < 1) {
        lev_set_error("b200lev_seqlp_forward: bad dimensions");
        return B200LEV_ERR_ARG;
    }
    if (outer * inner == 0) return B200LEV_OK;
    if (!hyp && T > 0) {
        lev_set_error("b200lev_seqlp_forward: NULL hyp");
        return B200LEV_ERR_ARG;
    }
    if ((T > 0 && !logits) || !len || !row_lp || !row_lse || !out) {
        lev_set_error("b200lev_seqlp_forward: NULL buffer");
        return B200LEV_ERR_ARG;
    }
#define CALL(DT_) slp_forward<DT_>(logits, outer, T, inner, V, hyp, has_eos, eos, len, row_lp, row_lse, out, (cudaStream_t)stream)
    SLP_DT_SWITCH(dtype, CALL)
#undef CALL
}

extern "C" int b200lev_seqlp_backward(const void* logits, int32_t dtype, int64_t outer, int64_t T,
                                      int64_t inner, int64_t V, const int64_t* hyp, const int32_t* len,
                                      const void* row_lse, const void* grad_out, void* grad_logits,
                                      void* stream) {
    if (outer < 0 || T < 0 || inner < 0 || V < 1) {
        lev_set_error("b200lev_seqlp_backward: bad dimensions");
        return B200LEV_ERR_ARG;
    }
    if (outer * inner * T == 0) return B200LEV_OK;
    if (!logits || !hyp || !len || !row_lse || !grad_out || !grad_logits) {
        lev_set_error("b200lev_seqlp_backward: NULL buffer");
        return B200LEV_ERR_ARG;
    }
#define CALL(DT_) slp_backward<DT_>(logits, outer, T, inner, V, hyp, len, row_lse, grad_out, grad_logits, (cudaStream_t)stream)
    SLP_DT_SWITCH(dtype, CALL)
#undef CALL
}
