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

// ======================= ctc_greedy_search (_decoding.py:507-560) ========================
// Row pass (one warp per (sequence, step) row of V classes, one read of the logits): the
// arg max (first index among equal maxima), the value the reference sums -- max - logsumexp
// for logits (DC:524-526,531), the max itself for probabilities -- and the row's logsumexp
// for the backward pass.  Path pass (one warp per sequence): blanks and repeats dropped
// (DC:532-534), steps beyond in_lens ignored (DC:536-545), survivors compacted to the front
// in order (DC:546-555: masked_select + masked_scatter_, so positions >= out_lens keep the raw
// arg max), values summed / multiplied over the valid steps (DC:551-554).
template <int DT>
struct SlpTop {
    typename SlpElem<DT>::Acc m, s;
    int64_t i;
};

template <int DT, bool PROBS>
__device__ __forceinline__ SlpTop<DT> slp_row_top(const typename SlpElem<DT>::T* z, int64_t V, int lane) {
    typedef typename SlpElem<DT>::Acc A;
    constexpr int VEC = SlpElem<DT>::VEC;
    A m = -(A)INFINITY, ms = -(A)INFINITY, s = (A)0;
    int64_t arg = 0;
    bool any = false;
    auto feed = [&](const A* x, int n, int64_t base) {
        A cm = x[0];
        for (int k = 1; k < n; ++k) cm = slp_max(cm, x[k]);
        if (cm > m || !any) {  // strictly greater: the first of equal maxima stays
            int ck = 0;         // (rare after the first few vectors: the index is found here only)
            for (int k = n - 1; k > 0; --k)
                if (x[k] == cm) ck = k;
            if (x[0] == cm) ck = 0;
            if (!PROBS && any) s *= slp_exp(m - cm);
            m = cm;
            ms = slp_scale(cm);
            arg = base + ck;
            any = true;
        }
        if (!PROBS && m > -(A)INFINITY)
            for (int k = 0; k < n; ++k) s += slp_expm(x[k], ms);
    };
    const bool aligned = ((reinterpret_cast<uintptr_t>(z) & 15) == 0);
    int64_t v0 = 0;
    if (aligned) {
        const int64_t nvec = V / VEC;
        int64_t c = lane;
        for (; c + 96 < nvec; c += 128) {  // four 128-bit loads in flight per lane
            uint4 r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) r[u] = *reinterpret_cast<const uint4*>(z + (c + 32 * u) * VEC);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                A x[VEC];
                slp_unpack<DT>(r[u], x);
                feed(x, VEC, (c + 32 * u) * VEC);
            }
        }
        for (; c < nvec; c += 32) {
            A x[VEC];
            slp_load_vec<DT>(z + c * VEC, x);
            feed(x, VEC, c * VEC);
        }
        v0 = nvec * VEC;
    }
    for (int64_t v = v0 + lane; v < V; v += 32) {
        A x[1] = {(A)SlpElem<DT>::ld(z + v)};
        feed(x, 1, v);
    }
    // butterfly: larger maximum wins, equal maxima -> smaller index; sums rescaled to the winner
    for (int o = 16; o > 0; o >>= 1) {
        const A m2 = __shfl_xor_sync(0xffffffffu, m, o);
        const A s2 = __shfl_xor_sync(0xffffffffu, s, o);
        const int64_t i2 = __shfl_xor_sync(0xffffffffu, arg, o);
        const bool a2 = __shfl_xor_sync(0xffffffffu, (int)any, o) != 0;
        if (a2 && (!any || m2 > m || (m2 == m && i2 < arg))) {
            if (!PROBS) s = (any && m > -(A)INFINITY ? s * slp_exp(m - m2) : (A)0) + s2;
            m = m2;
            arg = i2;
            any = true;
        } else if (a2 && !PROBS) {
            s += (m2 > -(A)INFINITY ? s2 * slp_exp(m2 - m) : (A)0);
        }
    }
    SlpTop<DT> r;
    r.m = m;
    r.s = s;
    r.i = arg;
    return r;
}

template <int DT, bool PROBS>
__global__ void __launch_bounds__(256)
lev_ctc_row_kernel(const typename SlpElem<DT>::T* __restrict__ logits, int64_t rows, int64_t V,
                   int64_t* __restrict__ arg, typename SlpElem<DT>::Acc* __restrict__ row_val,
                   typename SlpElem<DT>::Acc* __restrict__ row_lse) {
    typedef typename SlpElem<DT>::Acc A;
    const int lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
         row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const SlpTop<DT> top = slp_row_top<DT, PROBS>(logits + row * V, V, lane);
        if (lane == 0) {
            arg[row] = top.i;
            if (PROBS) {
                row_val[row] = top.m;
                row_lse[row] = (A)0;
            } else {
                const A lse = top.m + slp_log(top.s);
                row_lse[row] = lse;
                row_val[row] = SlpElem<DT>::round(top.m - lse);  // log_softmax in the logits' dtype
            }
        }
    }
}

template <int DT, bool PROBS>
__global__ void __launch_bounds__(256)
lev_ctc_path_kernel(const int64_t* __restrict__ arg, const typename SlpElem<DT>::Acc* __restrict__ row_val,
                    const int64_t* __restrict__ in_lens, int64_t outer, int64_t T, int64_t inner,
                    int64_t blank, int32_t* __restrict__ len, int64_t* __restrict__ paths,
                    int64_t* __restrict__ out_lens, typename SlpElem<DT>::T* __restrict__ max_out) {
    typedef typename SlpElem<DT>::Acc A;
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= outer * inner) return;
    const int64_t a = q / inner, b = q - a * inner;
    int64_t n = T;
    if (in_lens != nullptr) n = in_lens[q];
    const int64_t nv = n < 0 ? 0 : (n > T ? T : n);  // arange(T) < in_lens
    int64_t count = 0, last = -1;
    A acc = PROBS ? (A)1 : (A)0;
    for (int64_t t0 = 0; t0 < T; t0 += 32) {
        const int64_t t = t0 + lane;
        const int64_t cur = t < T ? arg[(a * T + t) * inner + b] : (int64_t)-1;
        int64_t prev = __shfl_up_sync(0xffffffffu, cur, 1);
        if (lane == 0) prev = last;
        last = __shfl_sync(0xffffffffu, cur, 31);
        const bool keep = t < nv && cur != blank && (t == 0 || cur != prev);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) paths[(a * T + count + __popc(bal & ((1u << lane) - 1u))) * inner + b] = cur;
        count += __popc(bal);
        if (t < nv) {
            const A v = row_val[(a * T + t) * inner + b];
            acc = PROBS ? SlpElem<DT>::round(acc * v) : acc + v;
        }
    }
    // positions past the reduced path keep the raw arg max (masked_scatter_ leaves them)
    for (int64_t t = count + lane; t < T; t += 32) paths[(a * T + t) * inner + b] = arg[(a * T + t) * inner + b];
    // fixed butterfly over the lanes' partial sums / products
    for (int o = 16; o > 0; o >>= 1) {
        const A other = __shfl_xor_sync(0xffffffffu, acc, o);
        acc = PROBS ? SlpElem<DT>::round(acc * other) : acc + other;
    }
    if (lane == 0) {
        out_lens[q] = count;
        len[q] = (int32_t)nv;
        SlpElem<DT>::st(max_out + q, acc);
    }
}

template <int DT>
static int slp_ctc(const void* logits, int64_t outer, int64_t T, int64_t inner, int64_t V,
                   const int64_t* in_lens, int64_t blank, int32_t probs, int64_t* arg, void* row_val,
                   void* row_lse, int32_t* len, int64_t* paths, int64_t* out_lens, void* max_out,
                   cudaStream_t st) {
    typedef typename SlpElem<DT>::T Tt;
    typedef typename SlpElem<DT>::Acc A;
    const int64_t rows = outer * inner * T, nseq = outer * inner;
    if (rows > 0) {
        if (probs)
            lev_launch(lev_ctc_row_kernel<DT, true>, dim3(slp_row_grid(rows)), dim3(256), 0, st,
                       (const Tt*)logits, rows, V, arg, (A*)row_val, (A*)row_lse);
        else
            lev_launch(lev_ctc_row_kernel<DT, false>, dim3(slp_row_grid(rows)), dim3(256), 0, st,
                       (const Tt*)logits, rows, V, arg, (A*)row_val, (A*)row_lse);
    }
    const dim3 grid((unsigned)((nseq + 7) / 8));
    if (probs)
        lev_launch(lev_ctc_path_kernel<DT, true>, grid, dim3(256), 0, st, (const int64_t*)arg,
                   (const A*)row_val, in_lens, outer, T, inner, blank, len, paths, out_lens, (Tt*)max_out);
    else
        lev_launch(lev_ctc_path_kernel<DT, false>, grid, dim3(256), 0, st, (const int64_t*)arg,
                   (const A*)row_val, in_lens, outer, T, inner, blank, len, paths, out_lens, (Tt*)max_out);
    return lev_check_cuda("lev_ctc_greedy");
}

extern "C" int b200lev_ctc_greedy(const void* logits, int32_t dtype, int64_t outer, int64_t T,
                                  int64_t inner, int64_t V, const int64_t* in_lens, int64_t blank,
                                  int32_t is_probs, int64_t* arg, void* row_val, void* row_lse,
                                  int32_t* len, int64_t* paths, int64_t* out_lens, void* max_out,
                                  void* stream) {
    if (outer < 0 || T < 0 || inner < 0 || V < 1 || blank < 0 || blank >= V) {
        lev_set_error("b200lev_ctc_greedy: bad dimensions or blank index");
        return B200LEV_ERR_ARG;
    }
    if (outer * inner == 0) return B200LEV_OK;
    if ((T > 0 && (!logits || !arg || !row_val || !row_lse || !paths)) || !len || !out_lens || !max_out) {
        lev_set_error("b200lev_ctc_greedy: NULL buffer");
        return B200LEV_ERR_ARG;
    }
#define CALL(DT_) slp_ctc<DT_>(logits, outer, T, inner, V, in_lens, blank, is_probs, arg, row_val, row_lse, len, paths, out_lens, max_out, (cudaStream_t)stream)
    SLP_DT_SWITCH(dtype, CALL)
#undef CALL
}
