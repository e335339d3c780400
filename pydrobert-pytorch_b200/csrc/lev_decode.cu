// lev_decode.cu -- SURVEY 8f #3, the N-best producers' step functions:
//   beam_search_advance  (_decoding.py:41-155)   top-k over the (old_width x V) extensions of a
//                                                beam + gather of the surviving prefixes
//   random_walk_advance  (_decoding.py:1207-1283) append one sampled token per path
// so that hypotheses -> sequence_log_probs -> minimum_error_rate_loss stays on the device.
//
// lev_beam_topk_kernel: one CTA per batch element.  The candidate score of (k, v) is
// log_probs_prev[n, k] + log_probs_t[n, k, v], rounded to the tensors' dtype as torch's add
// does.  Scores are mapped to order-preserving unsigned keys (NaN greatest, as torch.topk) and
// the K best are extracted one per pass: every thread scans its strided share of the
// old_width * V candidates for the best one strictly BELOW the previous winner in the total
// order (key descending, flat index ascending), a block reduction picks the pass's winner.
// K passes over data that sits in L2 (8 x 10 000 fp32 = 320 KB per element at config 2) -- a
// beam is a few paths wide.  Equal scores resolve to the lower flat index (torch.topk leaves
// the order of ties unspecified; the reference's own test only pins the values, tests/
// test_decoding.py:701-760).
//
// lev_path_extend_kernel: y_next[s, n, k] = y_prev[s, n, src[n, k]] below the new row, the new
// token at row lens_prev[n, src] (or at the new last row), and -- when the tensor grows -- the
// new token in the whole new last row, exactly what the reference's cat + scatter leaves there.
#include "lev_common.cuh"

// storage type <-> arithmetic (fp32 for the 16/32-bit types, fp64 for double); the 16-bit types
// only convert (same helpers in the CUDA and the emulated build)
template <typename T>
struct LevDec;
template <>
struct LevDec<float> {
    typedef float acc;
    static __device__ __forceinline__ float load(const float* p) { return *p; }
    static __device__ __forceinline__ float round(float x) { return x; }
    static __device__ __forceinline__ void store(float* p, float x) { *p = x; }
};
template <>
struct LevDec<double> {
    typedef double acc;
    static __device__ __forceinline__ double load(const double* p) { return *p; }
    static __device__ __forceinline__ double round(double x) { return x; }
    static __device__ __forceinline__ void store(double* p, double x) { *p = x; }
};
template <>
struct LevDec<__half> {
    typedef float acc;
    static __device__ __forceinline__ float load(const __half* p) { return __half2float(*p); }
    static __device__ __forceinline__ float round(float x) { return __half2float(__float2half(x)); }
    static __device__ __forceinline__ void store(__half* p, float x) { *p = __float2half(x); }
};
template <>
struct LevDec<__nv_bfloat16> {
    typedef float acc;
    static __device__ __forceinline__ float load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ float round(float x) { return __bfloat162float(__float2bfloat16(x)); }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, float x) { *p = __float2bfloat16(x); }
};

// order-preserving key of a double (covers fp32 / fp16 / bf16 exactly): larger value = larger
// key, NaN above +inf
__device__ __forceinline__ unsigned long long lev_dec_key(double v) {
    if (v != v) return ~0ull;
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

struct LevBeamCand {
    unsigned long long key;
    long long idx;
};
// a is better than b: larger key, then lower index
__device__ __forceinline__ bool lev_beam_better(const LevBeamCand& a, const LevBeamCand& b) {
    return a.key > b.key || (a.key == b.key && a.idx < b.idx);
}

// T: storage type; the candidate sum is formed in fp32 (fp64 for double) and rounded to T, as
// torch's `log_probs_prev.unsqueeze(2) + log_probs_t` is.
template <typename T>
__global__ void __launch_bounds__(256)
lev_beam_topk_kernel(const T* __restrict__ lpt, int64_t s_n, int64_t s_k, int64_t s_v,
                     const T* __restrict__ lpp, int64_t p_n, int64_t p_k, int64_t Kp, int64_t V, int K,
                     T* __restrict__ out_lp, int64_t* __restrict__ next_src, int64_t* __restrict__ y_t,
                     int width) {
    __shared__ LevBeamCand red[8];
    __shared__ LevBeamCand last;
    const int64_t n = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t total = Kp * V;
    const T* __restrict__ row = lpt + n * s_n;
    typedef typename LevDec<T>::acc ACC;
    auto cand_sum = [&](int64_t flat) -> ACC {
        const int64_t k = flat / V, v = flat - k * V;
        return LevDec<T>::round(LevDec<T>::load(row + k * s_k + v * s_v) + LevDec<T>::load(lpp + n * p_n + k * p_k));
    };
    auto cand_value = [&](int64_t flat) -> double { return (double)cand_sum(flat); };
    if (tid == 0) {
        last.key = ~0ull;
        last.idx = -1;  // everything is below (max key, index -1)
    }
    __syncthreads();
    for (int sel = 0; sel < K; ++sel) {
        const LevBeamCand prev = last;
        LevBeamCand best;
        best.key = 0ull;
        best.idx = -1;  // none
        bool have = false;
        for (int64_t f = tid; f < total; f += blockDim.x) {
            LevBeamCand c;
            c.key = lev_dec_key(cand_value(f));
            c.idx = f;
            if (!lev_beam_better(prev, c)) continue;  // not strictly below the previous winner
            if (!have || lev_beam_better(c, best)) {
                best = c;
                have = true;
            }
        }
        if (!have) best.idx = 0x7fffffffffffffffLL;  // loses against every real candidate
        // block arg-best: warp shuffles, then the 8 warp winners
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            LevBeamCand o;
            o.key = __shfl_xor_sync(LEV_FULL_MASK, best.key, d);
            o.idx = __shfl_xor_sync(LEV_FULL_MASK, best.idx, d);
            const bool ohave = o.idx != 0x7fffffffffffffffLL;
            const bool mhave = best.idx != 0x7fffffffffffffffLL;
            if (ohave && (!mhave || lev_beam_better(o, best))) best = o;
        }
        if (lane == 0) red[warp] = best;
        __syncthreads();
        if (tid == 0) {
            LevBeamCand w = red[0];
            for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
                const bool ohave = red[i].idx != 0x7fffffffffffffffLL;
                const bool mhave = w.idx != 0x7fffffffffffffffLL;
                if (ohave && (!mhave || lev_beam_better(red[i], w))) w = red[i];
            }
            last = w;
            const int64_t k = w.idx / V, v = w.idx - k * V;
            LevDec<T>::store(out_lp + n * width + sel, cand_sum(w.idx));
            next_src[n * width + sel] = k;
            y_t[n * width + sel] = v;
        }
        __syncthreads();
    }
    // too few extensions to fill the beam (_decoding.py:143-152): -inf, source 0
    for (int sel = K + tid; sel < width; sel += blockDim.x) {
        LevDec<T>::store(out_lp + n * width + sel, (ACC)(-__int_as_float(0x7f800000)));
        next_src[n * width + sel] = 0;
        y_t[n * width + sel] = 0;
    }
}

// y_next (S_out, N, W) from y_prev (S, N, Kp): columns k < K are live.
//   src == NULL: the path keeps its column (random walk); lens_prev == NULL: every prefix has
//   length S.  grown = S_out > S.
__global__ void __launch_bounds__(256)
lev_path_extend_kernel(const int64_t* __restrict__ y_prev, int64_t S, int64_t N, int64_t Kp,
                       const int64_t* __restrict__ src, const int64_t* __restrict__ lens_prev,
                       const int64_t* __restrict__ y_t, int64_t K, int64_t W, int64_t S_out,
                       int64_t* __restrict__ y_next, int64_t* __restrict__ lens_next) {
    const int64_t total = S_out * N * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i % W, n = (i / W) % N, s = i / (W * N);
        if (k >= K) {  // padding columns: the reference leaves them uninitialised; zeros here
            y_next[i] = 0;
            if (s == 0 && lens_next != nullptr) lens_next[n * W + k] = 0;
            continue;
        }
        const int64_t from = src != nullptr ? src[n * W + k] : k;
        const int64_t len = lens_prev != nullptr ? lens_prev[n * Kp + from] : S;
        const int64_t tok = y_t[n * W + k];
        int64_t val;
        if (s == len || s >= S)  // the new token's row, or the whole new last row (cat of y_t)
            val = tok;
        else
            val = y_prev[(s * N + n) * Kp + from];
        y_next[i] = val;
        if (s == 0 && lens_next != nullptr) lens_next[n * W + k] = len + 1;
    }
}

template <typename T>
static void lev_beam_topk_launch(const void* lpt, int64_t s_n, int64_t s_k, int64_t s_v, const void* lpp,
                                 int64_t p_n, int64_t p_k, int64_t N, int64_t Kp, int64_t V, int K,
                                 void* out_lp, int64_t* next_src, int64_t* y_t, int width, cudaStream_t st) {
    lev_launch(lev_beam_topk_kernel<T>, dim3((unsigned)N), dim3(256), 0, st, (const T*)lpt, s_n, s_k, s_v,
               (const T*)lpp, p_n, p_k, Kp, V, K, (T*)out_lp, next_src, y_t, width);
}

extern "C" int b200lev_beam_topk(const void* log_probs_t, int32_t dtype, int64_t N, int64_t Kp, int64_t V,
                                 int64_t stride_n, int64_t stride_k, int64_t stride_v,
                                 const void* log_probs_prev, int64_t prev_stride_n, int64_t prev_stride_k,
                                 int64_t width, void* log_probs_next, int64_t* next_src, int64_t* y_t,
                                 void* stream) {
    if (N <= 0 || width <= 0) return B200LEV_OK;
    if (!log_probs_t || !log_probs_prev || !log_probs_next || !next_src || !y_t || Kp < 0 || V < 0) {
        lev_set_error("b200lev_beam_topk: bad arguments");
        return B200LEV_ERR_ARG;
    }
    if (N >= ((int64_t)1 << 31) || width >= ((int64_t)1 << 30)) {
        lev_set_error("b200lev_beam_topk: dimension too large");
        return B200LEV_ERR_UNSUPPORTED;
    }
    const int64_t cand = Kp * V;
    const int K = (int)(width < cand ? width : cand);
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case 0:
            lev_beam_topk_launch<float>(log_probs_t, stride_n, stride_k, stride_v, log_probs_prev,
                                               prev_stride_n, prev_stride_k, N, Kp, V, K, log_probs_next,
                                               next_src, y_t, (int)width, st);
            break;
        case 1:
            lev_beam_topk_launch<__half>(log_probs_t, stride_n, stride_k, stride_v, log_probs_prev,
                                                prev_stride_n, prev_stride_k, N, Kp, V, K, log_probs_next,
                                                next_src, y_t, (int)width, st);
            break;
        case 2:
            lev_beam_topk_launch<__nv_bfloat16>(log_probs_t, stride_n, stride_k, stride_v,
                                                       log_probs_prev, prev_stride_n, prev_stride_k, N, Kp, V,
                                                       K, log_probs_next, next_src, y_t, (int)width, st);
            break;
        case 3:
            lev_beam_topk_launch<double>(log_probs_t, stride_n, stride_k, stride_v, log_probs_prev,
                                                 prev_stride_n, prev_stride_k, N, Kp, V, K, log_probs_next,
                                                 next_src, y_t, (int)width, st);
            break;
        default:
            lev_set_error("b200lev_beam_topk: unsupported dtype code %d", (int)dtype);
            return B200LEV_ERR_ARG;
    }
    return lev_check_cuda("lev_beam_topk_kernel");
}

extern "C" int b200lev_path_extend(const int64_t* y_prev, int64_t S, int64_t N, int64_t Kp,
                                   const int64_t* src, const int64_t* lens_prev, const int64_t* y_t,
                                   int64_t K, int64_t W, int64_t S_out, int64_t* y_next,
                                   int64_t* lens_next, void* stream) {
    const int64_t total = S_out * N * W;
    if (total <= 0) return B200LEV_OK;
    if (!y_t || !y_next || (S > 0 && !y_prev) || K > W || S_out < S || S_out > S + 1) {
        lev_set_error("b200lev_path_extend: bad arguments");
        return B200LEV_ERR_ARG;
    }
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    lev_launch(lev_path_extend_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, y_prev, S, N,
               Kp, src, lens_prev, y_t, K, W, S_out, y_next, lens_next);
    return lev_check_cuda("lev_path_extend_kernel");
}
