// lev_completion.cu -- K4: the sorted-unique compaction of optimal_completion.
//
// The reference (SM:492-517) turns the (H', R, N) position mask into per-prefix token
// sets by (i) OR-ing the mask across duplicate tokens with an (H', N, R, R) boolean
// temporary, (ii) sorting the reference tokens, (iii) masked_select/masked_scatter.
// Here the sort happens ONCE per reference instead of once per prefix:
//   lev_uid_kernel      : for every reference position j its rank uid[j] among the
//                         DISTINCT token values of that reference (ascending, as
//                         int64), plus the table dtok[rank] -> token value (a bitonic
//                         sort per reference);
//   (lev_dp.cu, MASK)   : sets bit uid[j] of the (prefix, pair) bitmap for every
//                         flagged position -- duplicates collapse onto one bit and
//                         ascending bit order IS ascending token order;
//   lev_completion_fill : enumerates the set bits into the (H', N, U) int64 output,
//                         padding the tail (coalesced: consecutive threads own
//                         consecutive U-element rows).
#include "lev_common.cuh"

// One CTA per reference: a bitonic sort of its (token, position) pairs in shared memory, then
// "starts a new value" flags, their prefix sum (= the rank among the DISTINCT tokens, ascending
// as int64) and the scatter back to the positions.  O(r log^2 r) compare-exchanges instead of the
// O(r^2) scan of every position against every other (16 k warp-instructions per 200-token
// reference there, 0.13 ms on config 3 -- a quarter of the whole call).
template <typename TT>
__global__ void __launch_bounds__(256)
lev_uid_kernel(const TT* __restrict__ tok, int64_t st, int64_t sn, const int32_t* __restrict__ packed,
               const int* __restrict__ state, const int32_t* __restrict__ ref_len,
               int32_t* __restrict__ uid, int64_t* __restrict__ dtok, int32_t* __restrict__ ndist,
               int64_t Rp, int n_pad) {
    LEV_DYN_SMEM(int64_t, key);                      // [n_pad] token values
    int* pos = reinterpret_cast<int*>(key + n_pad);  // [n_pad] positions, then ranks
    __shared__ int carry;
    const int64_t n = blockIdx.x;
    const int r = ref_len[n];
    const int tid = threadIdx.x, nt = blockDim.x;
    {
        // a call whose tokens span fewer than 32 values: the bit of a token is its offset from
        // the smallest one (ascending bit order is still ascending token order, duplicates still
        // collapse); lev_completion_fill turns bits back into tokens by the same rule
        int tmin;
        if (lev_tokens_direct(state, &tmin)) {
            for (int j = tid; j < r; j += nt) uid[n * Rp + j] = packed[n * Rp + j] - tmin;
            if (tid == 0) ndist[n] = 32;
            return;
        }
    }
    // the smallest power of two that holds this reference (uniform per CTA)
    int m = 32;
    while (m < r) m <<= 1;
    // K0's pair-major int32 table holds the tokens exactly unless it flagged one outside int32:
    // contiguous rows there, a stride of N elements between positions in the caller's tensor
    const bool wide = (*state & B200LEV_FLAG_WIDE_TOKENS) != 0;
    if (!wide) {
        // (token, position) in ONE 64-bit key: a compare-exchange is one compare and two 8-byte
        // swaps, and thread t owns the t-th exchange of a stage (no idle half)
        for (int j = tid; j < m; j += nt)
            key[j] = j < r ? (int64_t)(((uint64_t)(uint32_t)packed[n * Rp + j] << 32) | (uint32_t)j)
                           : (int64_t)0x7fffffffffffffffLL;  // sorts behind a real INT32_MAX token
        for (int k = 2; k <= m; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                __syncthreads();
                for (int t = tid; t < (m >> 1); t += nt) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int q = i | j;
                    const int64_t a = key[i], b = key[q];
                    if ((a > b) == ((i & k) == 0)) {
                        key[i] = b;
                        key[q] = a;
                    }
                }
            }
        }
        __syncthreads();
        // unpack: key -> token value, pos -> position
        for (int j = tid; j < m; j += nt) {
            const int64_t kk = key[j];
            pos[j] = (int)(uint32_t)kk;
            key[j] = (int64_t)(int32_t)(kk >> 32);
        }
    } else {
        for (int j = tid; j < m; j += nt) {
            key[j] = j < r ? (int64_t)tok[(int64_t)j * st + n * sn] : (int64_t)0x7fffffffffffffffLL;
            pos[j] = j < r ? j : 0x7fffffff;  // padding sorts behind a real INT64_MAX token
        }
        for (int k = 2; k <= m; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                __syncthreads();
                for (int t = tid; t < (m >> 1); t += nt) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int q = i | j;
                    const int64_t a = key[i], b = key[q];
                    const int pa = pos[i], pb = pos[q];
                    const bool a_gt_b = a > b || (a == b && pa > pb);
                    if (a_gt_b == ((i & k) == 0)) {
                        key[i] = b;
                        key[q] = a;
                        pos[i] = pb;
                        pos[q] = pa;
                    }
                }
            }
        }
    }
    __syncthreads();
    // ranks: warp 0 walks the sorted run in chunks of 32 (flag, inclusive scan, running carry)
    if (tid == 0) carry = 0;
    __syncthreads();
    if (tid < 32) {
        int base = 0;
        for (int c = 0; c < r; c += 32) {
            const int i = c + tid;
            const bool live = i < r;
            const int64_t x = live ? key[i] : 0;
            const int flag = live && (i == 0 || key[i - 1] != x) ? 1 : 0;
            int incl = flag;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(LEV_FULL_MASK, incl, o);
                if (tid >= o) incl += t;
            }
            const int rank = base + incl - 1;  // rank of x among the distinct values
            if (live) {
                uid[n * Rp + pos[i]] = rank;
                if (flag) dtok[n * Rp + rank] = x;
            }
            base += __shfl_sync(LEV_FULL_MASK, incl, 31);
        }
        if (tid == 0) ndist[n] = base;
    }
}

int lev_launch_uid(const b200lev_tokens_t* ref, const int32_t* packed, const int* state,
                   const int32_t* ref_len, int32_t* uid,
                   int64_t* dtok, int32_t* ndist, int64_t Rp, cudaStream_t st) {
    if (ref->N <= 0) return B200LEV_OK;
    int n_pad = 32;
    while (n_pad < ref->T) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * (sizeof(int64_t) + sizeof(int));
    if (smem > 200 * 1024) {
        lev_set_error("reference length %lld too long for the completion path", (long long)ref->T);
        return B200LEV_ERR_UNSUPPORTED;
    }
    dim3 grid((unsigned)ref->N), block(n_pad <= 256 ? 128 : 256);
    lev_prof_begin(LEV_PROF_COMP_UID, st);
#define LEV_UID_CASE(TT)                                                                        \
    {                                                                                           \
        auto kern = lev_uid_kernel<TT>;                                                         \
        if (smem > 48 * 1024)                                                                   \
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        lev_launch(kern, grid, block, smem, st, (const TT*)ref->data, ref->stride_t,            \
                   ref->stride_n, packed, state, ref_len, uid, dtok, ndist, Rp, n_pad);                       \
    }
    switch (ref->elem_bytes) {
        case 8: LEV_UID_CASE(int64_t) break;
        case 4: LEV_UID_CASE(int32_t) break;
        case 2: LEV_UID_CASE(int16_t) break;
        case 1: LEV_UID_CASE(int8_t) break;
        default:
            lev_set_error("unsupported token element size %d", (int)ref->elem_bytes);
            return B200LEV_ERR_ARG;
    }
#undef LEV_UID_CASE
    lev_prof_end(LEV_PROF_COMP_UID, st);
    return lev_check_cuda("lev_uid_kernel");
}

// One thread per (hypothesis prefix i, pair n) lists the distinct next tokens of its row.
// STAGED: the 32 rows of a warp are ONE contiguous run of 32 * U values in the output (rows of
// consecutive n are adjacent), most of it padding.  The warp keeps an image of that run in
// shared memory: it fills the image with `padding` (128-bit stores), each lane drops the few
// tokens of its row into it, and the image goes out with 128-bit loads and stores, 512
// contiguous bytes per instruction -- no per-element index arithmetic at all.  A thread writing
// its own U values directly touches 32 different sectors per instruction, a quarter of each:
// 2.4 TB/s of the 7 TB/s a plain fill reaches on this part.
template <bool STAGED>
__global__ void __launch_bounds__(256)
lev_completion_fill_kernel(const uint32_t* __restrict__ dbits, const int64_t* __restrict__ dtok,
                           const int* __restrict__ state, int64_t Rp, int Hout, int P, int Wd, int ref_group, int64_t U, int64_t padding,
                           int64_t* __restrict__ out, int64_t out_si, int64_t out_sn) {
    LEV_DYN_SMEM(int64_t, lev_fill_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // blockIdx.x: 256 consecutive pairs; blockIdx.y strides over the prefixes: a warp's 32 rows
    // always share their prefix, and nothing here divides
    const int n = blockIdx.x * 256 + threadIdx.x;
    const int n0 = n - lane;
    if (n0 >= P) return;
    const bool live = n < P;
    const int Ui = (int)U;
    const int total2 = 16 * Ui;  // 16-byte pieces of a run of 32 rows
    int64_t* __restrict__ img = lev_fill_smem + (int64_t)warp * 32 * U;
    // staged only when the warp's 32 rows are one contiguous piece of the output
    const bool run = STAGED && out_sn == U && n0 + 31 < P;
    const int64_t* __restrict__ dt = dtok + (int64_t)((live ? n : n0) / ref_group) * Rp;
    const longlong2 pad2 = make_longlong2(padding, padding);
    int tmin;
    const bool direct = lev_tokens_direct(state, &tmin);  // bit b of word 0 = token tmin + b
    // The two global reads of a row depend on each other (bitmap word -> token table) and both
    // miss L1: the first word of the NEXT prefix's bitmap is fetched an iteration ahead, and the
    // table look-ups go out four at a time before the first of them is used.
    uint32_t word0 = (live && (int)blockIdx.y < Hout) ? dbits[((int64_t)blockIdx.y * P + n) * Wd] : 0u;
    for (int i = blockIdx.y; i < Hout; i += gridDim.y) {
        const uint32_t w0 = word0;
        if (live && i + (int)gridDim.y < Hout) word0 = dbits[((int64_t)(i + gridDim.y) * P + n) * Wd];
        if (STAGED && run) {
            longlong2* __restrict__ img2 = reinterpret_cast<longlong2*>(img);
            for (int e = lane; e < total2; e += 32) img2[e] = pad2;
            __syncwarp();
        }
        int64_t* __restrict__ o = run ? img + lane * Ui : out + i * out_si + n * out_sn;
        if (live) {
            const uint32_t* __restrict__ w = dbits + ((int64_t)i * P + n) * Wd;
            int k = 0;
            for (int q = 0; q < Wd && k < Ui; ++q) {
                uint32_t bits = q == 0 ? w0 : w[q];
                const int64_t* __restrict__ dq = dt + q * 32;
                if (direct) {  // (q == 0: such bitmaps have one word in use)
                    while (bits && k < Ui) {
                        o[k++] = (int64_t)tmin + (__ffs((int)bits) - 1);
                        bits &= bits - 1u;
                    }
                }
                while (bits && k < Ui) {
                    int64_t t[4];
                    int nb = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool ok = bits != 0u;
                        const int b = ok ? __ffs((int)bits) - 1 : 0;
                        bits &= bits - 1u;  // (0 stays 0)
                        t[j] = dq[b];       // (slot 0 of the table always exists)
                        nb += ok ? 1 : 0;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j < nb && k < Ui) o[k++] = t[j];
                }
            }
            if (!run)
                for (; k < Ui; ++k) o[k] = padding;
        }
        if (STAGED && run) {
            __syncwarp();
            const int64_t first = i * out_si + (int64_t)n0 * U;
            if ((first & 1) == 0) {
                const longlong2* __restrict__ img2 = reinterpret_cast<const longlong2*>(img);
                longlong2* __restrict__ dst2 = reinterpret_cast<longlong2*>(out + first);
                for (int e = lane; e < total2; e += 32) dst2[e] = img2[e];
            } else {
                // the run starts on an odd element (odd U times odd row count): 8-byte stores
                int64_t* __restrict__ dst = out + first;
                for (int e = lane; e < 2 * total2; e += 32) dst[e] = img[e];
            }
            __syncwarp();  // the image is reused by the next prefix
        }
    }
}

int lev_launch_completion_fill(const uint32_t* dbits, const int64_t* dtok, const int* state, int64_t Rp,
                               int64_t Hout, int64_t P, int64_t Wd, int ref_group, int64_t U,
                               int64_t padding, int64_t* out, int64_t out_si, int64_t out_sn,
                               cudaStream_t st) {
    if (Hout <= 0 || P <= 0 || U <= 0) return B200LEV_OK;
    // (a CTA walks several prefixes once the grid holds a few waves)
    const int64_t gx = (P + 255) / 256;
    int64_t gy = (148 * 16 + gx - 1) / gx;
    if (gy > Hout) gy = Hout;
    if (gy > 65535) gy = 65535;
    dim3 block(256), grid((unsigned)gx, (unsigned)gy);
    const size_t smem = (size_t)8 * 32 * (size_t)U * sizeof(int64_t);
    lev_prof_begin(LEV_PROF_COMP_FILL, st);
    if (smem <= 96 * 1024) {
        auto kern = lev_completion_fill_kernel<true>;
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lev_launch(kern, grid, block, smem, st, dbits, dtok, state, Rp, (int)Hout, (int)P, (int)Wd, ref_group, U, padding,
                   out, out_si, out_sn);
    } else {
        lev_launch(lev_completion_fill_kernel<false>, grid, block, 0, st, dbits, dtok, state, Rp, (int)Hout, (int)P,
                   (int)Wd, ref_group, U, padding, out, out_si, out_sn);
    }
    lev_prof_end(LEV_PROF_COMP_FILL, st);
    return lev_check_cuda("lev_completion_fill_kernel");
}
