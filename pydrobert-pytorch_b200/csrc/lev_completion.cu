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
// consecutive n are adjacent), most of it padding.  Each lane enumerates the few set bits of its
// row into shared memory and publishes the count; then the warp writes the run with 128-bit
// stores, 512 contiguous bytes per store instruction (a value is the staged token if its slot
// is below the row's count, else `padding`, which never passes through shared memory).  A
// thread writing its own U values directly touches 32 different sectors per instruction, a
// quarter of each: 2.4 TB/s of the 7 TB/s a plain fill reaches on this part.
template <bool STAGED>
__global__ void __launch_bounds__(256)
lev_completion_fill_kernel(const uint32_t* __restrict__ dbits, const int64_t* __restrict__ dtok,
                           int64_t Rp, int Hout, int P, int Wd, int ref_group, int64_t U, int64_t padding,
                           int64_t* __restrict__ out, int64_t out_si, int64_t out_sn) {
    LEV_DYN_SMEM(int64_t, lev_fill_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // blockIdx.x: 256 consecutive pairs; blockIdx.y strides over the prefixes: a warp's 32 rows
    // always share their prefix, and nothing here divides
    const int n = blockIdx.x * 256 + threadIdx.x;
    const int n0 = n - lane;
    if (n0 >= P) return;
    const bool live = n < P;
    const int64_t Us = U | 1;  // odd row pitch: the lanes' 8-byte stores spread over the banks
    int64_t* __restrict__ buf = lev_fill_smem + (int64_t)warp * 32 * Us;
    // staged only when the warp's 32 rows are one contiguous piece of the output
    const bool run = STAGED && out_sn == U && n0 + 31 < P;
    const int64_t* __restrict__ dt = dtok + (int64_t)((live ? n : n0) / ref_group) * Rp;
    for (int i = blockIdx.y; i < Hout; i += gridDim.y) {
    int64_t* __restrict__ o = run ? buf + (int64_t)lane * Us : out + i * out_si + n * out_sn;
    int cnt = 0;
    if (live) {
        const uint32_t* __restrict__ w = dbits + ((int64_t)i * P + n) * Wd;
        int k = 0;
        const int Ui = (int)U;
        for (int q = 0; q < Wd && k < Ui; ++q) {
            uint32_t bits = w[q];
            while (bits && k < Ui) {
                const int b = __ffs((int)bits) - 1;
                bits &= bits - 1;
                o[k++] = dt[q * 32 + b];
            }
        }
        cnt = k;
        if (!run)
            for (; k < Ui; ++k) o[k] = padding;
    }
    if (STAGED) {
        __syncwarp();
        if (run) {
            int64_t* __restrict__ dst = out + i * out_si + (int64_t)n0 * U;
            // element e of the run = slot k of row `row`; lane handles e = 2 lane, 2 lane + 1,
            // then + 64 ...; the counts of the rows travel by shuffle.  All of it in 32 bits
            // (a run is 32 U elements) with the divisions by U hoisted out of the loop.
            const int Ui = (int)U, Usi = (int)Us, total = 32 * Ui;
            if (((i * out_si + (int64_t)n0 * U) & 1) != 0) {
                // the run starts on an odd element (odd U times odd row count): 8-byte stores
                for (int e = lane; e < total; e += 32) {
                    const int rw = e / Ui, kk = e - rw * Ui;
                    const int c = __shfl_sync(__activemask(), cnt, rw);
                    dst[e] = kk < c ? buf[rw * Usi + kk] : padding;
                }
                __syncwarp();
                continue;
            }
            const int q64 = 64 / Ui, r64 = 64 - q64 * Ui;
            int row = (2 * lane) / Ui, k = 2 * lane - row * Ui;
            for (int e = 2 * lane; e - 2 * lane < total; e += 64) {
                const bool in = e < total;  // (total is even: a pair never straddles the end)
                const int r0 = in ? row : 0;
                int r1 = r0, k1 = k + 1;
                if (k1 >= Ui) {
                    k1 = 0;
                    r1 = r0 + 1;
                }
                r1 = r1 > 31 ? 31 : r1;  // (only the unused second half of a final pair)
                const int c0 = __shfl_sync(LEV_FULL_MASK, cnt, r0);
                const int c1 = __shfl_sync(LEV_FULL_MASK, cnt, r1);
                if (in) {
                    longlong2 v;
                    v.x = k < c0 ? buf[r0 * Usi + k] : padding;
                    v.y = k1 < c1 ? buf[r1 * Usi + k1] : padding;
                    *reinterpret_cast<longlong2*>(dst + e) = v;
                }
                k += r64;
                row += q64;
                if (k >= Ui) {
                    k -= Ui;
                    ++row;
                }
            }
        }
        __syncwarp();  // the staging rows are reused by the next prefix
    }
    }
}

int lev_launch_completion_fill(const uint32_t* dbits, const int64_t* dtok, int64_t Rp,
                               int64_t Hout, int64_t P, int64_t Wd, int ref_group, int64_t U,
                               int64_t padding, int64_t* out, int64_t out_si, int64_t out_sn,
                               cudaStream_t st) {
    if (Hout <= 0 || P <= 0 || U <= 0) return B200LEV_OK;
    dim3 block(256), grid((unsigned)((P + 255) / 256), (unsigned)(Hout < 65535 ? Hout : 65535));
    const size_t smem = (size_t)8 * 32 * (size_t)(U | 1) * sizeof(int64_t);
    lev_prof_begin(LEV_PROF_COMP_FILL, st);
    if (smem <= 96 * 1024) {
        auto kern = lev_completion_fill_kernel<true>;
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lev_launch(kern, grid, block, smem, st, dbits, dtok, Rp, (int)Hout, (int)P, (int)Wd, ref_group, U, padding,
                   out, out_si, out_sn);
    } else {
        lev_launch(lev_completion_fill_kernel<false>, grid, block, 0, st, dbits, dtok, Rp, (int)Hout, (int)P,
                   (int)Wd, ref_group, U, padding, out, out_si, out_sn);
    }
    lev_prof_end(LEV_PROF_COMP_FILL, st);
    return lev_check_cuda("lev_completion_fill_kernel");
}
