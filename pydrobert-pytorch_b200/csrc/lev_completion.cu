// lev_completion.cu -- K4: the sorted-unique compaction of optimal_completion.
//
// The reference (SM:492-517) turns the (H', R, N) position mask into per-prefix token
// sets by (i) OR-ing the mask across duplicate tokens with an (H', N, R, R) boolean
// temporary, (ii) sorting the reference tokens, (iii) masked_select/masked_scatter.
// Here the sort happens ONCE per reference instead of once per prefix:
//   lev_uid_kernel      : for every reference position j its rank uid[j] among the
//                         DISTINCT token values of that reference (ascending, as
//                         int64), plus the table dtok[rank] -> token value;
//   (lev_dp.cu, MASK)   : sets bit uid[j] of the (prefix, pair) bitmap for every
//                         flagged position -- duplicates collapse onto one bit and
//                         ascending bit order IS ascending token order;
//   lev_completion_fill : enumerates the set bits into the (H', N, U) int64 output,
//                         padding the tail (coalesced: consecutive threads own
//                         consecutive U-element rows).
#include "lev_common.cuh"

template <typename TT>
__global__ void __launch_bounds__(256)
lev_uid_kernel(const TT* __restrict__ tok, int64_t st, int64_t sn, const int32_t* __restrict__ ref_len,
               int32_t* __restrict__ uid, int64_t* __restrict__ dtok, int32_t* __restrict__ ndist,
               int64_t Rp, int64_t R) {
    LEV_DYN_SMEM(int64_t, t64);                                       // [R]
    unsigned char* isfirst = reinterpret_cast<unsigned char*>(t64 + R);  // [R]
    __shared__ int nfirst;
    const int64_t n = blockIdx.x;
    const int r = ref_len[n];
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) nfirst = 0;
    for (int j = tid; j < r; j += nt) t64[j] = (int64_t)tok[(int64_t)j * st + n * sn];
    __syncthreads();
    int mine = 0;
    for (int j = tid; j < r; j += nt) {
        const int64_t x = t64[j];
        int f = 1;
        for (int k = 0; k < j; ++k)
            if (t64[k] == x) {
                f = 0;
                break;
            }
        isfirst[j] = (unsigned char)f;
        mine += f;
    }
    if (mine) atomicAdd(&nfirst, mine);
    __syncthreads();
    for (int j = tid; j < r; j += nt) {
        const int64_t x = t64[j];
        int rank = 0;
        for (int k = 0; k < r; ++k) rank += (isfirst[k] && t64[k] < x) ? 1 : 0;
        uid[n * Rp + j] = rank;
        if (isfirst[j]) dtok[n * Rp + rank] = x;
    }
    if (tid == 0) ndist[n] = nfirst;
}

int lev_launch_uid(const b200lev_tokens_t* ref, const int32_t* ref_len, int32_t* uid,
                   int64_t* dtok, int32_t* ndist, int64_t Rp, cudaStream_t st) {
    if (ref->N <= 0) return B200LEV_OK;
    const size_t smem = (size_t)ref->T * (sizeof(int64_t) + 1) + 16;
    if (smem > 200 * 1024) {
        lev_set_error("reference length %lld too long for the completion path", (long long)ref->T);
        return B200LEV_ERR_UNSUPPORTED;
    }
    dim3 grid((unsigned)ref->N), block(256);
    lev_prof_begin(LEV_PROF_COMP_UID, st);
#define LEV_UID_CASE(TT)                                                                        \
    {                                                                                           \
        auto kern = lev_uid_kernel<TT>;                                                         \
        if (smem > 48 * 1024)                                                                   \
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        lev_launch(kern, grid, block, smem, st, (const TT*)ref->data, ref->stride_t,            \
                   ref->stride_n, ref_len, uid, dtok, ndist, Rp, ref->T);                       \
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
// STAGED: the 32 lists of a warp go through shared memory and leave as one contiguous
// run of 32 * U values (rows of consecutive n are adjacent in the output), so every store
// instruction of the warp fills whole sectors; a thread writing its own U values directly
// touches 32 different sectors per instruction, a quarter of each.
template <bool STAGED>
__global__ void __launch_bounds__(256)
lev_completion_fill_kernel(const uint32_t* __restrict__ dbits, const int64_t* __restrict__ dtok,
                           int64_t Rp, int64_t rows /* Hout*P */, int64_t P, int64_t Wd,
                           int ref_group, int64_t U, int64_t padding, int64_t* __restrict__ out,
                           int64_t out_si, int64_t out_sn) {
    LEV_DYN_SMEM(int64_t, lev_fill_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t idx0 = idx - lane;  // first row of this warp
    if (idx0 >= rows) return;
    const bool live = idx < rows;
    const int64_t i = idx / P, n = idx - i * P;
    const int64_t Us = U | 1;  // odd row pitch: the lanes' 8-byte stores spread over the banks
    int64_t* __restrict__ buf = lev_fill_smem + (int64_t)warp * 32 * Us;
    // staged only when the warp's 32 rows are one contiguous piece of the output
    const int64_t i0 = idx0 / P;
    const bool run = STAGED && out_sn == U && idx0 + 31 < rows && (idx0 + 31) / P == i0;
    int64_t* __restrict__ o = run ? buf + (int64_t)lane * Us : out + i * out_si + n * out_sn;
    if (live) {
        const uint32_t* __restrict__ w = dbits + idx * Wd;
        const int64_t* __restrict__ dt = dtok + (n / ref_group) * Rp;
        int64_t k = 0;
        for (int64_t q = 0; q < Wd && k < U; ++q) {
            uint32_t bits = w[q];
            while (bits && k < U) {
                const int b = __ffs((int)bits) - 1;
                bits &= bits - 1;
                o[k++] = dt[q * 32 + b];
            }
        }
        for (; k < U; ++k) o[k] = padding;
    }
    if (STAGED) {
        __syncwarp();
        if (run) {
            int64_t* __restrict__ dst = out + i0 * out_si + (idx0 - i0 * P) * U;
            int64_t row = 0, k = lane;
            while (k >= U) {
                k -= U;
                ++row;
            }
            for (int64_t e = lane; e < 32 * U; e += 32) {
                dst[e] = buf[row * Us + k];
                k += 32;
                while (k >= U) {
                    k -= U;
                    ++row;
                }
            }
        }
    }
}

int lev_launch_completion_fill(const uint32_t* dbits, const int64_t* dtok, int64_t Rp,
                               int64_t Hout, int64_t P, int64_t Wd, int ref_group, int64_t U,
                               int64_t padding, int64_t* out, int64_t out_si, int64_t out_sn,
                               cudaStream_t st) {
    const int64_t rows = Hout * P;
    if (rows <= 0 || U <= 0) return B200LEV_OK;
    dim3 block(256), grid((unsigned)((rows + 255) / 256));
    const size_t smem = (size_t)8 * 32 * (size_t)(U | 1) * sizeof(int64_t);
    lev_prof_begin(LEV_PROF_COMP_FILL, st);
    if (smem <= 96 * 1024) {
        auto kern = lev_completion_fill_kernel<true>;
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lev_launch(kern, grid, block, smem, st, dbits, dtok, Rp, rows, P, Wd, ref_group, U, padding, out,
                   out_si, out_sn);
    } else {
        lev_launch(lev_completion_fill_kernel<false>, grid, block, 0, st, dbits, dtok, Rp, rows, P, Wd,
                   ref_group, U, padding, out, out_si, out_sn);
    }
    lev_prof_end(LEV_PROF_COMP_FILL, st);
    return lev_check_cuda("lev_completion_fill_kernel");
}
