// lev_pack.cu -- K0: lengths from eos + narrow/transpose tokens into pair-major int32.
//
// Replaces SM:137-143 (_lens_from_eos), the include_eos adjustment of SM:195-228 and
// the transposes of SM:181-183.  One pass over the token tensor: a CTA takes 32
// sequences, loads 32x32 tiles with the unit-stride axis on threadIdx.x (256-byte
// coalesced rows for a sequence-first int64 tensor), finds the first eos per
// sequence with a shared-memory atomicMin, and writes the tile transposed so that
// every sequence becomes one contiguous int32 row (what the DP kernels and the TMA
// bulk copies of the long-pair kernel want).
//
// HBM traffic: T*N*elem_bytes read + T*N*4 written, both fully coalesced.
#include "lev_common.cuh"

template <typename TT>
__global__ void __launch_bounds__(256, 6)
lev_pack_kernel(const TT* __restrict__ tok, int64_t T, int64_t N, int64_t st, int64_t sn,
                int has_eos, int64_t eos, int include_eos, int32_t* __restrict__ packed,
                int64_t Tp, uint16_t* __restrict__ packed16, int64_t Tp16,
                int32_t* __restrict__ lens, int32_t* flags, int32_t* state,
                int missing_flag, int transposed, const int32_t* __restrict__ ref_len,
                int ref_group, int G, int* __restrict__ ghist) {
    __shared__ int tile[32][33];
    __shared__ int first[32];
    __shared__ int blk_flags;
    __shared__ unsigned blk_umax, blk_nmax;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t n0 = (int64_t)blockIdx.x * 32;
    if (ty == 0) first[tx] = (int)T;
    if (tx == 0 && ty == 0) {
        blk_flags = 0;
        blk_umax = 0u;
        blk_nmax = 0u;
    }
    __syncthreads();
    int wide = 0;
    int my_first = (int)T;  // first eos seen by this thread (eos-padded tails hit it often)
    // int32 range of the (truncated) tokens this thread saw; turned into the biased
    // max(u) / max(~u), u = tok + 2^31, that the packed 16-bit DP path tests
    int lo = 0x7fffffff, hi = (int)0x80000000;
    const int Ti = (int)T;
    if (transposed) {
        // software pipeline: the loads of tile k+1 are issued before the stores of tile k,
        // so the two directions of HBM traffic overlap inside one CTA; addresses advance
        // by pointer increments only
        const bool n_ok = (n0 + tx) < N;
        const TT* __restrict__ src = tok + (n0 + tx) * sn + (int64_t)ty * st;
        const int64_t st8 = 8 * st, st32 = 32 * st;
        TT regs[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) regs[q] = (n_ok && ty + 8 * q < Ti) ? src[q * st8] : (TT)0;
        int32_t* dst32[4];
        uint16_t* dst16[4];
        bool row_ok[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t n = n0 + ty + 8 * q;
            row_ok[q] = n < N;
            dst32[q] = packed + n * Tp + tx;
            dst16[q] = packed16 != nullptr ? packed16 + n * Tp16 + tx : nullptr;
        }
        for (int t0 = 0; t0 < Ti; t0 += 32) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int tl = ty + 8 * q, t = t0 + tl;
                if (n_ok && t < Ti) {
                    const int64_t v = (int64_t)regs[q];
                    const int v32 = (int)v;
                    tile[tl][tx] = v32;
                    if (has_eos && v == eos && t < my_first) my_first = t;
                    if (sizeof(TT) == 8 && (int64_t)v32 != v) wide = 1;
                    lo = v32 < lo ? v32 : lo;
                    hi = v32 > hi ? v32 : hi;
                }
            }
            __syncthreads();
            if (t0 + 32 < Ti) {
                src += st32;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    regs[q] = (n_ok && t0 + 32 + ty + 8 * q < Ti) ? src[q * st8] : (TT)0;
            }
            if (t0 + tx < Ti) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (row_ok[q]) {
                        const int v = tile[tx][ty + 8 * q];
                        dst32[q][t0] = v;
                        if (dst16[q] != nullptr) dst16[q][t0] = (uint16_t)v;
                    }
                }
            }
            __syncthreads();
        }
        if (my_first < Ti) atomicMin(&first[tx], my_first);
        __syncthreads();
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int nl = ty + 8 * q;
            const int64_t n = n0 + nl;
            if (n < N) {
                for (int64_t t = tx; t < T; t += 32) {
                    const int64_t v = (int64_t)tok[t * st + n * sn];
                    packed[n * Tp + t] = (int)v;
                    if (packed16 != nullptr) packed16[n * Tp16 + t] = (uint16_t)(int)v;
                    if (has_eos && v == eos) atomicMin(&first[nl], (int)t);
                    if ((int64_t)(int)v != v) wide = 1;
                    lo = (int)v < lo ? (int)v : lo;
                    hi = (int)v > hi ? (int)v : hi;
                }
            }
        }
        __syncthreads();
    }
    if (wide) atomicOr(&blk_flags, B200LEV_FLAG_WIDE_TOKENS);
    unsigned umax = (unsigned)hi + 0x80000000u, nmax = ~((unsigned)lo + 0x80000000u);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned a = __shfl_xor_sync(LEV_FULL_MASK, umax, o);
        const unsigned b = __shfl_xor_sync(LEV_FULL_MASK, nmax, o);
        umax = a > umax ? a : umax;
        nmax = b > nmax ? b : nmax;
    }
    if (tx == 0) {
        atomicMax(&blk_umax, umax);
        atomicMax(&blk_nmax, nmax);
    }
    if (ty == 0 && n0 + tx < N) {
        int len = first[tx];
        if (has_eos && include_eos) {  // SM:198-218
            if (len == (int)T)
                atomicOr(&blk_flags, missing_flag);
            else
                len += 1;
        }
        lens[n0 + tx] = len;
        // hypothesis side only: (class, length) histogram for the group kernel's global
        // bucketing (the reference lengths were produced by the launch before this one)
        if (ghist != nullptr)
            atomicAdd(&ghist[lev_group_bin(ref_len[(n0 + tx) / ref_group], len, G, (int)T)], 1);
    }
    __syncthreads();
    if (tx == 0 && ty == 0) {
        if (blk_flags != 0) {
            if (flags != nullptr) atomicOr(flags, blk_flags);
            atomicOr(state, blk_flags);
        }
        atomicMax(reinterpret_cast<unsigned*>(state) + 1, blk_umax);
        atomicMax(reinterpret_cast<unsigned*>(state) + 2, blk_nmax);
    }
}

int lev_launch_pack(const b200lev_tokens_t* t, int has_eos, int64_t eos, int include_eos,
                    int32_t* packed, int64_t Tp, uint16_t* packed16, int64_t Tp16, int32_t* lens,
                    int32_t* flags, int32_t* state, int missing_flag, const int32_t* ref_len,
                    int ref_group, int G, int* ghist, cudaStream_t st) {
    if (t->N <= 0) return B200LEV_OK;
    if (t->T >= (int64_t)1 << 30) {
        lev_set_error("sequence dimension %lld too long", (long long)t->T);
        return B200LEV_ERR_UNSUPPORTED;
    }
    dim3 block(32, 8, 1);
    dim3 grid((unsigned)((t->N + 31) / 32), 1, 1);
    // unit stride along the batch axis => transpose through shared memory; otherwise
    // (batch_first, or an arbitrary view) read along the sequence axis directly.
    const int transposed = (t->stride_n == 1 && t->stride_t != 1);
    switch (t->elem_bytes) {
        case 8:
            lev_launch(lev_pack_kernel<int64_t>, grid, block, 0, st, (const int64_t*)t->data, t->T,
                       t->N, t->stride_t, t->stride_n, has_eos, eos, include_eos, packed, Tp,
                       packed16, Tp16, lens, flags, state, missing_flag, transposed, ref_len,
                       ref_group, G, ghist);
            break;
        case 4:
            lev_launch(lev_pack_kernel<int32_t>, grid, block, 0, st, (const int32_t*)t->data, t->T,
                       t->N, t->stride_t, t->stride_n, has_eos, eos, include_eos, packed, Tp,
                       packed16, Tp16, lens, flags, state, missing_flag, transposed, ref_len,
                       ref_group, G, ghist);
            break;
        case 2:
            lev_launch(lev_pack_kernel<int16_t>, grid, block, 0, st, (const int16_t*)t->data, t->T,
                       t->N, t->stride_t, t->stride_n, has_eos, eos, include_eos, packed, Tp,
                       packed16, Tp16, lens, flags, state, missing_flag, transposed, ref_len,
                       ref_group, G, ghist);
            break;
        case 1:
            lev_launch(lev_pack_kernel<int8_t>, grid, block, 0, st, (const int8_t*)t->data, t->T,
                       t->N, t->stride_t, t->stride_n, has_eos, eos, include_eos, packed, Tp,
                       packed16, Tp16, lens, flags, state, missing_flag, transposed, ref_len,
                       ref_group, G, ghist);
            break;
        default:
            lev_set_error("unsupported token element size %d", (int)t->elem_bytes);
            return B200LEV_ERR_ARG;
    }
    return lev_check_cuda("lev_pack_kernel");
}
