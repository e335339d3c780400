// lev_pack.cu -- K0: lengths from eos + narrow/transpose tokens into pair-major tables.
//
// Replaces SM:137-143 (_lens_from_eos), the include_eos adjustment of SM:195-228 and
// the transposes of SM:181-183.  One pass over the token tensor, HBM-bound:
//   read  T*N*elem_bytes            (coalesced: 32 lanes x 8 B per load instruction)
//   write T*N*4  [+ T*N*2]          (int32 rows [+ uint16 rows for the packed DP path])
//
// Sequence-first tensors (unit stride along the batch axis): a CTA owns 32 sequences; lane l
// of every warp loads tok[t][n0 + l] -- each load instruction is one coalesced 256-byte row,
// 16 of them in flight per lane -- so a lane holds consecutive tokens of ITS OWN sequence in
// registers and the transpose is free.  The four warps take four 32-position slices of a
// 128-position chunk, stage them in a shared [32][132] tile and write whole rows back as
// 128-bit vectors: the 32 rows of a CTA are one contiguous block of each pair-major table
// (scripts/micro/pack.cu: per-lane row stores reach 66 us on the cfg2 hypothesis tensor,
// this layout 40 us, a plain streaming copy of the same bytes 35 us).  First-eos position,
// the "fits in int32" flag and the token range stay per-lane scalars: no atomics in the loop.
//
// Other layouts (batch-first, arbitrary views): one warp per sequence, lanes along the
// sequence axis (coalesced when that axis has unit stride).
//
// Side products: warning / width flags (B200LEV_FLAG_*), the biased token range
// [max(u), max(~u)], u = tok + 2^31, that lets the DP kernels pick the packed 16-bit path
// on the device, and -- on the hypothesis side -- the (column class, length) histogram of
// the group kernel's bucketing.
#include "lev_common.cuh"

struct LevPackArgs {
    const void* tok;
    int64_t T, N, st, sn;
    int has_eos;
    int64_t eos;
    int include_eos;
    int32_t* packed;
    int64_t Tp;
    uint16_t* packed16;  // may be NULL
    int64_t Tp16;
    int32_t* lens;
    int32_t* flags;  // caller's warning flags, may be NULL
    int32_t* state;  // workspace state words: [0] flags, [1] max(u), [2] max(~u)
    int missing_flag;
    const int32_t* ref_len;  // histogram side product (hypothesis pass only)
    int ref_group, G;
    int* ghist;
    int bv_check;  // exit if the bit-vector kernels, enqueued first, took the batch
    int slice;     // positions per warp and chunk: 32, or 16 / 8 for T <= 64 / 32
    int hist_bins;  // > 0: a CTA walks many blocks and collects the histogram in shared memory
    // few, long sequences (N / 32 CTAs would leave most SMs idle): `split` CTAs share a block of
    // 32 sequences, each taking split_len positions (a multiple of the 128-position chunk).
    // They meet in split_first[n] (first eos - T, atomicMin on zeroes) and split_ticket[block];
    // the CTA that draws the last ticket runs the owner epilogue.  split = 1: off.
    int split, split_len;
    int* split_first;
    int* split_ticket;
};

// ---- epilogue pieces shared by both layouts ----------------------------------------------
// owner lane of a sequence: length rule (SM:198-218) and, on the hypothesis side, the
// (class, length) histogram for the group kernel's global bucketing (the reference lengths
// were produced by the launch before this one).  Returns the warning flags it raises.
__device__ __forceinline__ int lev_pack_owner(const LevPackArgs& a, int ref_len, int64_t n,
                                              int first, int* cta_hist = nullptr) {
    int len = first, flags = 0;
    if (a.has_eos && a.include_eos) {
        if (len == (int)a.T)
            flags = a.missing_flag;
        else
            len += 1;
    }
    a.lens[n] = len;
    if (a.ghist != nullptr)
        atomicAdd(&(cta_hist ? cta_hist : a.ghist)[lev_group_bin(ref_len, len, a.G, (int)a.T)], 1);
    return flags;
}

// (flags, max(u), max(~u)) over the warp, valid in every lane
__device__ __forceinline__ void lev_pack_warp_reduce(int& flags, unsigned& umax, unsigned& nmax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned x = __shfl_xor_sync(LEV_FULL_MASK, umax, o);
        const unsigned y = __shfl_xor_sync(LEV_FULL_MASK, nmax, o);
        const int f = __shfl_xor_sync(LEV_FULL_MASK, flags, o);
        umax = x > umax ? x : umax;
        nmax = y > nmax ? y : nmax;
        flags |= f;
    }
}

// One thread publishes.  The range words only grow, so a recent copy (cur_u, cur_n) tells
// whether the reduction can be skipped: thousands of reductions on one address would
// otherwise queue behind each other at the end of the kernel.
__device__ __forceinline__ void lev_pack_publish(const LevPackArgs& a, int flags, unsigned umax,
                                                 unsigned nmax, unsigned cur_u, unsigned cur_n) {
    if (flags != 0) {
        if (a.flags != nullptr) atomicOr(a.flags, flags);
        atomicOr(a.state, flags);
    }
    if (umax > cur_u) atomicMax(reinterpret_cast<unsigned*>(a.state) + 1, umax);
    if (nmax > cur_n) atomicMax(reinterpret_cast<unsigned*>(a.state) + 2, nmax);
}

#define LEV_PACK_OBSERVE(x, t)                                       \
    {                                                                \
        const int v32_ = (int)(x);                                   \
        if (sizeof(TT) == 8 && (int64_t)v32_ != (x)) wide = 1;       \
        if (a.has_eos && (x) == a.eos && (t) < first) first = (t);   \
        lo = v32_ < lo ? v32_ : lo;                                  \
        hi = v32_ > hi ? v32_ : hi;                                  \
    }

// ---- sequence-first: a CTA owns 32 sequences, chunks of 128 positions -------------------
// Warp w loads positions [tb + 32w, tb + 32w + 32) of the chunk, lane l the sequence n0 + l:
// every load instruction is one coalesced row of 32 tokens, NB of them in flight per lane.
// The lane's consecutive positions go into its row of the shared tile as 128-bit stores (row
// stride 132 words: conflict-free); after the barrier whole rows leave as contiguous 16-byte
// vectors -- the CTA's 32 output rows are one contiguous block of each table.
constexpr int LEV_PACK_CHUNK = 128;
constexpr int LEV_PACK_STRIDE = LEV_PACK_CHUNK + 4;
constexpr int LEV_PACK_NB = 16;  // loads in flight per lane

// What a lane learns from the tokens it moves, at ~5 instructions per token:
//   first  first position whose LOW word equals eos32 (exact unless the lane saw a token that
//          does not fit in int32, or eos itself does not fit: the caller rescans then)
//   wacc   nonzero iff some token does not fit in int32
//   lo/hi  range of the low words
struct LevPackSeen {
    int first, wacc, lo, hi;
};

// One SW-position slice of one sequence: batches of NB coalesced loads, tokens narrowed into
// the lane's row of the shared tile.  GUARD adds the t < T predicate (last slice only).
// SW = 32 for long sequences; 16 / 8 keep all four warps busy when T <= 64 / 32.
template <typename TT, bool GUARD, int SW>
__device__ __forceinline__ void lev_pack_slice(const TT* __restrict__ src, int st, int t0,
                                               int Ti, bool eos_fast, int eos32, int* trow,
                                               LevPackSeen& seen) {
    constexpr int NB = SW < LEV_PACK_NB ? SW : LEV_PACK_NB;
    LEV_OPAQUE_PTR(src);
#pragma unroll
    for (int h = 0; h < SW / NB; ++h) {
        TT raw[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u)
            raw[u] = (!GUARD || t0 + h * NB + u < Ti)
                         ? lev_ldg_stream(src + (int64_t)(h * NB + u) * (int64_t)st)
                         : (TT)0;
        int v[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int64_t x = (int64_t)raw[u];
            v[u] = (int)x;
            if (sizeof(TT) == 8) seen.wacc |= (int)(x >> 32) ^ (v[u] >> 31);
            const int t = t0 + h * NB + u;
            if (GUARD) {
                if (t < Ti) {
                    if (eos_fast && v[u] == eos32) seen.first = t < seen.first ? t : seen.first;
                    seen.lo = v[u] < seen.lo ? v[u] : seen.lo;
                    seen.hi = v[u] > seen.hi ? v[u] : seen.hi;
                }
            } else {
                if (eos_fast && v[u] == eos32) seen.first = t < seen.first ? t : seen.first;
            }
        }
        if (!GUARD) {
#pragma unroll
            for (int u = 0; u < NB; u += 2) {
                seen.lo = __vimin3_s32(seen.lo, v[u], v[u + 1]);
                seen.hi = __vimax3_s32(seen.hi, v[u], v[u + 1]);
            }
        }
#pragma unroll
        for (int c = 0; c < NB / 4; ++c)
            *reinterpret_cast<int4*>(trow + h * NB + 4 * c) =
                make_int4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    }
}

template <typename TT>
__global__ void __launch_bounds__(128, 8) lev_pack_seqfirst_kernel(const LevPackArgs a) {
    if (a.bv_check && lev_bv_took(a.state)) return;
    __shared__ __align__(16) int tile[32][LEV_PACK_STRIDE];
    __shared__ int first_s[32];
    __shared__ unsigned blk_u, blk_n;
    __shared__ int blk_flags;
    // A million sequences fall into a few hundred (class, length) bins: one global atomic per
    // sequence queues up behind the others on the same address (+120 us at 1 M x T=31).  A CTA
    // that walks many blocks counts in shared memory and adds each bin once at the end.
    LEV_DYN_SMEM(int, cta_hist);
    for (int b = threadIdx.x; b < a.hist_bins; b += blockDim.x) cta_hist[b] = 0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // a CTA walks blocks of 32 sequences (a few each: the grid stays small, so that standing by
    // for the bit-vector path costs a 1000-CTA launch and not one of N / 32 CTAs)
    const int64_t nitems = ((a.N + 31) / 32) * a.split;
    for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int64_t n0 = (item / a.split) * 32;
    const int t_lo = (int)(item % a.split) * a.split_len;
    const int64_t n = n0 + lane;
    const bool valid_seq = n < a.N;
    const int rows = (int)(a.N - n0 < 32 ? a.N - n0 : 32);
    const int Ti = (int)a.T, Tp = (int)a.Tp, Tp16 = (int)a.Tp16;
    const int st = (int)a.st;  // the launcher routes strides beyond int32 to the rows kernel
    // lanes past the batch re-read sequence 0: its tokens are legitimate members of the
    // tensor's range, and such a lane owns no output
    const TT* __restrict__ seq = reinterpret_cast<const TT*>(a.tok) + (valid_seq ? n : 0);
    const int eos32 = (int)a.eos;
    const bool eos_fast = a.has_eos && (int64_t)eos32 == a.eos;
    LevPackSeen seen = {Ti, 0, 0x7fffffff, (int)0x80000000};
    // fetched before the token stream so that the latency hides under it
    const int ref_len = (w == 0 && valid_seq && a.ghist != nullptr) ? a.ref_len[n / a.ref_group] : 0;
    unsigned cur_u = 0u, cur_n = 0u;
    if (w == 0) first_s[lane] = Ti;
    if (threadIdx.x == 0) {
        blk_u = 0u;
        blk_n = 0u;
        blk_flags = 0;
    }
    __syncthreads();
    const int sw = a.slice, chunk = 4 * sw;
    const int t_hi = (a.split > 1 && t_lo + a.split_len < Ti) ? t_lo + a.split_len : Ti;
    for (int tb = t_lo; tb < t_hi; tb += chunk) {
        const int t0 = tb + sw * w;
        const TT* sp = seq + (int64_t)t0 * st;
        int* trow = &tile[lane][sw * w];
        const bool full = t0 + sw <= Ti;
#define LEV_PACK_SLICE(SW_)                                                                      \
    if (full)                                                                                    \
        lev_pack_slice<TT, false, SW_>(sp, st, t0, Ti, eos_fast, eos32, trow, seen);             \
    else                                                                                         \
        lev_pack_slice<TT, true, SW_>(sp, st, t0, Ti, eos_fast, eos32, trow, seen);
        if (sw == 32) {
            LEV_PACK_SLICE(32)
        } else if (sw == 16) {
            LEV_PACK_SLICE(16)
        } else {
            LEV_PACK_SLICE(8)
        }
#undef LEV_PACK_SLICE
        __syncthreads();
        if (threadIdx.x == 0 && tb + chunk >= t_hi) {
            // a recent copy of the range words; the load completes under the row stores
            cur_u = lev_ldg_l2(reinterpret_cast<const unsigned*>(a.state) + 1);
            cur_n = lev_ldg_l2(reinterpret_cast<const unsigned*>(a.state) + 2);
        }
        // rows are 16-byte aligned and padded to a multiple of 4 (8 for the 16-bit table);
        // positions T..Tp-1 receive zeros
        const int width = Tp - tb < chunk ? Tp - tb : chunk;
        const int width16 = Tp16 - tb < chunk ? Tp16 - tb : chunk;
        for (int r = w; r < rows; r += 4) {
            int32_t* __restrict__ row32 = a.packed + (n0 + r) * a.Tp + tb;
            for (int c = lane; 4 * c < width; c += 32)
                *reinterpret_cast<int4*>(row32 + 4 * c) = *reinterpret_cast<const int4*>(&tile[r][4 * c]);
            if (a.packed16 != nullptr) {
                uint16_t* __restrict__ row16 = a.packed16 + (n0 + r) * a.Tp16 + tb;
                for (int c = lane; 8 * c < width16; c += 32) {
                    const int4 x = *reinterpret_cast<const int4*>(&tile[r][8 * c]);
                    const int4 y = *reinterpret_cast<const int4*>(&tile[r][8 * c + 4]);
                    *reinterpret_cast<uint4*>(row16 + 8 * c) =
                        make_uint4(__byte_perm(x.x, x.y, 0x5410), __byte_perm(x.z, x.w, 0x5410),
                                   __byte_perm(y.x, y.y, 0x5410), __byte_perm(y.z, y.w, 0x5410));
                }
            }
        }
        __syncthreads();
    }
    if (a.has_eos && (!eos_fast || (sizeof(TT) == 8 && seen.wacc != 0))) {
        // rare: a low-word match may be a wide token (or eos itself is wide) -- rescan this
        // lane's slices with the exact comparison
        seen.first = Ti;
        for (int tb = t_lo; tb < t_hi; tb += chunk)
            for (int t = tb + sw * w; t < Ti && t < tb + sw * w + sw; ++t)
                if ((int64_t)seq[(int64_t)t * st] == a.eos) {
                    seen.first = t < seen.first ? t : seen.first;
                    break;
                }
    }
    if (seen.first < Ti) atomicMin(&first_s[lane], seen.first);
    {
        int flags = seen.wacc != 0 ? B200LEV_FLAG_WIDE_TOKENS : 0;
        unsigned umax = (unsigned)seen.hi + 0x80000000u, nmax = ~((unsigned)seen.lo + 0x80000000u);
        lev_pack_warp_reduce(flags, umax, nmax);
        if (lane == 0) {
            atomicMax(&blk_u, umax);
            atomicMax(&blk_n, nmax);
            if (flags != 0) atomicOr(&blk_flags, flags);
        }
    }
    __syncthreads();
    if (w == 0) {
        int first = first_s[lane];
        bool owner = valid_seq;
        if (a.split > 1) {
            if (valid_seq && first < Ti) atomicMin(&a.split_first[n], first - Ti);
            __threadfence();
            __syncwarp();
            int ticket = 0;
            if (lane == 0) ticket = atomicAdd(&a.split_ticket[n0 >> 5], 1);
            ticket = __shfl_sync(LEV_FULL_MASK, ticket, 0);
            owner = owner && ticket == a.split - 1;  // every other slice of these rows has published
            if (owner) {
                __threadfence();
                first = Ti + __ldcg(&a.split_first[n]);
            }
        }
        int flags = owner ? lev_pack_owner(a, ref_len, n, first, a.hist_bins ? cta_hist : nullptr) : 0;
        unsigned umax = blk_u, nmax = blk_n;
        lev_pack_warp_reduce(flags, umax, nmax);
        if (lane == 0) lev_pack_publish(a, flags | blk_flags, umax, nmax, cur_u, cur_n);
    }
    __syncthreads();  // the next block re-initialises the shared scalars
    }
    for (int b = threadIdx.x; b < a.hist_bins; b += blockDim.x)
        if (cta_hist[b] != 0) atomicAdd(&a.ghist[b], cta_hist[b]);
}

// ---- sequence-first, short sequences (T <= 64): a WARP owns 32 sequences ------------------
// With only one or two 32-position slices per sequence the four-warp chunk above leaves warps
// idle and pays two block barriers per 1 K tokens (cfg4: 1 M sequences of 31).  Here every warp
// runs its own 32 sequences start to finish -- slice loads, its private [32][36] tile, row
// stores that cover 4 (int32) or 8 (uint16) rows per instruction, lane-local lengths -- and
// only __syncwarp stands between the phases.
template <typename TT>
__global__ void __launch_bounds__(128, 8) lev_pack_warpseq_kernel(const LevPackArgs a) {
    if (a.bv_check && lev_bv_took(a.state)) return;
    __shared__ __align__(16) int tiles[4][32][36];
    LEV_DYN_SMEM(int, cta_hist);
    for (int b = threadIdx.x; b < a.hist_bins; b += blockDim.x) cta_hist[b] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int(*tile)[36] = tiles[w];
    const int Ti = (int)a.T, Tp = (int)a.Tp, Tp16 = (int)a.Tp16;
    const int st = (int)a.st;
    const int eos32 = (int)a.eos;
    const bool eos_fast = a.has_eos && (int64_t)eos32 == a.eos;
    for (int64_t n0 = ((int64_t)blockIdx.x * 4 + w) * 32; n0 < a.N; n0 += (int64_t)gridDim.x * 128) {
        const int64_t n = n0 + lane;
        const bool valid_seq = n < a.N;
        const int rows = (int)(a.N - n0 < 32 ? a.N - n0 : 32);
        const TT* __restrict__ seq = reinterpret_cast<const TT*>(a.tok) + (valid_seq ? n : 0);
        LevPackSeen seen = {Ti, 0, 0x7fffffff, (int)0x80000000};
        const int ref_len = (valid_seq && a.ghist != nullptr) ? a.ref_len[n / a.ref_group] : 0;
        unsigned cur_u = 0u, cur_n = 0u;
        for (int tb = 0; tb < Ti; tb += 32) {
            if (tb + 32 <= Ti)
                lev_pack_slice<TT, false, 32>(seq + (int64_t)tb * st, st, tb, Ti, eos_fast, eos32, &tile[lane][0], seen);
            else
                lev_pack_slice<TT, true, 32>(seq + (int64_t)tb * st, st, tb, Ti, eos_fast, eos32, &tile[lane][0], seen);
            __syncwarp();
            if (lane == 0 && tb + 32 >= Ti) {  // a recent copy of the range words (see above)
                cur_u = lev_ldg_l2(reinterpret_cast<const unsigned*>(a.state) + 1);
                cur_n = lev_ldg_l2(reinterpret_cast<const unsigned*>(a.state) + 2);
            }
            const int cpr = (Tp - tb < 32 ? Tp - tb : 32) >> 2;  // 16-byte chunks per int32 row
            for (int i = lane; i < rows * cpr; i += 32) {
                const int r = i / cpr, c = i - r * cpr;
                *reinterpret_cast<int4*>(a.packed + (n0 + r) * a.Tp + tb + 4 * c) =
                    *reinterpret_cast<const int4*>(&tile[r][4 * c]);
            }
            if (a.packed16 != nullptr) {
                const int cpr16 = (Tp16 - tb < 32 ? Tp16 - tb : 32) >> 3;
                for (int i = lane; i < rows * cpr16; i += 32) {
                    const int r = i / cpr16, c = i - r * cpr16;
                    const int4 x = *reinterpret_cast<const int4*>(&tile[r][8 * c]);
                    const int4 y = *reinterpret_cast<const int4*>(&tile[r][8 * c + 4]);
                    *reinterpret_cast<uint4*>(a.packed16 + (n0 + r) * a.Tp16 + tb + 8 * c) =
                        make_uint4(__byte_perm(x.x, x.y, 0x5410), __byte_perm(x.z, x.w, 0x5410),
                                   __byte_perm(y.x, y.y, 0x5410), __byte_perm(y.z, y.w, 0x5410));
                }
            }
            __syncwarp();
        }
        if (a.has_eos && (!eos_fast || (sizeof(TT) == 8 && seen.wacc != 0))) {
            seen.first = Ti;  // rare: exact rescan (wide tokens / wide eos)
            for (int t = 0; t < Ti; ++t)
                if ((int64_t)seq[(int64_t)t * st] == a.eos) {
                    seen.first = t;
                    break;
                }
        }
        int flags = seen.wacc != 0 ? B200LEV_FLAG_WIDE_TOKENS : 0;
        if (valid_seq) flags |= lev_pack_owner(a, ref_len, n, seen.first, a.hist_bins ? cta_hist : nullptr);
        unsigned umax = (unsigned)seen.hi + 0x80000000u, nmax = ~((unsigned)seen.lo + 0x80000000u);
        lev_pack_warp_reduce(flags, umax, nmax);
        if (lane == 0) lev_pack_publish(a, flags, umax, nmax, cur_u, cur_n);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < a.hist_bins; b += blockDim.x)
        if (cta_hist[b] != 0) atomicAdd(&a.ghist[b], cta_hist[b]);
}

// ---- any other layout: one warp per sequence, lanes along the sequence axis -------------
template <typename TT>
__global__ void __launch_bounds__(128) lev_pack_rows_kernel(const LevPackArgs a) {
    if (a.bv_check && lev_bv_took(a.state)) return;
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const bool valid_seq = n < a.N;  // warp-uniform
    const int Ti = (int)a.T;
    int first = Ti, wide = 0, lo = 0x7fffffff, hi = (int)0x80000000;
    const int ref_len = (valid_seq && lane == 0 && a.ghist != nullptr) ? a.ref_len[n / a.ref_group] : 0;
    if (valid_seq) {
        const TT* __restrict__ src = reinterpret_cast<const TT*>(a.tok) + n * a.sn;
        int32_t* __restrict__ row32 = a.packed + n * a.Tp;
        uint16_t* __restrict__ row16 = a.packed16 ? a.packed16 + n * a.Tp16 : nullptr;
        for (int t = lane; t < Ti; t += 32) {
            const int64_t x = (int64_t)src[(int64_t)t * a.st];
            row32[t] = (int)x;
            if (row16 != nullptr) row16[t] = (uint16_t)(int)x;
            LEV_PACK_OBSERVE(x, t)
        }
    }
    // first eos of the sequence = minimum over the lanes
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int f = __shfl_xor_sync(LEV_FULL_MASK, first, o);
        first = f < first ? f : first;
    }
    int flags = wide ? B200LEV_FLAG_WIDE_TOKENS : 0;
    if (valid_seq && lane == 0) flags |= lev_pack_owner(a, ref_len, n, first);
    unsigned umax = (unsigned)hi + 0x80000000u, nmax = ~((unsigned)lo + 0x80000000u);
    lev_pack_warp_reduce(flags, umax, nmax);
    if (lane == 0)
        lev_pack_publish(a, flags, umax, nmax, lev_ldg_l2(reinterpret_cast<const unsigned*>(a.state) + 1),
                         lev_ldg_l2(reinterpret_cast<const unsigned*>(a.state) + 2));
}

int lev_launch_pack(const b200lev_tokens_t* t, int has_eos, int64_t eos, int include_eos,
                    int32_t* packed, int64_t Tp, uint16_t* packed16, int64_t Tp16, int32_t* lens,
                    int32_t* flags, int32_t* state, int missing_flag, const int32_t* ref_len,
                    int ref_group, int G, int* ghist, int bv_check, cudaStream_t st, int32_t* split_scratch) {
    if (t->N <= 0) return B200LEV_OK;
    if (t->T >= (int64_t)1 << 30) {
        lev_set_error("sequence dimension %lld too long", (long long)t->T);
        return B200LEV_ERR_UNSUPPORTED;
    }
    LevPackArgs a;
    a.tok = t->data;
    a.T = t->T;
    a.N = t->N;
    a.st = t->stride_t;
    a.sn = t->stride_n;
    a.has_eos = has_eos;
    a.eos = eos;
    a.include_eos = include_eos;
    a.packed = packed;
    a.Tp = Tp;
    a.packed16 = packed16;
    a.Tp16 = Tp16;
    a.lens = lens;
    a.flags = flags;
    a.state = state;
    a.missing_flag = missing_flag;
    a.ref_len = ref_len;
    a.ref_group = ref_group < 1 ? 1 : ref_group;
    a.G = G;
    a.ghist = ghist;
    a.bv_check = bv_check;
    a.slice = t->T <= 32 ? 8 : (t->T <= 64 ? 16 : 32);
    // unit stride along the batch axis => register-tile transpose; otherwise (batch_first,
    // or an arbitrary view) one warp per sequence.
    const bool seqfirst = (t->stride_n == 1 && t->stride_t != 1 && t->stride_t > -((int64_t)1 << 31) &&
                           t->stride_t < ((int64_t)1 << 31));
    const dim3 block(128, 1, 1);
    int64_t nblk = seqfirst ? (t->N + 31) / 32 : (t->N + 3) / 4;
    int64_t cap = 148 * 8;
    if (const char* e = getenv("B200LEV_PACK_CTAS")) cap = atoll(e) > 0 ? atoll(e) : cap;  // tests
    a.hist_bins = 0;
    size_t smem = 0;
    // few CTAs walking many blocks: on stand-by duty (a small grid exits faster), and for big
    // batches, where the histogram is then collected per CTA
    if (seqfirst && nblk > cap && (bv_check || nblk >= 8 * cap)) {
        const int64_t per_cta = (nblk + cap - 1) / cap;
        nblk = (nblk + per_cta - 1) / per_cta;
        const int64_t bins = (int64_t)LEV_GROUP_NCLS * (t->T + 1);
        if (ghist != nullptr && per_cta >= 4 && bins <= 8192) {
            a.hist_bins = (int)bins;
            smem = sizeof(int) * (size_t)bins;
        }
    }
    const bool warpseq = seqfirst && t->T <= 64;
    // few, long sequences: several CTAs per block of 32 sequences, split along the sequence axis
    a.split = 1;
    a.split_len = 0;
    a.split_first = a.split_ticket = nullptr;
    {
        int64_t want = 148 * 4;  // CTAs that keep the machine busy
        if (const char* e = getenv("B200LEV_PACK_SPLIT")) want = atoll(e);  // tests; 0 = off
        const int64_t nchunks = (t->T + LEV_PACK_CHUNK - 1) / LEV_PACK_CHUNK;
        if (seqfirst && !warpseq && split_scratch != nullptr && nchunks >= 2 && 4 * nblk <= want) {  // (less than a CTA per SM)
            int64_t S = (want + nblk - 1) / nblk;
            S = S < nchunks ? S : nchunks;
            const int64_t per = (nchunks + S - 1) / S;  // chunks per slice
            S = (nchunks + per - 1) / per;
            if (S > 1) {
                a.split = (int)S;
                a.split_len = (int)(per * LEV_PACK_CHUNK);
                a.split_first = split_scratch;
                a.split_ticket = split_scratch + t->N;
                if (cudaMemsetAsync(split_scratch, 0, sizeof(int32_t) * (size_t)(t->N + nblk), st) != cudaSuccess)
                    return lev_check_cuda("memset");
                nblk *= S;
            }
        }
    }
    const dim3 grid((unsigned)nblk, 1, 1);
    if (warpseq) {  // a warp per 32 sequences: 4 blocks per CTA and iteration
        int64_t n = ((t->N + 31) / 32 + 3) / 4;
        const bool many = n > cap && (bv_check || n >= 8 * cap);
        if (many) {
            const int64_t per_cta = (n + cap - 1) / cap;
            n = (n + per_cta - 1) / per_cta;
        }
        if (!a.hist_bins) {  // (the branch above sized it for 32-sequence CTAs)
            const int64_t bins = (int64_t)LEV_GROUP_NCLS * (t->T + 1);
            if (ghist != nullptr && many && bins <= 8192) {
                a.hist_bins = (int)bins;
                smem = sizeof(int) * (size_t)bins;
            }
        }
        nblk = n;
    }
    const dim3 wgrid((unsigned)nblk, 1, 1);
#define LEV_PACK_CASE(TT)                                                \
    if (warpseq)                                                         \
        lev_launch(lev_pack_warpseq_kernel<TT>, wgrid, block, smem, st, a); \
    else if (seqfirst)                                                   \
        lev_launch(lev_pack_seqfirst_kernel<TT>, grid, block, smem, st, a); \
    else                                                                 \
        lev_launch(lev_pack_rows_kernel<TT>, grid, block, 0, st, a);
    switch (t->elem_bytes) {
        case 8: LEV_PACK_CASE(int64_t) break;
        case 4: LEV_PACK_CASE(int32_t) break;
        case 2: LEV_PACK_CASE(int16_t) break;
        case 1: LEV_PACK_CASE(int8_t) break;
        default:
            lev_set_error("unsupported token element size %d", (int)t->elem_bytes);
            return B200LEV_ERR_ARG;
    }
#undef LEV_PACK_CASE
    return lev_check_cuda("lev_pack_kernel");
}
