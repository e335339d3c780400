// lev_bitvec.cuh -- pieces shared by the unit-cost bit-vector kernels (lev_bitvec.cu: the
// two-kernel uid + DP form; lev_bvfused.cu: the fused single-kernel form).
#pragma once
#include "lev_common.cuh"

#define LEV_BV_NOMATCH 0xffu
#define LEV_BV_EMPTY ((int)0x80000000)  // key of a free way (a token with this value: exact path)
constexpr int LEV_BV_NT = 4;     // tables per warp = runs of identical references per pass
constexpr int LEV_BV_WARPS = 4;  // warps per CTA (independent: no block barrier anywhere)

struct LevBvArgs {
    const void* ref;  // raw token tensors, stride 1 along the batch axis
    const void* hyp;
    int64_t ref_st, hyp_st;  // elements between positions
    int R, H, P, ref_group;
    int has_eos;
    int64_t eos;
    int include_eos;
    int32_t* ref_len;  // [Nref]  (same workspace slots K0 fills on the other paths)
    int32_t* hyp_len;  // [P]
    uint4* ref_uid;    // [ceil(R/16)][P]  16 uid bytes per (chunk, pair); run leaders only
    uint4* hyp_uid;    // [ceil(H/16)][P]
    unsigned char* lead;  // [P] lane (0..31) of the leader of the pair's run in its 32-block
    int32_t* flags;    // caller's warning flags, may be NULL
    int32_t* state;  // workspace state words; [3] != 0 = "not for this path" (lev_bv_took)
    int check_state;   // device-selected mode: veto through state[3] instead of multi-pass
    int slots_log2;    // hash buckets per table (power of two >= R), 4 ways each
    int mode, norm, exclude_last, Hout;
    float mult, padding;
    float* out;
    int64_t out_si;  // prefix: elements between output rows (pairs are adjacent)
    double* acc;     // FINAL, short-reference kernel: += [sum(out), sum(ref_len per pair), #pairs] (may be NULL)
};

// Bucketed hash: 4 ways per bucket, one 128-bit read of the keys and one 32-bit read of the
// position bytes settle a probe -- no loop, so 32 lanes with 32 different tokens cost the same
// as one.  A run leader builds its table with the first of a few multipliers under which no
// bucket overflows (load <= 1 token per bucket on average: a couple of tries at most).
// (a build with R = 128 distinct tokens in 128 buckets fails with p = 0.38: 16 tries leave
// 2e-7 of the runs to the slow exact path; the typical R ~ 100 needs 1.2 tries on average)
constexpr int LEV_BV_TRIES = 16;
__device__ __forceinline__ unsigned lev_bv_mult(int seed) {
    return (0x9E3779B1u * (2u * (unsigned)seed + 1u)) ^ ((unsigned)seed * 0x85EBCA6Au);
}
__device__ __forceinline__ unsigned lev_bv_hash(int v, unsigned mult, int buckets_log2) {
    return ((unsigned)v * mult) >> (32 - buckets_log2);
}

// Runs of identical references inside a block of 32 consecutive pairs, from each lane's
// "same as the lane before" bit: leader lane of the run, its index, number of runs.
struct LevBvRuns {
    int lead, index, count, len;
    unsigned mask;  // the lanes of my run
};
__device__ __forceinline__ LevBvRuns lev_bv_runs(bool same, int lane) {
    const unsigned starts = ~__ballot_sync(LEV_FULL_MASK, same && lane > 0);
    const unsigned upto = starts & (0xffffffffu >> (31 - lane));
    LevBvRuns r;
    r.lead = 31 - __clz((int)upto);
    r.index = __popc(upto) - 1;
    r.count = __popc(starts);
    const unsigned higher = r.lead == 31 ? 0u : (starts & ~(0xffffffffu >> (31 - r.lead)));
    const int next = higher ? __ffs((int)higher) - 1 : 32;
    r.len = next - r.lead;
    r.mask = (r.len == 32 ? 0xffffffffu : ((1u << r.len) - 1u)) << r.lead;
    return r;
}

__device__ __forceinline__ unsigned lev_bv_set_byte(unsigned word, unsigned byte, int k) {
    return (word & ~(0xffu << (8 * k))) | (byte << (8 * k));
}

// sum = t + pv over W 32-bit words (carry chain in one asm block)
template <int W>
__device__ __forceinline__ void lev_bv_add(const unsigned (&t)[W], const unsigned (&pv)[W],
                                           unsigned (&sum)[W]) {
#ifdef B200LEV_EMU
    unsigned long long carry = 0;
    for (int w = 0; w < W; ++w) {
        const unsigned long long s = (unsigned long long)t[w] + pv[w] + carry;
        sum[w] = (unsigned)s;
        carry = s >> 32;
    }
#else
    if (W == 1) {
        sum[0] = t[0] + pv[0];
    } else if (W == 2) {
        asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;"
            : "=r"(sum[0]), "=r"(sum[W > 1 ? 1 : 0])
            : "r"(t[0]), "r"(t[W > 1 ? 1 : 0]), "r"(pv[0]), "r"(pv[W > 1 ? 1 : 0]));
    } else if (W == 3) {
        asm("add.cc.u32 %0, %3, %6;\n\taddc.cc.u32 %1, %4, %7;\n\taddc.u32 %2, %5, %8;"
            : "=r"(sum[0]), "=r"(sum[W > 1 ? 1 : 0]), "=r"(sum[W > 2 ? 2 : 0])
            : "r"(t[0]), "r"(t[W > 1 ? 1 : 0]), "r"(t[W > 2 ? 2 : 0]), "r"(pv[0]),
              "r"(pv[W > 1 ? 1 : 0]), "r"(pv[W > 2 ? 2 : 0]));
    } else {
        asm("add.cc.u32 %0, %4, %8;\n\taddc.cc.u32 %1, %5, %9;\n\taddc.cc.u32 %2, %6, %10;\n\t"
            "addc.u32 %3, %7, %11;"
            : "=r"(sum[0]), "=r"(sum[W > 1 ? 1 : 0]), "=r"(sum[W > 2 ? 2 : 0]),
              "=r"(sum[W > 3 ? 3 : 0])
            : "r"(t[0]), "r"(t[W > 1 ? 1 : 0]), "r"(t[W > 2 ? 2 : 0]), "r"(t[W > 3 ? 3 : 0]),
              "r"(pv[0]), "r"(pv[W > 1 ? 1 : 0]), "r"(pv[W > 2 ? 2 : 0]), "r"(pv[W > 3 ? 3 : 0]));
    }
#endif
}

// One row of the recurrence; returns the score change at the top bit (reference column r).
template <int W>
__device__ __forceinline__ int lev_bv_step(const unsigned (&eq)[W], unsigned (&pv)[W],
                                           unsigned (&mv)[W]) {
    unsigned t[W], sum[W], ph[W], mh[W];
#pragma unroll
    for (int w = 0; w < W; ++w) t[w] = eq[w] & pv[w];
    lev_bv_add<W>(t, pv, sum);
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const unsigned xh = (sum[w] ^ pv[w]) | eq[w];
        ph[w] = mv[w] | ~(xh | pv[w]);
        mh[w] = pv[w] & xh;
    }
    const int delta = (int)(ph[W - 1] >> 31) - (int)(mh[W - 1] >> 31);
#pragma unroll
    for (int w = W - 1; w >= 0; --w) {
        const unsigned phs = w ? __funnelshift_l(ph[w - 1], ph[w], 1) : ((ph[0] << 1) | 1u);
        const unsigned mhs = w ? __funnelshift_l(mh[w - 1], mh[w], 1) : (mh[0] << 1);
        const unsigned xv = eq[w] | mv[w];
        pv[w] = mhs | ~(xv | phs);
        mv[w] = phs & xv;
    }
    return delta;
}


// ---------------------------------------------------------------------------------------
// pieces of the fused kernel (lev_bvfused.cu)
// ---------------------------------------------------------------------------------------
bool lev_bvshort_supports(int elem_bytes, int64_t R);
int lev_bvshort_launch(const LevBvArgs& a, int elem_bytes, cudaStream_t st);
bool lev_bvfused_supports(int elem_bytes);
int lev_bvfused_launch(const LevBvArgs& a, int elem_bytes, cudaStream_t st, void* after_probe);

struct LevBvRefScan {
    int first_eos;  // position of the first eos in my reference column (R: none)
    int exact0;     // the column holds a token the 32-bit keys cannot represent
    bool diff;      // my column differs from my left neighbour's (a run starts at this lane)
    bool vetoed;    // (may_veto) more than LEV_BV_NT runs after the first tokens
};

// One coalesced pass over a lane's reference column, 16 positions per half, the two halves'
// loads ping-pong: runs of identical references (SM:1426/1439 repeat every reference over its
// n-best list), first eos (SM:137-143, 198-218), tokens outside int32.
template <typename TT>
__device__ __forceinline__ LevBvRefScan lev_bv_ref_scan(const LevBvArgs& a, const TT* __restrict__ rsrc,
                                                        const int rst, const int64_t rcol,
                                                        const int lane, const bool may_veto) {
    LevBvRefScan s;
    s.first_eos = a.R;
    s.exact0 = 0;
    s.vetoed = false;
    const int prev_col = __shfl_up_sync(LEV_FULL_MASK, (int)rcol, 1);
    const bool other = prev_col != (int)rcol;
    const int eos_lo = (int)a.eos, eos_hi = (int)(a.eos >> 32);
    int dacc = 0, wacc = 0;  // differences to the left neighbour / to a sign extension
    constexpr int DH = 16;
    auto observe = [&](const TT (&buf)[DH], int t0) {
        unsigned eos_bits = 0, empty_bits = 0;
#pragma unroll
        for (int k = 0; k < DH; ++k) {
            const int64_t x = (int64_t)buf[k];
            const int lo = (int)x, hi = (int)(x >> 32);
            dacc |= lo ^ __shfl_up_sync(LEV_FULL_MASK, lo, 1);
            if (sizeof(TT) == 8) {
                dacc |= hi ^ __shfl_up_sync(LEV_FULL_MASK, hi, 1);
                wacc |= hi ^ (lo >> 31);
            }
            eos_bits |= (lo == eos_lo && hi == eos_hi) ? (1u << k) : 0u;
            empty_bits |= lo == LEV_BV_EMPTY ? 1u : 0u;
        }
        // positions past R were loaded as 0 (= what the neighbour loaded): mask them out
        const int live = a.R - t0;
        if (live < DH) eos_bits &= live > 0 ? (1u << live) - 1u : 0u;
        if (a.has_eos && eos_bits != 0 && s.first_eos == a.R) s.first_eos = t0 + __ffs((int)eos_bits) - 1;
        if (empty_bits) s.exact0 = 1;
    };
    const TT* rp = rsrc;
    TT bufA[DH], bufB[DH];
#pragma unroll
    for (int k = 0; k < DH; ++k) bufA[k] = (k < a.R) ? rp[(int64_t)k * rst] : (TT)0;
#pragma unroll 1
    for (int t0 = 0; t0 < a.R; t0 += 2 * DH) {
        LEV_OPAQUE_PTR(rp);
#pragma unroll
        for (int k = 0; k < DH; ++k) bufB[k] = (t0 + DH + k < a.R) ? rp[(int64_t)(DH + k) * rst] : (TT)0;
        observe(bufA, t0);
        if (may_veto && t0 == 0) {
            const unsigned same = __ballot_sync(LEV_FULL_MASK, lane > 0 && !(dacc != 0 && other));
            if (32 - __popc(same) > LEV_BV_NT) {
                s.vetoed = true;
                break;
            }
        }
        rp += 2 * DH * (int64_t)rst;
        LEV_OPAQUE_PTR(rp);
#pragma unroll
        for (int k = 0; k < DH; ++k) bufA[k] = (t0 + 2 * DH + k < a.R) ? rp[(int64_t)k * rst] : (TT)0;
        observe(bufB, t0 + DH);
    }
    s.diff = dacc != 0 && other;
    if (wacc != 0) s.exact0 = 1;
    return s;
}

// The lanes of a run build its hash table together (warp-collective: every lane calls it).
// Lane i of a run of n lanes takes positions i, i + n, ...: it claims a free way of the token's
// bucket with a CAS on the key word (an equal key already there = duplicate), then writes its
// position byte.  Which occurrence ends up representing a token is a race, and does not
// matter: a position byte only has to be the SAME for every occurrence and for the hypothesis
// look-ups, which all read the finished table.  Positions past the eos may get in too: their
// rows receive no bits, so matching them is no match.  A build whose bucket overflows is
// retried under the next multiplier.  Returns 1 if the run has no usable table (tokens outside
// int32, or LEV_BV_TRIES overflowing builds): the caller compares tokens one by one.
template <typename TT>
__device__ __forceinline__ int lev_bv_build_table(const LevBvArgs& a, int4* keys4, unsigned* posw,
                                                  const int nb, const TT* __restrict__ rsrc,
                                                  const int rst, const LevBvRuns& runs,
                                                  const bool active, const int tb, const int run_pos,
                                                  const int exact0, const int lane, int* seed_out) {
    constexpr int NT = LEV_BV_NT;
    int seed = 0;
    int exact = exact0;
    bool need = active && !exact;
#pragma unroll 1
    for (int attempt = 0; attempt < LEV_BV_TRIES; ++attempt) {
        unsigned clear = 0;  // tables (of this pass) being (re)built
#pragma unroll
        for (int t = 0; t < NT; ++t)
            if (__any_sync(LEV_FULL_MASK, need && tb == t)) clear |= 1u << t;
        if (clear == 0) break;
        __syncwarp();
        for (int i = lane; i < nb * NT; i += 32)
            if ((clear >> (i % NT)) & 1u) {
                keys4[i] = make_int4(LEV_BV_EMPTY, LEV_BV_EMPTY, LEV_BV_EMPTY, LEV_BV_EMPTY);
                posw[i] = (unsigned)a.R * 0x01010101u;  // free ways point at the all-zero row
            }
        __syncwarp();
        bool overflow = false, hopeless = false;
        if (need) {
            seed = attempt;
            const unsigned hmul = lev_bv_mult(seed);
            for (int t0 = run_pos; t0 < a.R; t0 += 8 * runs.len) {
                int vv[8];  // 8 of my positions at a time: the loads overlap
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    vv[k] = (t0 + k * runs.len < a.R) ? (int)rsrc[(int64_t)(t0 + k * runs.len) * rst] : 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int t = t0 + k * runs.len;
                    if (t >= a.R) break;
                    const int v = vv[k];
                    if (v == LEV_BV_EMPTY) {  // the one token value a key word cannot hold
                        hopeless = true;
                        continue;
                    }
                    const int idx = (int)lev_bv_hash(v, hmul, a.slots_log2) * NT + tb;
                    int* kk = reinterpret_cast<int*>(&keys4[idx]);
                    int w = 0;
                    for (; w < 4; ++w) {
                        const int old = atomicCAS(&kk[w], LEV_BV_EMPTY, v);
                        if (old == LEV_BV_EMPTY) {
                            reinterpret_cast<unsigned char*>(&posw[idx])[w] = (unsigned char)t;
                            break;
                        }
                        if (old == v) break;
                    }
                    overflow |= w == 4;
                }
            }
        }
        // the run retries (next multiplier) if any of its lanes met a full bucket
        const bool run_bad = (__ballot_sync(LEV_FULL_MASK, overflow) & runs.mask) != 0;
        const bool run_hopeless = (__ballot_sync(LEV_FULL_MASK, hopeless) & runs.mask) != 0;
        need = need && run_bad;
        if (active && run_hopeless) {
            need = false;
            exact = 1;
        }
        if (need && attempt == LEV_BV_TRIES - 1) {
            need = false;
            exact = 1;
        }
    }
    __syncwarp();
    *seed_out = seed;
    return exact;
}
