// lev_bitvec.cu -- unit-cost fast path: bit-parallel Levenshtein, one LANE per pair.
//
// When ins == del == sub (SM:168-174 reduces that case to unit costs and a multiplier) the
// DP row of SM:286-350 is determined by its +1/0/-1 differences, and Myers' bit-vector
// recurrence (J. ACM 46(3), 1999; Hyyro's edit-distance form) advances a whole row of up to
// 32*W reference positions with ~10 logic operations per 32-bit word instead of ~2.5
// instructions per CELL in the packed wavefront kernel (lev_group.cu).  The value the
// reference reads off each row -- column r, the distance between the full reference and a
// hypothesis prefix (SM:352-358, 386-405) -- is the running score of the recurrence, so
// the final and the prefix modes both fall out of one pass over the hypothesis.
//
// Two kernels, both with lane = pair and no transposition anywhere: a sequence-first token
// tensor (T, N) is already "position-major, pair-minor", so lane l of a warp reading
// tok[t][n0 + l] is one coalesced row, and the (H', N) output row of 32 neighbouring pairs is
// one 128-byte store.  K0 (lev_pack.cu) is not needed on this path.
//
//   lev_bv_uid_kernel   per lane: an open-addressing hash of the pair's reference tokens in
//       shared memory (slot-major, lane-minor: every lane stays in its own bank).  Emits, 16
//       positions per 128-bit store, ref_uid[j] = first position holding ref[j]'s token and
//       hyp_uid[i] = first reference position matching hyp[i] (0xff: none); finds the
//       lengths (SM:137-143, 195-228) and raises the warning flags on the way.  Reads the
//       raw tokens once: HBM-bound.
//   lev_bv_dp_kernel    per lane: the match masks Peq[uid] of its reference in shared memory
//       (same interleaving), then one Myers step per hypothesis token.  The reference is
//       RIGHT-aligned in the 32*W-bit vector: the o = 32W - r low bits start with Pv = 0 and
//       never match, which makes each of them a copy of the boundary row D[0][i] = i, so the
//       boundary enters the first real bit through the ordinary shift and the score is always
//       read at the top bit of the last word.
//
// Eligibility (lev_bitvec_launch): unit costs after SM:168-174, final or prefix mode, at most
// 128 reference positions, unit stride along the batch axis of both token tensors and of
// the prefix output.  Everything else keeps the wavefront kernels.
#include "lev_common.cuh"

#define LEV_BV_NOMATCH 0xffu

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
    uint4* ref_uid;    // [ceil(R/16)][P]  16 uid bytes per (chunk, pair)
    uint4* hyp_uid;    // [ceil(H/16)][P]
    int32_t* flags;    // caller's warning flags, may be NULL
    int slots_log2;    // hash slots per pair (power of two >= 2 R)
    int mode, norm, exclude_last, Hout;
    float mult, padding;
    float* out;
    int64_t out_si;  // prefix: elements between output rows (pairs are adjacent)
};

__device__ __forceinline__ unsigned lev_bv_hash(int v, int slots_log2) {
    return ((unsigned)v * 0x9E3779B1u) >> (32 - slots_log2);
}

// ---------------------------------------------------------------------------------------
// uid pre-pass
// ---------------------------------------------------------------------------------------
template <typename TT>
__global__ void __launch_bounds__(32) lev_bv_uid_kernel(const LevBvArgs a) {
    LEV_DYN_SMEM(int, smem);
    const int lane = threadIdx.x;
    const int nslots = 1 << a.slots_log2, smask = nslots - 1;
    int* keys = smem;                                                        // [nslots][32]
    unsigned char* pos = reinterpret_cast<unsigned char*>(smem + nslots * 32);  // [nslots][32]
    const int64_t pair = (int64_t)blockIdx.x * 32 + lane;
    const bool valid = pair < a.P;
    const int64_t pc = valid ? pair : (int64_t)a.P - 1;  // lanes past the batch shadow the last pair
    {
        uint4* pz = reinterpret_cast<uint4*>(pos);
        for (int i = lane; i < nslots * 2; i += 32) pz[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
    }
    __syncwarp();
    const TT* __restrict__ rsrc = reinterpret_cast<const TT*>(a.ref) + pc / a.ref_group;
    const TT* __restrict__ hsrc = reinterpret_cast<const TT*>(a.hyp) + pc;
    int myflags = 0, wide = 0;

    // ---- reference: insert, ref_uid[j] = first position of the token at j ---------------
    int rlen = a.R;
    bool open = true;  // no eos seen yet
    for (int c = 0; c * 16 < a.R; ++c) {
        TT buf[16];
#pragma unroll
        for (int k = 0; k < 16; ++k)
            buf[k] = (c * 16 + k < a.R) ? rsrc[(int64_t)(c * 16 + k) * a.ref_st] : (TT)0;
        unsigned q[4] = {~0u, ~0u, ~0u, ~0u};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int t = c * 16 + k;
            const int64_t x = (int64_t)buf[k];
            const int v = (int)x;
            bool take = open && t < a.R;
            if (take && a.has_eos && x == a.eos) {  // SM:137-143, 198-218
                open = false;
                rlen = t + (a.include_eos ? 1 : 0);
                take = a.include_eos != 0;
            }
            if (take) {
                if (sizeof(TT) == 8 && (int64_t)v != x) wide = 1;
                unsigned slot = lev_bv_hash(v, a.slots_log2), u;
                while (true) {
                    const unsigned p = pos[slot * 32 + lane];
                    const int key = keys[slot * 32 + lane];
                    if (p == LEV_BV_NOMATCH) {
                        pos[slot * 32 + lane] = (unsigned char)t;
                        keys[slot * 32 + lane] = v;
                        u = (unsigned)t;
                        break;
                    }
                    if (key == v) {
                        u = p;
                        break;
                    }
                    slot = (slot + 1) & smask;
                }
                q[k >> 2] = (q[k >> 2] & ~(0xffu << (8 * (k & 3)))) | (u << (8 * (k & 3)));
            }
        }
        if (valid) a.ref_uid[(int64_t)c * a.P + pair] = make_uint4(q[0], q[1], q[2], q[3]);
    }
    if (open && a.has_eos && a.include_eos) myflags |= B200LEV_FLAG_REF_NO_EOS;

    // ---- hypothesis: look up, hyp_uid[i] = first reference position with hyp[i]'s token ---
    int hlen = a.H;
    open = true;
    for (int c = 0; c * 16 < a.H; ++c) {
        TT buf[16];
#pragma unroll
        for (int k = 0; k < 16; ++k)
            buf[k] = (c * 16 + k < a.H) ? lev_ldg_stream(hsrc + (int64_t)(c * 16 + k) * a.hyp_st) : (TT)0;
        unsigned q[4] = {~0u, ~0u, ~0u, ~0u};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int t = c * 16 + k;
            const int64_t x = (int64_t)buf[k];
            const int v = (int)x;
            bool take = open && t < a.H;
            if (take && a.has_eos && x == a.eos) {
                open = false;
                hlen = t + (a.include_eos ? 1 : 0);
                take = a.include_eos != 0;
            }
            if (take) {
                if (sizeof(TT) == 8 && (int64_t)v != x) wide = 1;
                unsigned slot = lev_bv_hash(v, a.slots_log2), u;
                while (true) {
                    const unsigned p = pos[slot * 32 + lane];
                    const int key = keys[slot * 32 + lane];
                    if (p == LEV_BV_NOMATCH || key == v) {
                        u = p;
                        break;
                    }
                    slot = (slot + 1) & smask;
                }
                q[k >> 2] = (q[k >> 2] & ~(0xffu << (8 * (k & 3)))) | (u << (8 * (k & 3)));
            }
        }
        if (valid) a.hyp_uid[(int64_t)c * a.P + pair] = make_uint4(q[0], q[1], q[2], q[3]);
    }
    if (open && a.has_eos && a.include_eos) myflags |= B200LEV_FLAG_HYP_NO_EOS;

    if (sizeof(TT) == 8 && wide && valid) {
        // rare: some token of this pair does not fit in int32, so equal low words prove
        // nothing -- redo this lane's uids by exact comparison (O(r (r + h)) cached loads)
        for (int c = 0; c * 16 < a.R; ++c) {
            unsigned q[4] = {~0u, ~0u, ~0u, ~0u};
            for (int k = 0; k < 16; ++k) {
                const int j = c * 16 + k;
                if (j >= rlen) break;
                const TT x = rsrc[(int64_t)j * a.ref_st];
                int u = j;
                for (int e = 0; e < j; ++e)
                    if (rsrc[(int64_t)e * a.ref_st] == x) {
                        u = e;
                        break;
                    }
                q[k >> 2] = (q[k >> 2] & ~(0xffu << (8 * (k & 3)))) | ((unsigned)u << (8 * (k & 3)));
            }
            a.ref_uid[(int64_t)c * a.P + pair] = make_uint4(q[0], q[1], q[2], q[3]);
        }
        for (int c = 0; c * 16 < a.H; ++c) {
            unsigned q[4] = {~0u, ~0u, ~0u, ~0u};
            for (int k = 0; k < 16; ++k) {
                const int i = c * 16 + k;
                if (i >= hlen) break;
                const TT x = hsrc[(int64_t)i * a.hyp_st];
                unsigned u = LEV_BV_NOMATCH;
                for (int e = 0; e < rlen; ++e)
                    if (rsrc[(int64_t)e * a.ref_st] == x) {
                        u = (unsigned)e;
                        break;
                    }
                q[k >> 2] = (q[k >> 2] & ~(0xffu << (8 * (k & 3)))) | (u << (8 * (k & 3)));
            }
            a.hyp_uid[(int64_t)c * a.P + pair] = make_uint4(q[0], q[1], q[2], q[3]);
        }
    }

    if (valid) {
        a.hyp_len[pair] = hlen;
        if (pair % a.ref_group == 0) a.ref_len[pair / a.ref_group] = rlen;
        if (a.norm && rlen == 0) myflags |= B200LEV_FLAG_EMPTY_REF;  // SM:360-366, 397-404
    } else {
        myflags = 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) myflags |= __shfl_xor_sync(LEV_FULL_MASK, myflags, o);
    if (lane == 0 && myflags != 0 && a.flags != nullptr) atomicOr(a.flags, myflags);
}

// ---------------------------------------------------------------------------------------
// the DP
// ---------------------------------------------------------------------------------------
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

template <int W, int MODE>
__global__ void __launch_bounds__(32) lev_bv_dp_kernel(const LevBvArgs a) {
    LEV_DYN_SMEM(unsigned, M);  // Peq[(R + 1)][W][32 lanes]; row R stays zero (no match)
    const int lane = threadIdx.x;
    const int64_t pair = (int64_t)blockIdx.x * 32 + lane;
    const bool valid = pair < a.P;
    const int64_t pc = valid ? pair : (int64_t)a.P - 1;
    const int r = a.ref_len[pc / a.ref_group], h = a.hyp_len[pc];
    const int Z = a.R;
    {
        uint4* mz = reinterpret_cast<uint4*>(M);
        const int n16 = (a.R + 1) * W * 8;
        for (int i = lane; i < n16; i += 32) mz[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
    // the reference occupies bits [o, 32 W)
    const int o = 32 * W - r;
    int rmax = r;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const int other = __shfl_xor_sync(LEV_FULL_MASK, rmax, d);
        rmax = other > rmax ? other : rmax;
    }
    for (int c = 0; c * 16 < rmax; ++c) {
        const uint4 q4 = a.ref_uid[(int64_t)c * a.P + pc];
        const unsigned q[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int j = c * 16 + k;
            if (j < r) {
                const unsigned u = (q[k >> 2] >> (8 * (k & 3))) & 0xffu;
                const int b = j + o;
                M[(u * W + (b >> 5)) * 32 + lane] |= 1u << (b & 31);
            }
        }
    }
    unsigned pv[W], mv[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const int lo = 32 * w;
        pv[w] = o <= lo ? ~0u : (o >= lo + 32 ? 0u : (~0u << (o - lo)));
        mv[w] = 0u;
    }
    int score = r;
    // rows this warp has to run: the longest hypothesis among its pairs
    const int mysteps = a.exclude_last ? (h > 0 ? h - 1 : 0) : h;  // SM:286-288
    int steps = mysteps;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const int other = __shfl_xor_sync(LEV_FULL_MASK, steps, d);
        steps = other > steps ? other : steps;
    }
    const float rf = (float)r;
    const float y = r > 0 ? __frcp_rn(rf) : 0.0f;
    const int first_pad = h + (a.exclude_last ? 0 : 1);
    float* __restrict__ orow = a.out + pair;  // prefix: row 0 of this pair
    auto emit = [&](int i, int rv) {          // SM:352-388: one output row
        float val = __fmul_rn((float)rv, a.mult);
        if (a.norm) {
            if (r == 0) {
                val = i > 0 ? 1.0f : 0.0f;
            } else {  // correctly rounded val / r (Markstein, see lev_group.cu)
                const float q0 = __fmul_rn(val, y);
                val = __fmaf_rn(__fmaf_rn(-rf, q0, val), y, q0);
            }
        }
        if (valid) *orow = i < first_pad ? val : a.padding;
        orow += a.out_si;
    };
    int fin = r;  // FINAL: hypotheses without rows keep D[r][0] = r
    if (MODE == LEV_MODE_PREFIX && a.Hout > 0) emit(0, r);
    int i = 1;
    for (int c = 0; c * 16 < steps; ++c) {
        const uint4 q4 = a.hyp_uid[(int64_t)c * a.P + pc];
        const unsigned q[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
        for (int k = 0; k < 16; ++k, ++i) {
            if (i > steps) break;
            unsigned u = (q[k >> 2] >> (8 * (k & 3))) & 0xffu;
            u = u < (unsigned)Z ? u : (unsigned)Z;
            unsigned eq[W];
#pragma unroll
            for (int w = 0; w < W; ++w) eq[w] = M[(u * W + w) * 32 + lane];
            score += lev_bv_step<W>(eq, pv, mv);
            if (MODE == LEV_MODE_PREFIX) {
                emit(i, score);
            } else if (i == h) {
                fin = score;
            }
        }
    }
    if (MODE == LEV_MODE_PREFIX) {
        for (; i < a.Hout; ++i) {  // rows past every hypothesis of this warp
            if (valid) *orow = a.padding;
            orow += a.out_si;
        }
    } else if (valid) {  // SM:390-405
        float val = __fmul_rn((float)fin, a.mult);
        if (a.norm) val = (r == 0) ? (h > 0 ? 1.0f : 0.0f) : val / rf;
        a.out[pair] = val;
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// EXPERIMENTAL, off unless B200LEV_BITVEC=1.  Parity-green on the GPU and in the emulator, but
// on cfg2 (131072 pairs, R = H = 101, W = 4) the two kernels take 390 + 165 us against 314 us
// for the whole wavefront path: per-lane tables cost 40 KB (hash) and 52 KB (Peq) of shared
// memory per WARP, i.e. one warp per scheduler, and both kernels are serial dependency
// chains (ncu: 0.2 IPC, "wait"/scoreboard stalls).  What it needs to win -- tables shared
// by the lanes of an n-best group, or 2-4 lanes per pair, and a probe loop that does not
// run at the pace of the slowest of 32 lanes -- is listed in DESIGN.md.
static bool lev_bv_enabled() {
    const char* e = getenv("B200LEV_BITVEC");
    return e != nullptr && atoi(e) != 0;
}
static int64_t lev_bv_min_pairs() {
    int64_t v = 1024;  // below this the warp-per-pair kernels have the lower latency
    if (const char* e = getenv("B200LEV_BITVEC_MIN_PAIRS")) v = atoll(e);
    return v;
}

bool lev_bitvec_eligible(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp, int mode,
                         bool count_mode, bool float_path, int ins_i, int del_i, int sub_i,
                         int64_t out_sn) {
    if (!lev_bv_enabled()) return false;
    if (mode != LEV_MODE_FINAL && mode != LEV_MODE_PREFIX) return false;
    if (count_mode || float_path || ins_i != 1 || del_i != 1 || sub_i != 1) return false;
    if (ref->T > 128 || ref->T < 1 || hyp->T >= (1 << 20)) return false;
    if (hyp->N < lev_bv_min_pairs()) return false;
    if (ref->elem_bytes != hyp->elem_bytes) return false;
    if ((ref->N > 1 && ref->stride_n != 1) || (hyp->N > 1 && hyp->stride_n != 1)) return false;
    if (mode == LEV_MODE_PREFIX && out_sn != 1) return false;
    return true;
}

template <typename TT>
static void lev_bv_launch_uid(const LevBvArgs& a, size_t smem, cudaStream_t st) {
    auto kern = lev_bv_uid_kernel<TT>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lev_launch(kern, dim3((unsigned)((a.P + 31) / 32)), dim3(32), smem, st, a);
}

template <int W>
static void lev_bv_launch_dp(const LevBvArgs& a, cudaStream_t st) {
    const size_t smem = sizeof(unsigned) * (size_t)(a.R + 1) * W * 32;
    const dim3 grid((unsigned)((a.P + 31) / 32)), block(32);
    if (a.mode == LEV_MODE_PREFIX) {
        auto kern = lev_bv_dp_kernel<W, LEV_MODE_PREFIX>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lev_launch(kern, grid, block, smem, st, a);
    } else {
        auto kern = lev_bv_dp_kernel<W, LEV_MODE_FINAL>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lev_launch(kern, grid, block, smem, st, a);
    }
}

// The caller has checked lev_bitvec_eligible.  `uid_ref` / `uid_hyp` are workspace regions of
// P * round_up(R, 16) and P * round_up(H, 16) bytes.
int lev_bitvec_launch(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                      const b200lev_opts_t* o, int mode, float mult, int32_t* ref_len,
                      int32_t* hyp_len, void* uid_ref, void* uid_hyp, int32_t* flags, float* out,
                      int64_t out_si, int Hout, cudaStream_t st) {
    LevBvArgs a;
    memset(&a, 0, sizeof(a));
    a.ref = ref->data;
    a.hyp = hyp->data;
    a.ref_st = ref->stride_t;
    a.hyp_st = hyp->stride_t;
    a.R = (int)ref->T;
    a.H = (int)hyp->T;
    a.P = (int)hyp->N;
    a.ref_group = o->ref_group;
    a.has_eos = o->has_eos;
    a.eos = o->eos;
    a.include_eos = o->include_eos;
    a.ref_len = ref_len;
    a.hyp_len = hyp_len;
    a.ref_uid = (uint4*)uid_ref;
    a.hyp_uid = (uint4*)uid_hyp;
    a.flags = flags;
    a.slots_log2 = 5;
    while ((1 << a.slots_log2) < 2 * a.R) ++a.slots_log2;
    a.mode = mode;
    a.norm = o->norm;
    a.exclude_last = mode == LEV_MODE_PREFIX ? o->exclude_last : 0;
    a.Hout = Hout;
    a.mult = mult;
    a.padding = (float)o->padding;
    a.out = out;
    a.out_si = out_si;
    const size_t smem_uid = (size_t)(1 << a.slots_log2) * 32 * (sizeof(int) + 1);
    if (getenv("B200LEV_TRACE"))
        fprintf(stderr, "b200lev: bit-vector path, mode %d, R=%d H=%d P=%d W=%d\n", mode, a.R, a.H,
                a.P, (a.R + 31) / 32);
    lev_prof_begin(LEV_PROF_PACK_HYP, st);
    switch (ref->elem_bytes) {
        case 8: lev_bv_launch_uid<int64_t>(a, smem_uid, st); break;
        case 4: lev_bv_launch_uid<int32_t>(a, smem_uid, st); break;
        case 2: lev_bv_launch_uid<int16_t>(a, smem_uid, st); break;
        case 1: lev_bv_launch_uid<int8_t>(a, smem_uid, st); break;
        default:
            lev_set_error("unsupported token element size %d", (int)ref->elem_bytes);
            return B200LEV_ERR_ARG;
    }
    lev_prof_end(LEV_PROF_PACK_HYP, st);
    int rc = lev_check_cuda("lev_bv_uid_kernel");
    if (rc) return rc;
    lev_prof_begin(LEV_PROF_DP, st);
    const int W = (a.R + 31) / 32;
    switch (W) {
        case 1: lev_bv_launch_dp<1>(a, st); break;
        case 2: lev_bv_launch_dp<2>(a, st); break;
        case 3: lev_bv_launch_dp<3>(a, st); break;
        default: lev_bv_launch_dp<4>(a, st); break;
    }
    lev_prof_end(LEV_PROF_DP, st);
    return lev_check_cuda("lev_bv_dp_kernel");
}
