// lev_group.cu -- K1s: the throughput kernel for large ragged batches of short/mid pairs.
//
// Same anti-diagonal wavefront as lev_dp.cu, re-shaped so that warps do not idle on
// padding (north star item 3):
//   * a pair is handled by a GROUP of G lanes (G = 1..32, chosen from the padded
//     reference length), each lane owning C adjacent DP columns in registers, so the
//     ramp of the skewed wavefront costs G-1 steps instead of 31 and one __shfl_up_sync
//     serves C cells;
//   * the whole row fits one strip (G*C >= r+1, right-aligned, virtual +BIG columns on
//     the left) -- no shared-memory boundary column;
//   * a CTA takes a TILE of consecutive pairs and buckets them in shared memory by
//     (column class C, hypothesis length) with a counting sort; a warp then runs 32/G
//     pairs of the SAME class and near-equal length in lock step, largest first, pulled
//     from a shared-memory work queue.  Lane utilisation is (r+1)/(G*C) instead of
//     (r+1)/(32*C), and no step is spent on a pair that has already finished.
//   * prefix values are parked in shared memory by the one lane that owns column r and
//     written out after the pair by all lanes (scale, IEEE division by the reference
//     length, padding fill: SM:356-386), so the epilogue arithmetic is off the DP loop.
//
// Integer costs only (cost row, or (cost, count) rows for the error-rate family);
// FINAL and PREFIX modes.  Everything else stays on lev_dp.cu.
#include <cstdlib>

#include "lev_common.cuh"

#define LEVG_NCLS 6
__device__ __forceinline__ int levg_class_cols(int cls) { return 8 + 4 * cls; }  // 8..28

struct LevGroupGeom {
    int G;      // lanes per pair
    int tile;   // pairs per CTA
    int Hs;     // shared-memory row stride (ints) of the per-pair token / prefix rows
    int nbins;  // LEVG_NCLS * (H + 1)
    int allow16;  // costs and lengths admit the packed 2 x int16 DPX path
};

// "infinity" of the packed path: BIG16 + (largest reachable value) must stay < 2^15
#define LEVG_BIG16 16000

// smallest class whose strip covers columns 0..r
__device__ __forceinline__ int levg_class_of(int r, int G) {
    const int need = (r + G) / G;  // ceil((r + 1) / G)
    int cls = (need - 8 + 3) >> 2;
    return cls < 0 ? 0 : cls;
}

template <bool COUNT, int MODE, int C>
__device__ __forceinline__ void levg_run(const LevParams& p, const int G, const int pair,
                                         const int r, const int h, const int steps,
                                         const int maxsteps, const int* __restrict__ hyp_row,
                                         int* __restrict__ pref_row) {
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const int refcol = pair >= 0 ? pair / p.ref_group : 0;
    const int32_t* __restrict__ rtok = p.ref_tok + (int64_t)refcol * p.Rp;
    const int insc = p.ins_i, delc = p.del_i, subc = p.sub_i;
    const int j0 = r - G * C + gl * C + 1;  // right-aligned: the group's last column is r
    int v[C], m[C], rt[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int j = j0 + c;
        v[c] = (j >= 0) ? j * delc : LEV_BIG_I32;  // SM:258-263
        m[c] = j >= 0 ? j : 0;                      // SM:260
        rt[c] = (j >= 1 && pair >= 0) ? rtok[j - 1] : 0;
    }
    (void)m;
    int pl_v = LEV_BIG_I32, pl_m = 0;
    (void)pl_m;
    const bool owner = (gl == G - 1);
    const int nsteps = maxsteps + G - 1;
    for (int s = 1; s <= nsteps; ++s) {
        const int sh_v = __shfl_up_sync(LEV_FULL_MASK, v[C - 1], 1, G);
        const int in_v = (gl == 0) ? LEV_BIG_I32 : sh_v;
        const int diag_v = pl_v;
        pl_v = in_v;
        int in_m = 0, diag_m = 0;
        if (COUNT) {
            const int sh_m = __shfl_up_sync(LEV_FULL_MASK, m[C - 1], 1, G);
            in_m = (gl == 0) ? 0 : sh_m;
            diag_m = pl_m;
            pl_m = in_m;
        }
        (void)in_m; (void)diag_m;
        const int i = s - gl;
        if ((unsigned)(i - 1) < (unsigned)steps) {
            const int ht = hyp_row[i - 1];
            if (COUNT) {  // SM:292-314
                int dc = diag_v, dm = diag_m, lc = in_v, lm = in_m;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int uc = v[c], um = m[c];
                    const bool neq = rt[c] != ht;
                    const int sub_c = dc + (neq ? subc : 0);
                    const int ins_c = uc + insc;
                    const bool ps = ins_c >= sub_c;
                    int cc = ps ? sub_c : ins_c;
                    int mm = ps ? dm + (neq ? 1 : 0) : um + 1;
                    const int del_c = lc + delc;
                    const bool keep = del_c >= cc;
                    cc = keep ? cc : del_c;
                    mm = keep ? mm : lm + 1;
                    dc = uc;
                    dm = um;
                    lc = cc;
                    lm = mm;
                    v[c] = cc;
                    m[c] = mm;
                }
            } else {
                int dg = diag_v, lf = in_v;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int up = v[c];
                    int sb = dg;
                    if (rt[c] != ht) sb += subc;
                    const int t = __viaddmin_s32(up, insc, sb);
                    lf = __viaddmin_s32(lf, delc, t);
                    dg = up;
                    v[c] = lf;
                }
            }
            if (MODE == LEV_MODE_PREFIX && owner) pref_row[i] = COUNT ? m[C - 1] : v[C - 1];
        }
    }
    if (MODE == LEV_MODE_FINAL && owner && pair >= 0) {  // SM:390-405
        float val = (float)(COUNT ? m[C - 1] : v[C - 1]) * p.mult;
        if (p.norm) val = (r == 0) ? (h > 0 ? 1.0f : 0.0f) : val / (float)r;
        p.out[pair] = val;
    }
}

// Packed path: TWO pairs per lane group, one in each 16-bit half of every register, driven
// by the 2-wide DPX instructions (VIADDMNMX.S16x2, VIMNMX.U16x2).  Per 2 cells: LOP3 (token
// xor), VIMNMX.U16x2 (-> 0/1 per half), IMAD (diag + neq*sub, FMA pipe), 2 x VIADDMNMX.S16x2.
// Needs 16-bit-injective tokens (K0 measured max - min < 65536), non-negative integer costs
// and values below LEVG_BIG16.  Cost row only.  Both pairs run `maxsteps` rows; rows past a
// pair's own length are never read back.
template <int C>
__device__ __forceinline__ void levg_run16(const LevParams& p, const int G, const int pairA,
                                           const int rA, const int pairB, const int rB,
                                           const int maxsteps,
                                           const unsigned* __restrict__ hyp_row,
                                           unsigned* __restrict__ pref_row) {
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const int32_t* __restrict__ rtA = p.ref_tok + (int64_t)(pairA >= 0 ? pairA / p.ref_group : 0) * p.Rp;
    const int32_t* __restrict__ rtB = p.ref_tok + (int64_t)(pairB >= 0 ? pairB / p.ref_group : 0) * p.Rp;
    const unsigned ins2 = (unsigned)p.ins_i * 0x00010001u, del2 = (unsigned)p.del_i * 0x00010001u;
    const unsigned subc = (unsigned)p.sub_i;
    const unsigned BIG2 = (unsigned)LEVG_BIG16 * 0x00010001u;
    const int j0A = rA - G * C + gl * C + 1, j0B = rB - G * C + gl * C + 1;
    unsigned v[C], rt[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int jA = j0A + c, jB = j0B + c;
        const unsigned vA = (jA >= 0) ? (unsigned)(jA * p.del_i) : (unsigned)LEVG_BIG16;
        const unsigned vB = (jB >= 0) ? (unsigned)(jB * p.del_i) : (unsigned)LEVG_BIG16;
        v[c] = vA | (vB << 16);
        const unsigned tA = (jA >= 1 && pairA >= 0) ? ((unsigned)rtA[jA - 1] & 0xffffu) : 0u;
        const unsigned tB = (jB >= 1 && pairB >= 0) ? ((unsigned)rtB[jB - 1] & 0xffffu) : 0u;
        rt[c] = tA | (tB << 16);
    }
    unsigned pl = BIG2;
    const bool owner = (gl == G - 1);
    const int nsteps = maxsteps + G - 1;
    for (int s = 1; s <= nsteps; ++s) {
        const unsigned sh = __shfl_up_sync(LEV_FULL_MASK, v[C - 1], 1, G);
        const unsigned in = (gl == 0) ? BIG2 : sh;
        unsigned dg = pl;
        pl = in;
        const int i = s - gl;
        if ((unsigned)(i - 1) < (unsigned)maxsteps) {
            const unsigned ht = hyp_row[i - 1];
            unsigned lf = in;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const unsigned up = v[c];
                const unsigned n01 = __vminu2(rt[c] ^ ht, 0x00010001u);
                const unsigned sb = n01 * subc + dg;
                const unsigned t = __viaddmin_s16x2(up, ins2, sb);
                lf = __viaddmin_s16x2(lf, del2, t);
                dg = up;
                v[c] = lf;
            }
            if (owner) pref_row[i] = v[C - 1];
        }
    }
}

template <bool COUNT, int MODE>
__global__ void __launch_bounds__(256) lev_group_kernel(const LevParams p, const LevGroupGeom geo) {
    LEV_DYN_SMEM(int, smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int G = geo.G, PPW = 32 / G, TILE = geo.tile, Hs = geo.Hs, H1 = p.H + 1;
    const int tile0 = blockIdx.x * TILE;
    const int ntile = min(TILE, p.P - tile0);
    // shared-memory carve-up
    int* hist = smem;                         // [nbins + 1]
    int* order = hist + geo.nbins + 1;        // [TILE + 2 * LEVG_NCLS * PPW]
    short* rl_s = reinterpret_cast<short*>(order + TILE + 2 * LEVG_NCLS * PPW);  // [TILE]
    short* hl_s = rl_s + TILE;                                               // [TILE]
    int* wbase = reinterpret_cast<int*>(hl_s + TILE);  // 2*TILE shorts: int-aligned
    int* hyp_w = wbase + (size_t)warp * 2 * PPW * Hs;  // [PPW][Hs]
    int* pref_w = hyp_w + PPW * Hs;                     // [PPW][Hs]
    __shared__ int next_task, ntasks, seg_end[LEVG_NCLS];

    // tokens that do not fit in int32: leave the whole batch to the 64-bit compare path of
    // lev_warp_kernel, which the host enqueues right behind this kernel
    if (*p.wide_flag & B200LEV_FLAG_WIDE_TOKENS) return;
    // K0 left the (biased) token range next to the flag word: [1] = max(u), [2] = max(~u),
    // u = token + 2^31.  All tokens inside a 65536-wide window <=> their low 16 bits are
    // injective <=> the packed path may compare 16-bit halves.
    const unsigned umax = (unsigned)p.wide_flag[1], umin = ~(unsigned)p.wide_flag[2];
    const bool packed = !COUNT && geo.allow16 && (umax < umin || umax - umin < 65536u);
    const int PPT = packed ? 2 * PPW : PPW;  // pairs per task
    // ---- 1. lengths, classes, histogram over (class desc, hyp length desc) ----
    for (int b = tid; b <= geo.nbins; b += blockDim.x) hist[b] = 0;
    for (int q = tid; q < TILE + 2 * LEVG_NCLS * PPW; q += blockDim.x) order[q] = -1;
    if (tid == 0) next_task = 0;
    __syncthreads();
    for (int q = tid; q < ntile; q += blockDim.x) {
        const int pair = tile0 + q;
        const int r = p.ref_len[pair / p.ref_group];
        const int h = p.hyp_len[pair];
        rl_s[q] = (short)r;
        hl_s[q] = (short)h;
        if (MODE != LEV_MODE_MASK && r == 0 && p.norm && p.flags != nullptr)
            atomicOr(p.flags, B200LEV_FLAG_EMPTY_REF);  // SM:360-366, 397-404
        const int cls = levg_class_of(r, G);
        atomicAdd(&hist[(LEVG_NCLS - 1 - cls) * H1 + (p.H - h)], 1);
    }
    __syncthreads();
    // ---- 2. exclusive scan (one warp; bins are few), class segments padded to PPW ----
    if (warp == 0) {
        int carry = 0;
        for (int cseg = 0; cseg < LEVG_NCLS; ++cseg) {
            for (int b0 = 0; b0 < H1; b0 += 32) {
                const int b = b0 + lane;
                const int cnt = b < H1 ? hist[cseg * H1 + b] : 0;
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(LEV_FULL_MASK, incl, o);
                    if (lane >= o) incl += t;
                }
                if (b < H1) hist[cseg * H1 + b] = carry + incl - cnt;
                carry += __shfl_sync(LEV_FULL_MASK, incl, 31);
            }
            carry = (carry + PPT - 1) / PPT * PPT;
            if (lane == 0) seg_end[cseg] = carry;
        }
        if (lane == 0) ntasks = carry / PPT;
    }
    __syncthreads();
    // ---- 3. scatter pair indices into sorted order ----
    for (int q = tid; q < ntile; q += blockDim.x) {
        const int cls = levg_class_of(rl_s[q], G);
        const int pos = atomicAdd(&hist[(LEVG_NCLS - 1 - cls) * H1 + (p.H - hl_s[q])], 1);
        order[pos] = q;
    }
    __syncthreads();

    // ---- 4. work queue: a task = PPW same-class pairs, one per lane group ----
    const int g = lane / G;
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&next_task, 1);
        t = __shfl_sync(LEV_FULL_MASK, t, 0);
        if (t >= ntasks) break;
        if (packed) {
            // ---- packed task: 2*PPW pairs, group g runs pairs 2g (low half) and 2g+1 ----
            const int qA = order[t * PPT + 2 * g], qB = order[t * PPT + 2 * g + 1];
            const int rA = qA >= 0 ? rl_s[qA] : 0, rB = qB >= 0 ? rl_s[qB] : 0;
            const int hA = qA >= 0 ? hl_s[qA] : 0, hB = qB >= 0 ? hl_s[qB] : 0;
            const int hmax = hA > hB ? hA : hB;
            int maxsteps = p.exclude_last ? (hmax > 0 ? hmax - 1 : 0) : hmax;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int other = __shfl_xor_sync(LEV_FULL_MASK, maxsteps, o);
                maxsteps = other > maxsteps ? other : maxsteps;
            }
            int cseg = 0;
            while (t * PPT >= seg_end[cseg]) ++cseg;
            const int cls = LEVG_NCLS - 1 - cseg;
            unsigned short* hyp16 = reinterpret_cast<unsigned short*>(hyp_w);
            for (int k = 0; k < PPT; ++k) {
                const int qk = order[t * PPT + k];
                if (qk < 0) continue;
                const int hk = hl_s[qk];
                const int sk = p.exclude_last ? (hk > 0 ? hk - 1 : 0) : hk;
                const int32_t* __restrict__ src = p.hyp_tok + (int64_t)(tile0 + qk) * p.Hp;
                for (int i = lane; i < sk; i += 32)
                    hyp16[(((k >> 1) * Hs + i) << 1) + (k & 1)] = (unsigned short)src[i];
            }
            __syncwarp();
            const unsigned* hrow = reinterpret_cast<const unsigned*>(hyp_w) + g * Hs;
            unsigned* prow = reinterpret_cast<unsigned*>(pref_w) + g * Hs;
            const int pA = qA >= 0 ? tile0 + qA : -1, pB = qB >= 0 ? tile0 + qB : -1;
            switch (cls) {
                case 0: levg_run16<8>(p, G, pA, rA, pB, rB, maxsteps, hrow, prow); break;
                case 1: levg_run16<12>(p, G, pA, rA, pB, rB, maxsteps, hrow, prow); break;
                case 2: levg_run16<16>(p, G, pA, rA, pB, rB, maxsteps, hrow, prow); break;
                case 3: levg_run16<20>(p, G, pA, rA, pB, rB, maxsteps, hrow, prow); break;
                case 4: levg_run16<24>(p, G, pA, rA, pB, rB, maxsteps, hrow, prow); break;
                default: levg_run16<28>(p, G, pA, rA, pB, rB, maxsteps, hrow, prow); break;
            }
            __syncwarp();
            // epilogue for all 2*PPW pairs (SM:279-285, 340-346, 356-386 / 390-405)
            for (int k = 0; k < PPT; ++k) {
                const int qk = order[t * PPT + k];
                if (qk < 0) continue;
                const int rk = rl_s[qk], hk = hl_s[qk];
                const unsigned* row = reinterpret_cast<const unsigned*>(pref_w) + (k >> 1) * Hs;
                const int sh16 = (k & 1) * 16;
                const float rf = (float)rk;
                if (MODE == LEV_MODE_PREFIX) {
                    const int first_pad = hk + (p.exclude_last ? 0 : 1);
                    const int64_t col = (int64_t)(tile0 + qk) * p.out_sn;
                    for (int i = lane; i < p.Hout; i += 32) {
                        float val;
                        if (i >= first_pad) {
                            val = p.padding;
                        } else {
                            const int raw = (i == 0) ? rk * p.del_i : (int)((row[i] >> sh16) & 0xffffu);
                            val = (float)raw * p.mult;
                            if (p.norm) val = (rk == 0) ? (i > 0 ? 1.0f : 0.0f) : val / rf;
                        }
                        p.out[(int64_t)i * p.out_si + col] = val;
                    }
                } else if (lane == 0) {
                    const int raw = (hk == 0) ? rk * p.del_i : (int)((row[hk] >> sh16) & 0xffffu);
                    float val = (float)raw * p.mult;
                    if (p.norm) val = (rk == 0) ? (hk > 0 ? 1.0f : 0.0f) : val / rf;
                    p.out[tile0 + qk] = val;
                }
            }
            __syncwarp();
            continue;
        }
        const int q = order[t * PPW + g];
        const int pair = q >= 0 ? tile0 + q : -1;
        const int r = q >= 0 ? rl_s[q] : 0;
        const int h = q >= 0 ? hl_s[q] : 0;
        const int steps = p.exclude_last ? (h > 0 ? h - 1 : 0) : h;  // SM:286-288
        int maxsteps = steps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int other = __shfl_xor_sync(LEV_FULL_MASK, maxsteps, o);
            maxsteps = other > maxsteps ? other : maxsteps;
        }
        // class of the task = class of its first pair (segments are class-pure)
        int cseg = 0;
        while (t * PPW >= seg_end[cseg]) ++cseg;
        const int cls = LEVG_NCLS - 1 - cseg;
        // stage the hypothesis tokens of the PPW pairs: one contiguous row each
        for (int k = 0; k < PPW; ++k) {
            const int qk = order[t * PPW + k];
            if (qk < 0) continue;
            const int hk = hl_s[qk];
            const int sk = p.exclude_last ? (hk > 0 ? hk - 1 : 0) : hk;
            const int32_t* __restrict__ src = p.hyp_tok + (int64_t)(tile0 + qk) * p.Hp;
            for (int i = lane; i < sk; i += 32) hyp_w[k * Hs + i] = src[i];
        }
        __syncwarp();
        const int* hyp_row = hyp_w + g * Hs;
        int* pref_row = pref_w + g * Hs;
        switch (cls) {
            case 0: levg_run<COUNT, MODE, 8>(p, G, pair, r, h, steps, maxsteps, hyp_row, pref_row); break;
            case 1: levg_run<COUNT, MODE, 12>(p, G, pair, r, h, steps, maxsteps, hyp_row, pref_row); break;
            case 2: levg_run<COUNT, MODE, 16>(p, G, pair, r, h, steps, maxsteps, hyp_row, pref_row); break;
            case 3: levg_run<COUNT, MODE, 20>(p, G, pair, r, h, steps, maxsteps, hyp_row, pref_row); break;
            case 4: levg_run<COUNT, MODE, 24>(p, G, pair, r, h, steps, maxsteps, hyp_row, pref_row); break;
            default: levg_run<COUNT, MODE, 28>(p, G, pair, r, h, steps, maxsteps, hyp_row, pref_row); break;
        }
        __syncwarp();
        if (MODE == LEV_MODE_PREFIX) {
            // SM:279-285 (row 0), 340-346 (rows 1..), 356-386 (scale, norm, tail padding)
            for (int k = 0; k < PPW; ++k) {
                const int qk = order[t * PPW + k];
                if (qk < 0) continue;
                const int rk = rl_s[qk], hk = hl_s[qk];
                const int first_pad = hk + (p.exclude_last ? 0 : 1);
                const int64_t col = (int64_t)(tile0 + qk) * p.out_sn;
                const float rf = (float)rk;
                for (int i = lane; i < p.Hout; i += 32) {
                    float val;
                    if (i >= first_pad) {
                        val = p.padding;
                    } else {
                        const int raw = (i == 0) ? (COUNT ? rk : rk * p.del_i) : pref_w[k * Hs + i];
                        val = (float)raw * p.mult;
                        if (p.norm) val = (rk == 0) ? (i > 0 ? 1.0f : 0.0f) : val / rf;
                    }
                    p.out[(int64_t)i * p.out_si + col] = val;
                }
            }
            __syncwarp();
        }
    }
}

// Returns 1 if the group kernel took the job, 0 if it does not apply (caller falls back
// to the warp-per-pair kernel), < 0 on error.
int lev_launch_group(const LevParams& p, int mode, bool count_mode, cudaStream_t st) {
    if (mode == LEV_MODE_MASK) return 0;
    // small batches: latency matters, one warp per pair (B200LEV_GROUP_MIN_PAIRS
    // overrides the switch-over point; tests use it to drive both kernels)
    int min_pairs = 4096;
    if (const char* e = getenv("B200LEV_GROUP_MIN_PAIRS")) min_pairs = atoi(e);
    if (p.P < min_pairs) return 0;
    int G = 1;
    while (G <= 32 && G * 28 < p.R + 1) G <<= 1;
    if (G > 32 || p.H > 1000 || p.R > 30000) return 0;
    LevGroupGeom geo;
    geo.G = G;
    const int PPW = 32 / G;
    geo.Hs = ((p.H + 2 + 31) / 32) * 32 + G;  // stride == G (mod 32): conflict-free rows
    geo.nbins = LEVG_NCLS * (p.H + 1);
    const int maxc = p.ins_i > p.del_i ? (p.ins_i > p.sub_i ? p.ins_i : p.sub_i)
                                       : (p.del_i > p.sub_i ? p.del_i : p.sub_i);
    geo.allow16 = (!count_mode && p.ins_i >= 0 && p.del_i >= 0 && p.sub_i >= 0 &&
                   (int64_t)maxc * (p.R + p.H + 2) < LEVG_BIG16)
                      ? 1
                      : 0;
    if (const char* e = getenv("B200LEV_GROUP_PACKED16")) geo.allow16 = geo.allow16 && atoi(e);
    const size_t per_warp = (size_t)2 * PPW * geo.Hs * sizeof(int);
    const int nwarps = 8;
    int tile = 512;
    if (tile < 16 * PPW) tile = 16 * PPW;
    geo.tile = tile;
    const size_t smem = sizeof(int) * (geo.nbins + 1 + tile + 2 * LEVG_NCLS * PPW) +
                        sizeof(short) * (2 * tile) + per_warp * nwarps + 16;
    if (smem > 200 * 1024) return 0;
    const int64_t blocks = ((int64_t)p.P + tile - 1) / tile;
#define LEVG_LAUNCH(COUNT_, MODE_)                                                              \
    {                                                                                           \
        auto kern = lev_group_kernel<COUNT_, MODE_>;                                            \
        if (smem > 48 * 1024) {                                                                 \
            cudaError_t e = cudaFuncSetAttribute(                                               \
                kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                  \
            if (e != cudaSuccess) {                                                             \
                lev_set_error("cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e)); \
                return B200LEV_ERR_CUDA;                                                        \
            }                                                                                   \
        }                                                                                       \
        lev_launch(kern, dim3((unsigned)blocks), dim3(32 * nwarps), smem, st, p, geo);          \
    }
    if (!count_mode) {
        if (mode == LEV_MODE_FINAL) LEVG_LAUNCH(false, LEV_MODE_FINAL)
        else LEVG_LAUNCH(false, LEV_MODE_PREFIX)
    } else {
        if (mode == LEV_MODE_FINAL) LEVG_LAUNCH(true, LEV_MODE_FINAL)
        else LEVG_LAUNCH(true, LEV_MODE_PREFIX)
    }
#undef LEVG_LAUNCH
    const int rc = lev_check_cuda("lev_group_kernel");
    return rc ? rc : 1;
}
