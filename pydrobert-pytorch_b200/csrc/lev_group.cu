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
//   * persistent CTAs pull TILES of consecutive pairs from a global counter and bucket
//     each tile in shared memory by (column class C, hypothesis length) with a counting
//     sort; a warp then runs 32/G pairs (2 x 32/G in the packed path) of the SAME class
//     and near-equal length in lock step, largest first, from a shared-memory work queue.
//     Lane utilisation is (r+1)/(G*C) instead of (r+1)/(32*C), and no step is spent on a
//     pair that has already finished;
//   * the hypothesis rows of the NEXT task are fetched with cp.async (LDGSTS) into the
//     other half of a double buffer while the current task runs;
//   * prefix values are parked in shared memory by the one lane that owns column r and
//     written out after the task by all lanes (scale, IEEE division by the reference
//     length, padding fill: SM:356-386), off the DP loop.
//
// PACKED path: when K0 found all tokens inside a 65536-wide window (their low 16 bits are
// then injective) and costs/lengths keep every value below LEVG_BIG16, TWO pairs share
// each register, one per 16-bit half, and the 2-wide DPX instructions do the work.  Per 2
// cells: LOP3 (token xor), VIMNMX.U16x2 (-> 0/1 per half), IMAD (diag + neq*sub, FMA
// pipe), 2 x VIADDMNMX.S16x2.  Otherwise the 32-bit path of lev_dp.cu's cell update runs.
//
// Integer costs only (cost row, or (cost, count) rows for the error-rate family);
// FINAL and PREFIX modes.  Everything else stays on lev_dp.cu.
#include <cstdlib>

#include "lev_common.cuh"

#define LEVG_NCLS 6  // column classes C = 8, 12, ..., 28
// "infinity" of the packed path: BIG16 + (largest reachable value) must stay < 2^15
#define LEVG_BIG16 16000

struct LevGroupGeom {
    int G;         // lanes per pair
    int tile;      // pairs per tile
    int ntiles;
    int Hs;        // words per prefix row (>= H + 2)
    int row_words; // 32-bit words per staged hypothesis row
    int nbins;     // LEVG_NCLS * (H + 1)
    int allow16;   // costs and lengths admit the packed 2 x int16 DPX path
    int nwarps;
};

// smallest class whose strip covers columns 0..r
__device__ __forceinline__ int levg_class_of(int r, int G) {
    const int need = (r + G) / G;  // ceil((r + 1) / G)
    const int cls = (need - 8 + 3) >> 2;
    return cls < 0 ? 0 : cls;
}

// IEEE-754 correctly rounded a / b from the correctly rounded reciprocal y = RN(1/b)
// (Markstein): q0 = RN(a*y), r = a - b*q0 (exact in an FMA), q = RN(q0 + r*y).  Valid for
// the operands here (0 <= a < 2^24, 1 <= b < 2^24 - 1, no exponent corner cases) and
// 3 instructions per element once y is hoisted out of the row loop.
__device__ __forceinline__ float levg_div(float a, float b, float y) {
    const float q0 = __fmul_rn(a, y);
    const float r = __fmaf_rn(-b, q0, a);
    return __fmaf_rn(r, y, q0);
}

// ---------------------------------------------------------------------------------------
// 32-bit path: one pair per lane group
// ---------------------------------------------------------------------------------------
template <bool COUNT, int MODE, int C>
__device__ __forceinline__ void levg_run32(const LevParams& p, const int G, const int pair,
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

// ---------------------------------------------------------------------------------------
// packed path: two pairs per lane group, one per 16-bit half.  Both run `maxsteps` rows;
// rows past a pair's own length are never read back.  Cost row only.
// ---------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ void levg_run16(const LevParams& p, const int G, const int pairA,
                                           const int rA, const int pairB, const int rB,
                                           const int maxsteps,
                                           const unsigned short* __restrict__ hypA,
                                           const unsigned short* __restrict__ hypB,
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
            const unsigned ht = (unsigned)hypA[i - 1] | ((unsigned)hypB[i - 1] << 16);
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

template <bool COUNT, int MODE, bool PACKED>
__global__ void __launch_bounds__(256, 2) lev_group_kernel(const LevParams p, const LevGroupGeom geo) {
    LEV_DYN_SMEM(int, smem);
    // Which of the two builds of this kernel runs is decided on the device from what K0
    // found in the tokens (no host round trip): wider than int32 -> neither (the 64-bit
    // compare path of lev_warp_kernel, enqueued right behind, takes the batch); inside a
    // 65536-wide window and small values -> PACKED; else the 32-bit build.
    if (*p.wide_flag & B200LEV_FLAG_WIDE_TOKENS) return;
    {
        // [1] = max(u), [2] = max(~u), u = token + 2^31 (lev_pack.cu)
        const unsigned umax = (unsigned)p.wide_flag[1], umin = ~(unsigned)p.wide_flag[2];
        const bool narrow = !COUNT && geo.allow16 && (umax < umin || umax - umin < 65536u);
        if (narrow != PACKED) return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = geo.G, PPW = 32 / G, TILE = geo.tile, Hs = geo.Hs, H1 = p.H + 1;
    const int PPT = PACKED ? 2 * PPW : PPW;  // pairs per task
    const int RW = geo.row_words;
    // shared-memory carve-up
    int* hist = smem;                                                               // [nbins + 1]
    int* order = hist + geo.nbins + 1;                                              // [TILE + NCLS*PPT]
    short* rl_s = reinterpret_cast<short*>(order + TILE + LEVG_NCLS * PPT);        // [TILE]
    short* hl_s = rl_s + TILE;                                                      // [TILE]
    int* wbase = reinterpret_cast<int*>(hl_s + TILE) + ((TILE & 1) ? 1 : 0);
    wbase = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(wbase) + 15) & ~(uintptr_t)15);
    const int per_warp = 2 * PPT * RW + PPW * Hs;
    int* stage_w = wbase + (size_t)warp * per_warp;  // [2][PPT][RW]   hypothesis rows
    int* pref_w = stage_w + 2 * PPT * RW;            // [PPW][Hs]      prefix values
    __shared__ int next_task, ntasks, seg_end[LEVG_NCLS], cur_tile;
    int* tile_counter = const_cast<int*>(p.wide_flag) + 3;
    const int g = lane / G;

    for (;;) {
        // ---- 0. next tile (persistent CTAs, global counter) ----
        __syncthreads();
        if (tid == 0) cur_tile = atomicAdd(tile_counter, 1);
        __syncthreads();
        const int tile = cur_tile;
        if (tile >= geo.ntiles) break;
        const int tile0 = tile * TILE;
        const int ntile = min(TILE, p.P - tile0);
        // ---- 1. lengths, classes, histogram over (class desc, hyp length desc) ----
        for (int b = tid; b <= geo.nbins; b += blockDim.x) hist[b] = 0;
        for (int q = tid; q < TILE + LEVG_NCLS * PPT; q += blockDim.x) order[q] = -1;
        if (tid == 0) next_task = 0;
        __syncthreads();
        for (int q = tid; q < ntile; q += blockDim.x) {
            const int pair = tile0 + q;
            const int r = p.ref_len[pair / p.ref_group];
            const int h = p.hyp_len[pair];
            rl_s[q] = (short)r;
            hl_s[q] = (short)h;
            if (r == 0 && p.norm && p.flags != nullptr)
                atomicOr(p.flags, B200LEV_FLAG_EMPTY_REF);  // SM:360-366, 397-404
            atomicAdd(&hist[(LEVG_NCLS - 1 - levg_class_of(r, G)) * H1 + (p.H - h)], 1);
        }
        __syncthreads();
        // ---- 2. exclusive scan (one warp), class segments padded to whole tasks ----
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
            const int pos = atomicAdd(
                &hist[(LEVG_NCLS - 1 - levg_class_of(rl_s[q], G)) * H1 + (p.H - hl_s[q])], 1);
            order[pos] = q;
        }
        __syncthreads();

        // ---- 4. work queue; hypothesis rows of the next task prefetched with cp.async ----
        auto claim = [&]() {
            int t = 0;
            if (lane == 0) t = atomicAdd(&next_task, 1);
            return __shfl_sync(LEV_FULL_MASK, t, 0);
        };
        auto stage = [&](int t, int buf) {
            // 16-byte chunks of the PPT token rows (uint16 rows when PACKED, int32 otherwise)
            const int chunks_per_row = RW / 4;
            char* dst0 = reinterpret_cast<char*>(stage_w + buf * PPT * RW);
            for (int c = lane; c < PPT * chunks_per_row; c += 32) {
                const int k = c / chunks_per_row, ch = c - k * chunks_per_row;
                const int qk = order[t * PPT + k];
                if (qk < 0) continue;
                const int hk = hl_s[qk];
                const int sk = p.exclude_last ? (hk > 0 ? hk - 1 : 0) : hk;
                const int bytes = sk * (PACKED ? 2 : 4);
                if (ch * 16 >= bytes) continue;
                const char* src = PACKED
                    ? reinterpret_cast<const char*>(p.hyp_tok16 + (int64_t)(tile0 + qk) * p.Hp16)
                    : reinterpret_cast<const char*>(p.hyp_tok + (int64_t)(tile0 + qk) * p.Hp);
                lev_cp_async16(dst0 + (size_t)k * RW * 4 + ch * 16, src + ch * 16);
            }
        };
        int cur = claim(), buf = 0;
        if (cur < ntasks) stage(cur, buf);
        lev_cp_async_commit();
        while (cur < ntasks) {
            const int nxt = claim();
            if (nxt < ntasks) stage(nxt, buf ^ 1);
            lev_cp_async_commit();
            lev_cp_async_wait<1>();  // everything but the newest group: this task's rows landed
            __syncwarp();
            const int t = cur;
            int cseg = 0;
            while (t * PPT >= seg_end[cseg]) ++cseg;  // class segments are task-pure
            const int cls = LEVG_NCLS - 1 - cseg;
            const int* rows = stage_w + buf * PPT * RW;
            if (PACKED) {
                // group g runs pairs 2g (low halves) and 2g+1 (high halves) of the task
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
                const unsigned short* hA_row = reinterpret_cast<const unsigned short*>(rows + (2 * g) * RW);
                const unsigned short* hB_row = reinterpret_cast<const unsigned short*>(rows + (2 * g + 1) * RW);
                unsigned* prow = reinterpret_cast<unsigned*>(pref_w) + g * Hs;
                const int pA = qA >= 0 ? tile0 + qA : -1, pB = qB >= 0 ? tile0 + qB : -1;
                switch (cls) {
                    case 0: levg_run16<8>(p, G, pA, rA, pB, rB, maxsteps, hA_row, hB_row, prow); break;
                    case 1: levg_run16<12>(p, G, pA, rA, pB, rB, maxsteps, hA_row, hB_row, prow); break;
                    case 2: levg_run16<16>(p, G, pA, rA, pB, rB, maxsteps, hA_row, hB_row, prow); break;
                    case 3: levg_run16<20>(p, G, pA, rA, pB, rB, maxsteps, hA_row, hB_row, prow); break;
                    case 4: levg_run16<24>(p, G, pA, rA, pB, rB, maxsteps, hA_row, hB_row, prow); break;
                    default: levg_run16<28>(p, G, pA, rA, pB, rB, maxsteps, hA_row, hB_row, prow); break;
                }
            } else {
                const int q = order[t * PPT + g];
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
                const int* hyp_row = rows + g * RW;
                int* prow = pref_w + g * Hs;
                switch (cls) {
                    case 0: levg_run32<COUNT, MODE, 8>(p, G, pair, r, h, steps, maxsteps, hyp_row, prow); break;
                    case 1: levg_run32<COUNT, MODE, 12>(p, G, pair, r, h, steps, maxsteps, hyp_row, prow); break;
                    case 2: levg_run32<COUNT, MODE, 16>(p, G, pair, r, h, steps, maxsteps, hyp_row, prow); break;
                    case 3: levg_run32<COUNT, MODE, 20>(p, G, pair, r, h, steps, maxsteps, hyp_row, prow); break;
                    case 4: levg_run32<COUNT, MODE, 24>(p, G, pair, r, h, steps, maxsteps, hyp_row, prow); break;
                    default: levg_run32<COUNT, MODE, 28>(p, G, pair, r, h, steps, maxsteps, hyp_row, prow); break;
                }
            }
            __syncwarp();
            // ---- epilogue (SM:279-285, 340-346, 356-386 / 390-405) ----
            // 32/PPT lanes per pair, each striding over that pair's rows: the per-pair
            // set-up (lengths, reciprocal, output column) is done once per lane
            if (MODE == LEV_MODE_PREFIX || PACKED) {
                const int LPP = PPT >= 32 ? 1 : 32 / PPT;  // lanes per pair
                for (int kk = lane; kk < PPT * LPP; kk += 32) {
                    const int k = kk / LPP, sub = kk - k * LPP;
                    const int qk = order[t * PPT + k];
                    if (qk < 0) continue;
                    const int rk = rl_s[qk], hk = hl_s[qk];
                    const unsigned* row =
                        reinterpret_cast<const unsigned*>(pref_w) + (PACKED ? (k >> 1) : k) * Hs;
                    const int sh16 = PACKED ? (k & 1) * 16 : 0;
                    const unsigned msk = PACKED ? 0xffffu : 0xffffffffu;
                    const float rf = (float)rk;
                    const float y = rk > 0 ? __frcp_rn(rf) : 0.0f;
                    const int row0 = COUNT ? rk : rk * p.del_i;
                    const bool norm = p.norm != 0;
                    if (MODE == LEV_MODE_PREFIX) {
                        const int first_pad = hk + (p.exclude_last ? 0 : 1);
                        float* o = p.out + (int64_t)(tile0 + qk) * p.out_sn + (int64_t)sub * p.out_si;
                        const int64_t ostep = (int64_t)LPP * p.out_si;
                        for (int i = sub; i < p.Hout; i += LPP, o += ostep) {
                            float val = p.padding;
                            if (i < first_pad) {
                                const int raw = (i == 0) ? row0 : (int)((row[i] >> sh16) & msk);
                                val = __fmul_rn((float)raw, p.mult);
                                if (norm) val = (rk == 0) ? (i > 0 ? 1.0f : 0.0f) : levg_div(val, rf, y);
                            }
                            *o = val;
                        }
                    } else if (sub == 0) {  // PACKED FINAL: the value parked at row h
                        const int raw = (hk == 0) ? row0 : (int)((row[hk] >> sh16) & msk);
                        float val = __fmul_rn((float)raw, p.mult);
                        if (norm) val = (rk == 0) ? (hk > 0 ? 1.0f : 0.0f) : levg_div(val, rf, y);
                        p.out[tile0 + qk] = val;
                    }
                }
            }
            __syncwarp();
            cur = nxt;
            buf ^= 1;
        }
        lev_cp_async_wait<0>();
    }
}

// geometry + shared-memory footprint of one build; false if it cannot keep 2 CTAs per SM
static bool levg_geometry(const LevParams& p, bool packed, LevGroupGeom* geo, size_t* smem) {
    const int PPW = 32 / geo->G, PPT = packed ? 2 * PPW : PPW;
    // staged hypothesis rows: uint16 (packed) or int32; whole 16-byte chunks per row
    geo->row_words = packed ? (int)((p.H + 7) / 8) * 4 : (int)((p.H + 3) / 4) * 4;
    if (geo->row_words < 4) geo->row_words = 4;
    geo->Hs = ((p.H + 2 + 3) / 4) * 4;
    geo->tile = 512;
    if (geo->tile < 16 * PPT) geo->tile = 16 * PPT;
    geo->ntiles = (int)(((int64_t)p.P + geo->tile - 1) / geo->tile);
    geo->nwarps = 8;
    const size_t per_warp = sizeof(int) * ((size_t)2 * PPT * geo->row_words + (size_t)PPW * geo->Hs);
    *smem = sizeof(int) * (geo->nbins + 1 + geo->tile + LEVG_NCLS * PPT) +
            sizeof(short) * (2 * geo->tile) + per_warp * geo->nwarps + 64;
    return *smem <= 110 * 1024;
}

template <bool COUNT, int MODE, bool PACKED>
static int levg_launch_one(const LevParams& p, LevGroupGeom geo, cudaStream_t st) {
    size_t smem = 0;
    if (!levg_geometry(p, PACKED, &geo, &smem)) return 0;
    auto kern = lev_group_kernel<COUNT, MODE, PACKED>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            lev_set_error("cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
            return B200LEV_ERR_CUDA;
        }
    }
    int blocks = 148 * 2;
    if (blocks > geo.ntiles) blocks = geo.ntiles;
    lev_launch(kern, dim3((unsigned)blocks), dim3(32 * geo.nwarps), smem, st, p, geo);
    const int rc = lev_check_cuda("lev_group_kernel");
    return rc ? rc : 1;
}

// Returns 1 if the group kernels took the job, 0 if they do not apply (caller falls back
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
    memset(&geo, 0, sizeof(geo));
    geo.G = G;
    geo.nbins = LEVG_NCLS * (p.H + 1);
    const int maxc = p.ins_i > p.del_i ? (p.ins_i > p.sub_i ? p.ins_i : p.sub_i)
                                       : (p.del_i > p.sub_i ? p.del_i : p.sub_i);
    geo.allow16 = (!count_mode && p.ins_i >= 0 && p.del_i >= 0 && p.sub_i >= 0 &&
                   (int64_t)maxc * (p.R + p.H + 2) < LEVG_BIG16)
                      ? 1
                      : 0;
    if (const char* e = getenv("B200LEV_GROUP_PACKED16")) geo.allow16 = geo.allow16 && atoi(e);
    // the 32-bit build must be launchable (it is the one that runs when the tokens turn
    // out not to fit 16 bits); the packed build is optional
    LevGroupGeom tmp = geo;
    size_t smem = 0;
    if (!levg_geometry(p, false, &tmp, &smem)) return 0;
    tmp = geo;
    if (geo.allow16 && !levg_geometry(p, true, &tmp, &smem)) geo.allow16 = 0;
    // the persistent CTAs pull tiles from state word 3: rewind it for this launch
    if (cudaMemsetAsync(const_cast<int*>(p.wide_flag) + 3, 0, sizeof(int), st) != cudaSuccess)
        return lev_check_cuda("memset");
    // both builds are enqueued; the device-side token range decides which one works
    int rc;
#define LEVG_BOTH(COUNT_, MODE_)                                                   \
    rc = levg_launch_one<COUNT_, MODE_, false>(p, geo, st);                        \
    if (rc <= 0) return rc;                                                        \
    if (!COUNT_ && geo.allow16) {                                                  \
        rc = levg_launch_one<COUNT_, MODE_, true>(p, geo, st);                     \
        if (rc <= 0) return rc < 0 ? rc : B200LEV_ERR_UNSUPPORTED;                 \
    }
    if (!count_mode) {
        if (mode == LEV_MODE_FINAL) { LEVG_BOTH(false, LEV_MODE_FINAL) }
        else { LEVG_BOTH(false, LEV_MODE_PREFIX) }
    } else {
        if (mode == LEV_MODE_FINAL) { LEVG_BOTH(true, LEV_MODE_FINAL) }
        else { LEVG_BOTH(true, LEV_MODE_PREFIX) }
    }
#undef LEVG_BOTH
    return 1;
}
