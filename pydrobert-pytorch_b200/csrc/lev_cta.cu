// lev_cta.cu -- K2: one CTA per pair, strips pipelined across warps.  For long pairs and
// for batches too small to fill the chip with one warp per pair (north-star item 1b).
//
// Same cell update, right-aligned strips and lane skew as lev_dp.cu, but the strips of ONE
// pair run CONCURRENTLY: warp w owns strips w, w+NW, ...; strip k consumes the boundary
// column (row i, last column of strip k-1) that strip k-1's lane 31 publishes to shared
// memory, a progress counter per boundary (release/acquire through __threadfence_block +
// volatile shared accesses) keeping the consumer at most ~40 rows behind.  All live
// diagonals stay in registers; shared memory holds only the boundary columns and the
// pair's tokens, which are staged with TMA bulk copies (cp.async.bulk -> SASS UBLKCP)
// completing on an mbarrier: two descriptors-free 1-D copies per pair, issued by one thread.
//
// Pairs are pulled from a global counter (ragged lengths -> dynamic balance).  FINAL and
// PREFIX modes; integer costs (cost row or (cost, count) rows) and fp32 costs.
#include <cstdlib>

#include "lev_arith.cuh"

// rows between two publications of a strip's boundary progress (power of two)
#ifndef LEV_CTA_PUB
#define LEV_CTA_PUB 8
#endif
#ifdef LEV_CTA_NOFENCE  // timing experiments only
#define LEV_CTA_FENCE() ((void)0)
#else
#define LEV_CTA_FENCE() __threadfence_block()
#endif

struct LevCtaGeom {
    int NW;    // warps per CTA
    int Smax;  // boundary columns provisioned
    int Hs;    // words per boundary column (H + 64)
    int Rs;    // words reserved for the reference tokens (multiple of 4)
    int Hts;   // words reserved for the hypothesis tokens (multiple of 4)
    int ordered;  // pairs are taken in the order lev_cta_order_kernel wrote (largest first)
    int shift;    // order key = (r * h) >> shift, < 1024
};

// Largest-first order of the pairs (LPT): a counting sort on the cell count r*h quantised
// to 1024 levels.  One CTA; the order lands in the (otherwise unused) slot table.
__global__ void __launch_bounds__(1024) lev_cta_order_kernel(const LevParams p, const LevCtaGeom geo) {
    __shared__ int hist[1024];
    const int tid = threadIdx.x;
    hist[tid] = 0;
    __syncthreads();
    for (int q = tid; q < p.P; q += 1024) {
        const long long key = ((long long)p.ref_len[q / p.ref_group] * p.hyp_len[q]) >> geo.shift;
        atomicAdd(&hist[1023 - (int)(key > 1023 ? 1023 : key)], 1);
    }
    __syncthreads();
    // exclusive scan of 1024 bins: warp scans + one pass over the 32 warp totals
    const int lane = tid & 31, warp = tid >> 5;
    __shared__ int wsum[32];
    const int cnt = hist[tid];
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(LEV_FULL_MASK, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(LEV_FULL_MASK, w, o);
            if (lane >= o) w += t;
        }
        wsum[lane] = w - wsum[lane];
    }
    __syncthreads();
    hist[tid] = wsum[warp] + incl - cnt;
    __syncthreads();
    for (int q = tid; q < p.P; q += 1024) {
        const long long key = ((long long)p.ref_len[q / p.ref_group] * p.hyp_len[q]) >> geo.shift;
        p.gmeta_order[atomicAdd(&hist[1023 - (int)(key > 1023 ? 1023 : key)], 1)] = q;
    }
}

template <typename V, bool COUNT>
struct LevCtaChan {
    static constexpr bool FLT_COST = !std::is_same<V, int>::value && !COUNT;
    static constexpr int NCH = COUNT ? 2 : (FLT_COST ? 3 : 1);
};

template <typename V, bool COUNT, int MODE, int C>
__device__ __forceinline__ void lev_cta_strip(const LevParams& p, const int pair, const int r,
                                              const int h, const int steps, const int k,
                                              const int S, const int* __restrict__ ref_s,
                                              const int* __restrict__ hyp_s, V* __restrict__ bnd,
                                              int* __restrict__ done, const int Hs, const float rcp) {
    constexpr bool IS_INT = std::is_same<V, int>::value;
    constexpr bool FLT_COST = !IS_INT && !COUNT;
    constexpr int NCH = LevCtaChan<V, COUNT>::NCH;
    constexpr int W = 32 * C;
    const int lane = threadIdx.x & 31;
    const V BIG = LevArith<V>::big();
    const V insc = LevArith<V>::ins(p), delc = LevArith<V>::del(p), subc = LevArith<V>::sub(p);
    const bool first = (k == 0), last = (k == S - 1);
    const int jb = r - (S - k) * W;    // boundary column left of lane 0 (< 0 iff first)
    const int j0 = jb + lane * C + 1;  // this lane's first column
    // boundary k-1 is read, boundary k written; channels: value | count or (origin, fl(k*d))
    const V* __restrict__ in_v = bnd + (size_t)(k > 0 ? k - 1 : 0) * NCH * Hs;
    const V* __restrict__ in_a = in_v + Hs;
    const V* __restrict__ in_b = in_v + 2 * Hs;
    V* __restrict__ out_v = bnd + (size_t)k * NCH * Hs;
    V* __restrict__ out_a = out_v + Hs;
    V* __restrict__ out_b = out_v + 2 * Hs;
    (void)in_a; (void)in_b; (void)out_a; (void)out_b;
    V v[C], m[C], jd[C];
    int rt[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int j = j0 + c;
        v[c] = (j >= 0) ? (V)j * delc : BIG;     // SM:258-263
        m[c] = (V)(j >= 0 ? j : 0);              // SM:260
        jd[c] = (j >= 0) ? (V)j * delc : (V)0;
        rt[c] = (j >= 1) ? ref_s[j - 1] : 0;
    }
    (void)m; (void)jd;
    V ob = BIG, oj = (V)0;
    (void)ob; (void)oj;
    V pl_v = first ? BIG : (V)jb * delc;
    V pl_m = (V)(first ? 0 : jb);
    (void)pl_m;
    if (steps <= 0) {
        if (MODE == LEV_MODE_FINAL && last && lane == 31)
            p.out[pair] = lev_finalize((float)(COUNT ? m[C - 1] : v[C - 1]), p, r, h > 0);
        return;
    }
    // Shared memory is addressed through 32-bit shared-space addresses computed once, so
    // that the step loop carries no generic->shared conversions:
    //   row of this lane at step s is i = s - lane; its hypothesis token sits at hyp_a + 4*s;
    //   lane 0 reads row s of the left boundary at in_*_a + 4*s; lane 31 writes row s - 31.
    const lev_saddr hyp_a = lev_saddr_of(hyp_s) - 4u * (unsigned)(1 + lane);
    const lev_saddr in_v_a = lev_saddr_of(in_v + 32), in_a_a = lev_saddr_of(in_a + 32),
                    in_b_a = lev_saddr_of(in_b + 32);
    const lev_saddr out_v_a = lev_saddr_of(out_v + 1), out_a_a = lev_saddr_of(out_a + 1),
                    out_b_a = lev_saddr_of(out_b + 1);
    (void)in_a_a; (void)in_b_a; (void)out_a_a; (void)out_b_a;
    float* __restrict__ orow = p.out + (int64_t)pair * p.out_sn - (int64_t)31 * p.out_si;
    const float rf = (float)r, mult = p.mult;
    const bool norm = p.norm != 0;
    const bool is0 = (lane == 0), is31 = (lane == 31);
    const int* done_in = done + (k > 0 ? k - 1 : 0);
    const int nsteps = steps + 31;
    auto as_v = [](int bits) -> V { return IS_INT ? (V)bits : (V)__int_as_float(bits); };
    auto as_bits = [](V x) -> int { return IS_INT ? (int)x : __float_as_int((float)x); };
    int ht_next = lev_lds32(hyp_a + 4u);  // token of step 1 (software-pipelined one step ahead)

    // one wavefront step; ALL = every lane has a row to update (32 <= s <= steps), which
    // removes the divergent branch from the steady state
    // the three parts of a wavefront step.  flow / publish only act on every LEV_CTA_PUB-th
    // row; the steady state below runs them at fixed places of an unrolled block of rows.
    auto flow = [&](const int s) {  // wait for the next LEV_CTA_PUB rows of the left boundary
        if (!first && (s & (LEV_CTA_PUB - 1)) == 1 && s <= steps) {
            const int need = min(s + LEV_CTA_PUB - 1, steps);
            while (lev_ld_volatile_shared(done_in) < need) __nanosleep(32);
            LEV_CTA_FENCE();
        }
    };
    auto publish = [&](const int s) {  // rows <= s - 31 of this strip's boundary are out
        if (!last) {
            const int i31 = s - 31;
            if (i31 >= 1 && ((i31 & (LEV_CTA_PUB - 1)) == 0 || i31 == steps)) {
                __syncwarp();
                if (is31) {
                    LEV_CTA_FENCE();
                    lev_st_volatile_shared(done + k, i31);
                }
            }
        }
    };
    // ALL = every lane has a row to update (32 <= s <= steps): no divergent branch.  Lane 0's
    // left neighbour is the boundary column: EVERY lane reads that word (one broadcast LDS) and
    // selects, instead of lane 0 branching away to fetch it.
    // PRE: the boundary word(s) and the token of this row were loaded ahead (steady-state blocks)
    auto core = [&](auto ALL, const int s, auto PRE, const int pre_v, const int pre_a, const int pre_b,
                    const int pre_ht) {
        constexpr bool all_active = decltype(ALL)::value;
        constexpr bool pre = decltype(PRE)::value;
        (void)pre_a; (void)pre_b;
        const V sh_v = __shfl_up_sync(LEV_FULL_MASK, v[C - 1], 1);
        // past the last row lane 0 is idle; whatever it reads there is never used
        const V bnd_v = first ? BIG : as_v(pre ? pre_v : lev_lds32_sync(in_v_a + 4u * (unsigned)s));
        const V hand_v = is0 ? bnd_v : sh_v;
        const V diag_v = pl_v;
        pl_v = hand_v;
        V hand_m = (V)0, diag_m = (V)0, hand_ob = BIG, hand_oj = (V)0;
        if (COUNT) {
            const V sh_m = __shfl_up_sync(LEV_FULL_MASK, m[C - 1], 1);
            const V bnd_m = first ? (V)0 : as_v(pre ? pre_a : lev_lds32_sync(in_a_a + 4u * (unsigned)s));
            hand_m = is0 ? bnd_m : sh_m;
            diag_m = pl_m;
            pl_m = hand_m;
        }
        if (FLT_COST) {
            const V sh_ob = __shfl_up_sync(LEV_FULL_MASK, ob, 1);
            const V sh_oj = __shfl_up_sync(LEV_FULL_MASK, oj, 1);
            const V bnd_ob = first ? BIG : as_v(pre ? pre_a : lev_lds32_sync(in_a_a + 4u * (unsigned)s));
            const V bnd_oj = first ? (V)0 : as_v(pre ? pre_b : lev_lds32_sync(in_b_a + 4u * (unsigned)s));
            hand_ob = is0 ? bnd_ob : sh_ob;
            hand_oj = is0 ? bnd_oj : sh_oj;
        }
        (void)hand_m; (void)diag_m; (void)hand_ob; (void)hand_oj;
        int ht;
        if (pre) {
            ht = pre_ht;
        } else {
            ht = ht_next;
            ht_next = lev_lds32(hyp_a + 4u * (unsigned)(s + 1));
        }
        const bool active = all_active || (unsigned)(s - lane - 1) < (unsigned)steps;
        if (active) {
            if (COUNT) {  // SM:292-314
                V dc = diag_v, dm = diag_m, lc = hand_v, lm = hand_m;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const V uc = v[c], um = m[c];
                    // (integer cells: "differs" from one DPX add-and-clamp, the cost and count terms by
                    // multiply / add instead of two selects)
                    const V neq01 = IS_INT ? (V)(int)__viaddmin_u32((unsigned)rt[c], 0u - (unsigned)ht, 1u)
                                           : (rt[c] != ht ? (V)1 : (V)0);
                    const V sub_c = IS_INT ? dc + neq01 * subc : dc + (neq01 != (V)0 ? subc : (V)0);
                    const V ins_c = uc + insc;
                    const bool ps = ins_c >= sub_c;
                    V cc = ps ? sub_c : ins_c;
                    V mm = ps ? dm + neq01 : um + (V)1;
                    const V del_c = lc + delc;
                    const bool keep = del_c >= cc;
                    cc = keep ? cc : del_c;
                    mm = keep ? mm : lm + (V)1;
                    dc = uc;
                    dm = um;
                    lc = cc;
                    lm = mm;
                    v[c] = cc;
                    m[c] = mm;
                }
            } else if (FLT_COST) {  // SM:290-293, 316-317; deletion run tracked by origin
                V dg = diag_v, obr = hand_ob, ojr = hand_oj;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const V up = v[c];
                    const V a = up + insc;
                    const V sb = dg + ((rt[c] != ht) ? subc : (V)0);
                    const V t = a < sb ? a : sb;
                    const V cand = obr + (jd[c] - ojr);
                    const bool fresh = !(cand < t);
                    v[c] = fresh ? t : cand;
                    obr = fresh ? t : obr;
                    ojr = fresh ? jd[c] : ojr;
                    dg = up;
                }
                ob = obr;
                oj = ojr;
            } else {
                int dg = (int)diag_v, lf = (int)hand_v;
                const unsigned nht = 0u - (unsigned)ht;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int up = (int)v[c];
                    // "tokens differ" as 0 / 1 from one DPX add-and-clamp, times the cost on the
                    // FMA pipe: 3 ALU-pipe instructions per cell instead of 4
                    const int sb = (int)(__viaddmin_u32((unsigned)rt[c], nht, 1u) * (unsigned)subc) + dg;
                    const int t = __viaddmin_s32(up, (int)insc, sb);
                    lf = __viaddmin_s32(lf, (int)delc, t);
                    dg = up;
                    v[c] = (V)lf;
                }
            }
        }
        if (is31 && active) {
            if (!last) {
                lev_sts32(out_v_a + 4u * (unsigned)s, as_bits(v[C - 1]));
                if (COUNT) lev_sts32(out_a_a + 4u * (unsigned)s, as_bits(m[C - 1]));
                if (FLT_COST) {
                    lev_sts32(out_a_a + 4u * (unsigned)s, as_bits(ob));
                    lev_sts32(out_b_a + 4u * (unsigned)s, as_bits(oj));
                }
            } else if (MODE == LEV_MODE_PREFIX) {  // SM:340-346, 356-378
                float val = (float)(COUNT ? m[C - 1] : v[C - 1]) * mult;
                if (norm) {
                    if (r == 0) {
                        val = 1.0f;
                    } else {
                        const float q0 = __fmul_rn(val, rcp);
                        val = __fmaf_rn(__fmaf_rn(-rf, q0, val), rcp, q0);  // == val / rf
                    }
                }
                orow[(int64_t)s * p.out_si] = val;
            }
        }
    };
    auto step = [&](auto ALL, const int s) {
        flow(s);
        core(ALL, s, std::false_type(), 0, 0, 0, 0);
        publish(s);
    };
#ifdef LEV_CTA_CLOCK
    const long long clk0 = clock64();
#endif
    int s = 1;
    for (; s <= 32 && s <= nsteps; ++s) {  // ramp-up (and row 32, so that blocks start at s = 1 mod PUB)
        if (s <= 31 || s > steps)
            step(std::false_type(), s);
        else
            step(std::true_type(), s);
    }
    // steady state in blocks of LEV_CTA_PUB rows starting at s = 1 (mod PUB): one wait for the
    // left boundary at the top, the progress word published after the row with s - 31 = 0 (mod
    // PUB) -- fixed places, so the rows themselves carry no flow-control tests
    static_assert(LEV_CTA_PUB == 8, "the block below assumes 31 = PUB - 1 (mod PUB)");
    // (fetching the block's boundary words and tokens up front instead of one shared-memory
    // load per row was measured too: no gain, 17 more registers)
    for (; s + LEV_CTA_PUB - 1 <= steps; s += LEV_CTA_PUB) {
        flow(s);
#pragma unroll
        for (int u = 0; u < LEV_CTA_PUB; ++u) {
            core(std::true_type(), s + u, std::false_type(), 0, 0, 0, 0);
            if (u == LEV_CTA_PUB - 2) publish(s + u);  // s + u - 31 = 0 (mod PUB)
        }
    }
    for (; s <= steps; ++s) step(std::true_type(), s);    // what is left of the steady state
    for (; s <= nsteps; ++s) step(std::false_type(), s);  // ramp-down
#ifdef LEV_CTA_CLOCK
    if (lane == 0 && pair == 0) {  // debug: cycles of this strip's step loop
        p.gmeta[2 + 2 * k] = (int)(clock64() - clk0);
        p.gmeta[3 + 2 * k] = nsteps;
    }
#endif
    if (MODE == LEV_MODE_FINAL && last && lane == 31)  // SM:390-405
        p.out[pair] = lev_finalize((float)(COUNT ? m[C - 1] : v[C - 1]), p, r, h > 0);
}

template <typename V, bool COUNT, int MODE>
__global__ void __launch_bounds__(128, 4) lev_cta_kernel(const LevParams p, const LevCtaGeom geo) {
    LEV_DYN_SMEM(int, smem);
    if (*p.wide_flag & B200LEV_FLAG_WIDE_TOKENS) return;  // stand-by lev_warp_kernel takes over
    constexpr int NCH = LevCtaChan<V, COUNT>::NCH;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = geo.NW;
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(smem);  // 16 bytes reserved
    int* ref_s = smem + 4;
    int* hyp_s = ref_s + geo.Rs;
    int* done = hyp_s + geo.Hts;
    V* bnd = reinterpret_cast<V*>(done + ((geo.Smax + 3) & ~3));
    __shared__ int cur_pair;
    int* counter = const_cast<int*>(p.wide_flag) + 3;
    if (tid == 0) lev_mbar_init(mbar, 1);
    unsigned parity = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) cur_pair = atomicAdd(counter, 1);
        __syncthreads();
        if (cur_pair >= p.P) break;
        const int pair = geo.ordered ? p.gmeta_order[cur_pair] : cur_pair;  // longest first
        const int refcol = pair / p.ref_group;
        const int r = p.ref_len[refcol], h = p.hyp_len[pair];
        const int steps = p.exclude_last ? (h > 0 ? h - 1 : 0) : h;  // SM:286-288
        if (tid == 0 && r == 0 && p.norm && p.flags != nullptr)
            atomicOr(p.flags, B200LEV_FLAG_EMPTY_REF);  // SM:360-366, 397-404
        // ---- stage the pair's tokens: two TMA bulk copies completing on the mbarrier ----
        const unsigned rbytes = (unsigned)((r * 4 + 15) & ~15), hbytes = (unsigned)((steps * 4 + 15) & ~15);
        if (tid == 0 && rbytes + hbytes > 0) {
            lev_mbar_expect_tx(mbar, rbytes + hbytes);
            if (rbytes) lev_bulk_g2s(ref_s, p.ref_tok + (int64_t)refcol * p.Rp, rbytes, mbar);
            if (hbytes) lev_bulk_g2s(hyp_s, p.hyp_tok + (int64_t)pair * p.Hp, hbytes, mbar);
        }
        for (int q = tid; q < geo.Smax; q += blockDim.x) done[q] = 0;
        // PREFIX: row 0 and the padded tail do not depend on the DP (SM:279-285, 379-386)
        if (MODE == LEV_MODE_PREFIX) {
            const int first_pad = h + (p.exclude_last ? 0 : 1);
            if (tid == 0 && first_pad > 0 && p.Hout > 0) {
                const float v0 = COUNT ? (float)r : (float)((V)r * LevArith<V>::del(p));
                p.out[(int64_t)pair * p.out_sn] = lev_finalize(v0, p, r, false);
            }
            for (int i = first_pad + tid; i < p.Hout; i += blockDim.x)
                p.out[(int64_t)i * p.out_si + (int64_t)pair * p.out_sn] = p.padding;
        }
        __syncthreads();
        if (rbytes + hbytes > 0) {
            lev_mbar_wait(mbar, parity);
            parity ^= 1;
        }
        const float rcp = r > 0 ? __frcp_rn((float)r) : 0.0f;
        // smallest C whose NW concurrent strips cover the row; C = 8 beyond (several rounds)
        const int cols = r + 1;
        int S;
#define LEV_CTA_RUN(C_)                                                                          \
    S = (cols + 32 * C_ - 1) / (32 * C_);                                                        \
    for (int k = warp; k < S; k += NW)                                                           \
        lev_cta_strip<V, COUNT, MODE, C_>(p, pair, r, h, steps, k, S, ref_s, hyp_s, bnd, done,   \
                                          geo.Hs, rcp);
        if (cols <= 32 * NW) { LEV_CTA_RUN(1) }
        else if (cols <= 64 * NW) { LEV_CTA_RUN(2) }
        else if (cols <= 128 * NW) { LEV_CTA_RUN(4) }
        else if (cols <= 256 * NW) { LEV_CTA_RUN(8) }
        else { LEV_CTA_RUN(16) }
#undef LEV_CTA_RUN
        (void)lane;
    }
}

template <typename V, bool COUNT, int MODE>
static int lev_cta_launch_one(const LevParams& p, cudaStream_t st) {
    constexpr int NCH = LevCtaChan<V, COUNT>::NCH;
    LevCtaGeom geo;
    // enough warps to cover a row with C = 1 strips, at most 4: with C up to 16 columns per
    // lane four concurrent strips span 2048 columns, and 2-3 CTAs share an SM
    int NW = (p.R + 1 + 31) / 32;
    // (8 / 16 / 32 warps per pair were measured on config 5: 8 ties with 4 on 256 pairs and loses
    // 30 % on 1184, more warps lose everywhere -- the strips of a pair run at the pace of the
    // first one, so extra warps only add pipeline fill)
    const int maxw = 4;
    geo.NW = NW < 2 ? 2 : (NW > maxw ? maxw : NW);
    geo.Smax = (p.R + 1 + 511) / 512;
    if (geo.Smax < geo.NW) geo.Smax = geo.NW;
    // largest-first order when the batch is ragged enough to matter
    geo.ordered = p.P > 1 ? 1 : 0;
    geo.shift = 0;
    while ((((long long)p.R * p.H) >> geo.shift) > 1023) ++geo.shift;
    geo.Hs = p.H + 64;
    geo.Rs = (int)p.Rp + 4;
    geo.Hts = (int)p.Hp + 4;
    const size_t smem = sizeof(int) * ((size_t)4 + geo.Rs + geo.Hts + ((geo.Smax + 3) & ~3) +
                                       (size_t)geo.Smax * NCH * geo.Hs);
    if (smem > 200 * 1024) return 0;
    auto kern = lev_cta_kernel<V, COUNT, MODE>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            lev_set_error("cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
            return B200LEV_ERR_CUDA;
        }
    }
    if (cudaMemsetAsync(const_cast<int*>(p.wide_flag) + 3, 0, sizeof(int), st) != cudaSuccess)
        return lev_check_cuda("memset");
    if (geo.ordered) lev_launch(lev_cta_order_kernel, dim3(1), dim3(1024), 0, st, p, geo);
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    const int by_threads = 2048 / (32 * geo.NW);
    if (per_sm > by_threads) per_sm = by_threads;
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = (int64_t)148 * per_sm;
    if (blocks > p.P) blocks = p.P;
    lev_prof_begin(LEV_PROF_DP, st);
    lev_launch(kern, dim3((unsigned)blocks), dim3(32 * geo.NW), smem, st, p, geo);
    lev_prof_end(LEV_PROF_DP, st);
    const int rc = lev_check_cuda("lev_cta_kernel");
    return rc ? rc : 1;
}

// Returns 1 if the CTA-per-pair kernel took the job, 0 if it does not apply, < 0 on error.
int lev_launch_cta(const LevParams& p, int mode, bool count_mode, bool float_path, cudaStream_t st) {
    if (mode == LEV_MODE_MASK) return 0;
    // worth it when rows are long: measured crossover against the warp-per-pair kernel
    // (scripts/micro/k1_vs_k2.py, square pairs): R >= 400 for P <= 128, R >= 800 for
    // P <= 1024, R >= 1600 at P = 2048; below R = 300 one warp per pair always wins (a
    // 512 x 101 batch: 42 us against 88).  B200LEV_CTA_KERNEL=0/1 forces the choice (tests).
    bool use = (p.R >= 640 && (int64_t)p.P <= 1536) || (p.R >= 384 && p.P <= 256) || p.R >= 1400;
    if (const char* e = getenv("B200LEV_CTA_KERNEL")) use = atoi(e) != 0;
    if (!use) return 0;
    if (!float_path) {
        if (!count_mode) {
            if (mode == LEV_MODE_FINAL) return lev_cta_launch_one<int, false, LEV_MODE_FINAL>(p, st);
            return lev_cta_launch_one<int, false, LEV_MODE_PREFIX>(p, st);
        }
        if (mode == LEV_MODE_FINAL) return lev_cta_launch_one<int, true, LEV_MODE_FINAL>(p, st);
        return lev_cta_launch_one<int, true, LEV_MODE_PREFIX>(p, st);
    }
    if (!count_mode) {
        if (mode == LEV_MODE_FINAL) return lev_cta_launch_one<float, false, LEV_MODE_FINAL>(p, st);
        return lev_cta_launch_one<float, false, LEV_MODE_PREFIX>(p, st);
    }
    if (mode == LEV_MODE_FINAL) return lev_cta_launch_one<float, true, LEV_MODE_FINAL>(p, st);
    return lev_cta_launch_one<float, true, LEV_MODE_PREFIX>(p, st);
}
