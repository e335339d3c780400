// lev_common.cuh -- shared declarations for the Levenshtein kernels.
//
// "SM" in comments is src/pydrobert/torch/_string.py of the reference.
#pragma once
#include "simt.h"

#include "../../include/b200lev.h"

enum LevMode { LEV_MODE_FINAL = 0, LEV_MODE_PREFIX = 1, LEV_MODE_MASK = 2 };

// "Infinity" for the int32 DP: virtual columns left of column 0 hold it so that
// column 0 needs no special case.  Costs are bounded on the host so that BIG plus
// any reachable path cost stays below 2^31.
#define LEV_BIG_I32 (1 << 29)

static inline int64_t lev_round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- group kernel bucketing (lev_group.cu, histogram built by lev_pack.cu) ----
#define LEV_GROUP_NCLS 6  // column classes C = 8, 12, ..., 28 per lane
// lanes per pair for a padded reference length R: smallest power of two >= 2 with
// G * 28 >= R + 1; 0 if the group kernel does not apply
static inline int lev_group_lanes(int64_t R) {
    int G = 2;
    while (G <= 32 && (int64_t)G * 28 < R + 1) G <<= 1;
    return G <= 32 ? G : 0;
}
// smallest class whose strip of G*C columns covers columns 0..r
__host__ __device__ static inline int lev_group_class(int r, int G) {
    const int need = (r + G) / G;  // ceil((r + 1) / G)
    const int cls = (need - 8 + 3) >> 2;
    return cls < 0 ? 0 : cls;
}
// histogram bin of a pair: classes descending, then hypothesis lengths descending
__host__ __device__ static inline int lev_group_bin(int r, int h, int G, int H) {
    return (LEV_GROUP_NCLS - 1 - lev_group_class(r, G)) * (H + 1) + (H - h);
}

// Workspace layout (device scratch supplied by the caller).  All offsets are
// multiples of 128 bytes.
struct LevLayout {
    int64_t R, H, Nref, P;
    int64_t Rp, Hp;  // padded row lengths of the packed int32 token tables
    int64_t Hp16;    // padded row length of the 16-bit hypothesis table (multiple of 8)
    int64_t Hout;    // output rows of the prefix / mask modes
    int64_t Wd;      // 32-bit words of one (prefix, pair) distinct-token bitmap
    size_t off_ref_tok, off_hyp_tok, off_hyp_tok16, off_ref_len, off_hyp_len, off_flags;
    // group kernel (lev_group.cu): [ghist nbins][gcursor nbins][gmeta 16] follow the 4 state
    // words directly (one memset clears state + ghist); slots = sorted task table
    int64_t nbins;
    size_t off_ghist, off_gcursor, off_gmeta, off_slots;
    // raw prefix rows written by the group kernel, read by lev_prefix_finalize_kernel
    int64_t Hr, Hr16;
    size_t off_raw;
    size_t off_uid, off_dtok, off_ndist, off_dbits;
    // bit-vector path (lev_bitvec.cu): 16-byte uid chunks [ceil(R/16)][P]; the hypothesis
    // chunks [ceil(H/16)][P] reuse the packed-hypothesis region, which that path leaves idle
    size_t off_bv_ref;
    // pack kernel, few long sequences: [N first-eos words][ceil(N / 32) tickets] per side
    size_t off_split_ref, off_split_hyp;
    size_t bytes;
};

// kind: 0 = final value, 1 = completion (mask mode), 2 = prefix table
static inline LevLayout lev_layout(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                                   int kind, int exclude_last) {
    const int for_completion = (kind == 1);
    LevLayout L;
    L.R = ref->T;
    L.H = hyp->T;
    L.Nref = ref->N;
    L.P = hyp->N;
    L.Rp = lev_round_up(L.R > 0 ? L.R : 1, 4);
    L.Hp = lev_round_up(L.H > 0 ? L.H : 1, 4);
    L.Hp16 = lev_round_up(L.H > 0 ? L.H : 1, 8);
    L.Hout = L.H + (exclude_last ? 0 : 1);
    if (for_completion && L.Hout < 1) L.Hout = 1;  // SM:271-278
    L.Wd = (L.R + 31) / 32;
    if (L.Wd < 1) L.Wd = 1;
    size_t o = 0;
    auto take = [&](size_t n) {
        size_t at = o;
        o += (size_t)lev_round_up((int64_t)n, 128);
        return at;
    };
    L.off_ref_tok = take(sizeof(int32_t) * (size_t)L.Nref * L.Rp);
    L.off_hyp_tok = take(sizeof(int32_t) * (size_t)L.P * L.Hp);
    L.off_hyp_tok16 = take(sizeof(uint16_t) * (size_t)L.P * L.Hp16);
    L.off_ref_len = take(sizeof(int32_t) * (size_t)L.Nref);
    L.off_hyp_len = take(sizeof(int32_t) * (size_t)L.P);
    L.nbins = LEV_GROUP_NCLS * (L.H + 1);
    L.off_flags = take(sizeof(int32_t) * (size_t)(4 + 2 * L.nbins + 16));  // state + group tables
    L.off_ghist = L.off_flags + 4 * sizeof(int32_t);
    L.off_gcursor = L.off_ghist + sizeof(int32_t) * (size_t)L.nbins;
    L.off_gmeta = L.off_gcursor + sizeof(int32_t) * (size_t)L.nbins;
    L.off_slots = take(16 * (size_t)(L.P + LEV_GROUP_NCLS * 32));
    L.Hr = lev_round_up(L.H + 1, 4);
    L.Hr16 = lev_round_up(L.H + 1, 8);
    L.off_raw = (kind == 2) ? take(sizeof(int32_t) * (size_t)L.P * L.Hr) : 0;
    L.off_bv_ref = (kind != 1 && L.R <= 128) ? take((size_t)L.P * (size_t)lev_round_up(L.R > 0 ? L.R : 1, 16)) : 0;
    L.off_uid = L.off_dtok = L.off_ndist = L.off_dbits = 0;
    if (for_completion) {
        L.off_uid = take(sizeof(int32_t) * (size_t)L.Nref * L.Rp);
        L.off_dtok = take(sizeof(int64_t) * (size_t)L.Nref * L.Rp);
        L.off_ndist = take(sizeof(int32_t) * (size_t)L.Nref);
        L.off_dbits = take(sizeof(uint32_t) * (size_t)L.Hout * L.P * L.Wd);
    }
    L.off_split_ref = take(sizeof(int32_t) * (size_t)(L.Nref + (L.Nref + 31) / 32));
    L.off_split_hyp = take(sizeof(int32_t) * (size_t)(L.P + (L.P + 31) / 32));
    L.bytes = o;
    return L;
}

// Everything the DP kernels need, by value.
struct LevParams {
    const int32_t* ref_tok;  // [Nref][Rp]
    const int32_t* hyp_tok;  // [P][Hp]
    const uint16_t* hyp_tok16;  // [P][Hp16] low 16 bits of every hypothesis token
    int64_t Hp16;
    const int32_t* ref_len;  // [Nref]
    const int32_t* hyp_len;  // [P]
    int64_t Rp, Hp;
    int R, H, P, ref_group;
    // costs: integer path uses the *_i fields, float path the *_f fields
    int ins_i, del_i, sub_i;
    float ins_f, del_f, sub_f;
    float mult;  // SM:168-174
    int norm, exclude_last;
    float padding;
    // FINAL: out[n];  PREFIX: out[i*out_si + n*out_sn]
    float* out;
    int64_t out_si, out_sn;
    int Hout;
    // MASK
    const int32_t* uid;  // [Nref][Rp] rank of ref[j] among the pair's distinct tokens
    const int32_t* ndist;  // [Nref] number of distinct tokens of the reference
    uint32_t* dbits;     // [Hout][P][Wd]
    int Wd;
    int* umax;
    int* flags;            // caller's warning flags (may be NULL)
    const int* wide_flag;  // never NULL: workspace state words K0 fills: [0] flags (incl.
                           // B200LEV_FLAG_WIDE_TOKENS), [1], [2] biased token range
    // the caller's tensors, for the 64-bit compare path only
    const void* ref_raw;
    const void* hyp_raw;
    int64_t ref_st, ref_sn, hyp_st, hyp_sn;
    int ref_eb, hyp_eb;
    int only_if_wide;  // lev_warp_kernel: exit unless the wide-token flag is set
    int mask16;        // MASK: lev_mask16_kernel ran first; lev_warp_kernel skips the pairs it took
    int bv_check;      // the bit-vector kernels ran first: exit if they took the batch
    // group kernel tables (workspace)
    int* ghist;
    int* gcursor;
    int* gmeta;   // [0] = number of tasks
    int4* slots;  // sorted (pair, r, h, class) per task slot; pair < 0 = padding
    int* gmeta_order;  // same storage viewed as int[P]: pair order of the CTA kernel
    int32_t* raw32;          // [P][Hr]   raw prefix rows (32-bit group path)
    unsigned short* raw16;   // [P][Hr16] same storage, packed path
    int64_t Hr, Hr16;
    int nbins;
};


#if defined(__CUDACC__) || defined(B200LEV_EMU)
// K0's token range (state[1] = max(u), state[2] = max(~u), u = token + 2^31): all tokens of the
// call lie in one 65 536-wide window, so their low 16 bits tell them apart
__device__ __forceinline__ bool lev_tokens_narrow(const int* state) {
    const unsigned umax = (unsigned)state[1], umin = ~(unsigned)state[2];
    return umax < umin || umax - umin < 65536u;
}
// Mask mode, small alphabets (character-level targets): when every token of the call lies in a
// window of 32 values, "token - smallest token" IS an order-preserving index into a one-word
// bitmap -- no sort per reference, no rank -> token table (lev_completion.cu)
__device__ __forceinline__ bool lev_tokens_direct(const int* state, int* tmin) {
    const unsigned umax = (unsigned)state[1], umin = ~(unsigned)state[2];
    *tmin = (int)(umin ^ 0x80000000u);
    return !(state[0] & B200LEV_FLAG_WIDE_TOKENS) && umax >= umin && umax - umin < 32u;
}
// the packed mask kernel (lev_mask16.cu) and lev_warp_kernel split a mask-mode batch by these
// two rules: the first is per call, the second per pair
__device__ __forceinline__ bool lev_mask16_tokens_ok(const int* state) {
    return !(state[0] & B200LEV_FLAG_WIDE_TOKENS) && lev_tokens_narrow(state);
}
__device__ __forceinline__ bool lev_mask16_takes(const LevParams& p, int pair) {
    return p.ndist[pair / p.ref_group] <= 32;
}
#endif

// optional per-kernel timing (b200lev_profile): CUDA events recorded on the launch stream
// around each phase; slots of b200lev_profile_read()
enum LevProfSlot { LEV_PROF_PACK_REF = 0, LEV_PROF_PACK_HYP, LEV_PROF_SORT, LEV_PROF_DP,
                   LEV_PROF_FINALIZE, LEV_PROF_STANDBY, LEV_PROF_BV_UID, LEV_PROF_BV_DP,
                   LEV_PROF_COMP_UID, LEV_PROF_COMP_FILL, LEV_PROF_ERR_SUM, LEV_PROF_MASK16,
                   LEV_PROF_NSLOTS };
void lev_prof_begin(int slot, cudaStream_t st);
void lev_prof_end(int slot, cudaStream_t st);

// host-side status plumbing (lev_abi.cu)
void lev_set_error(const char* fmt, ...);
int lev_check_cuda(const char* what);

// kernels' host launchers
int lev_launch_pack(const b200lev_tokens_t* t, int has_eos, int64_t eos, int include_eos,
                    int32_t* packed, int64_t Tp, uint16_t* packed16, int64_t Tp16, int32_t* lens,
                    int32_t* flags, int32_t* state, int missing_flag, const int32_t* ref_len,
                    int ref_group, int G, int* ghist, int bv_check, cudaStream_t st,
                    int32_t* split_scratch = nullptr);  // [N + ceil(N / 32)] words, or NULL
int lev_launch_dp(const LevParams& p, int mode, bool count_mode, bool float_path,
                  cudaStream_t st);
int lev_launch_group(const LevParams& p, int mode, bool count_mode, cudaStream_t st);
int lev_launch_mask16(const LevParams& p, bool float_path, cudaStream_t st);
int lev_launch_cta(const LevParams& p, int mode, bool count_mode, bool float_path, cudaStream_t st);
// lanes per pair if the shapes admit the group kernel (its histogram is then built at
// pack time), else 0
int lev_group_eligible(int64_t R, int64_t H, int64_t P);
// Device-side choice between the bit-vector path and the wavefront path (no host round trip).
// When the shapes and costs admit the bit-vector kernels they are enqueued FIRST: the uid
// kernel finds out whether every block of 32 consecutive pairs holds at most 4 runs of
// identical references (n-best batches do) and vetoes through state[3] otherwise; every
// kernel of the wavefront path, enqueued behind, then starts with lev_bv_took() and exits at
// once if the work is already done (and the bit-vector DP kernel with the opposite test).
__device__ __forceinline__ bool lev_bv_took(const int* state) {
    return *reinterpret_cast<const volatile int*>(state + 3) == 0;
}
// unit-cost bit-vector path (lev_bitvec.cu)
bool lev_bitvec_eligible(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp, int mode,
                         bool count_mode, bool float_path, int ins_i, int del_i, int sub_i,
                         int64_t out_sn, int ref_group, bool* short_form, bool* grouped);
int lev_bitvec_launch(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                      const b200lev_opts_t* o, int mode, float mult, int32_t* ref_len,
                      int32_t* hyp_len, void* uid_ref, void* uid_hyp, void* lead,
                      int32_t* state, int32_t* flags, float* out, int64_t out_si, int Hout,
                      cudaStream_t st, void* after_uid, bool short_form, double* acc);
// 0: off, 1: forced (tests; runs whatever the references look like), 2: device-selected
int lev_bitvec_mode();
int lev_launch_uid(const b200lev_tokens_t* ref, const int32_t* packed, const int* state,
                   const int32_t* ref_len, int32_t* uid,
                   int64_t* dtok, int32_t* ndist, int64_t Rp, cudaStream_t st);
int lev_launch_completion_fill(const uint32_t* dbits, const int64_t* dtok, const int* state, int64_t Rp,
                               int64_t Hout, int64_t P, int64_t Wd, int ref_group, int64_t U,
                               int64_t padding, int64_t* out, int64_t out_si, int64_t out_sn,
                               cudaStream_t st);
