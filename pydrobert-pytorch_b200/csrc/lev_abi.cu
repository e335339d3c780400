// lev_abi.cu -- the extern "C" boundary declared in include/b200lev.h: argument
// checking, the uniform-cost shortcut (SM:168-174), int/float path selection, workspace
// carving and kernel sequencing.  No device allocation, no host synchronisation.
#include <mutex>
#include <cmath>
#include <cstdarg>
#include <cstdio>

#include "lev_common.cuh"

static thread_local char g_err[512] = "";

void lev_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int lev_check_cuda(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        lev_set_error("%s: %s", what, cudaGetErrorString(e));
        return B200LEV_ERR_CUDA;
    }
    return B200LEV_OK;
}

// ---- optional per-kernel timing ----------------------------------------------------------
#ifndef B200LEV_EMU
static bool g_prof_on = false;
static cudaEvent_t g_prof_ev[LEV_PROF_NSLOTS][2];
static bool g_prof_used[LEV_PROF_NSLOTS];
static bool g_prof_init = false;
void lev_prof_begin(int slot, cudaStream_t st) {
    if (g_prof_on) cudaEventRecord(g_prof_ev[slot][0], st);
}
void lev_prof_end(int slot, cudaStream_t st) {
    if (g_prof_on) {
        cudaEventRecord(g_prof_ev[slot][1], st);
        g_prof_used[slot] = true;
    }
}
extern "C" int b200lev_profile(int enable) {
    if (enable && !g_prof_init) {
        for (int s = 0; s < LEV_PROF_NSLOTS; ++s)
            for (int k = 0; k < 2; ++k)
                if (cudaEventCreate(&g_prof_ev[s][k]) != cudaSuccess) return lev_check_cuda("cudaEventCreate");
        g_prof_init = true;
    }
    for (int s = 0; s < LEV_PROF_NSLOTS; ++s) g_prof_used[s] = false;
    g_prof_on = enable != 0;
    return B200LEV_OK;
}
extern "C" int b200lev_profile_read(float* ms, int n) {
    for (int s = 0; s < n && s < LEV_PROF_NSLOTS; ++s) {
        ms[s] = -1.0f;
        if (g_prof_init && g_prof_used[s]) {
            if (cudaEventSynchronize(g_prof_ev[s][1]) != cudaSuccess) return lev_check_cuda("cudaEventSynchronize");
            cudaEventElapsedTime(&ms[s], g_prof_ev[s][0], g_prof_ev[s][1]);
        }
    }
    return B200LEV_OK;
}
#else
void lev_prof_begin(int, cudaStream_t) {}
void lev_prof_end(int, cudaStream_t) {}
extern "C" int b200lev_profile(int) { return B200LEV_OK; }
extern "C" int b200lev_profile_read(float* ms, int n) {
    for (int s = 0; s < n; ++s) ms[s] = -1.0f;
    return B200LEV_OK;
}
#endif

extern "C" int b200lev_abi_version(void) { return B200LEV_ABI_VERSION; }
extern "C" const char* b200lev_last_error(void) { return g_err; }

extern "C" int b200lev_copy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch,
                                    size_t width_bytes, size_t height, int32_t to_device,
                                    void* stream) {
    if (width_bytes == 0 || height == 0) return B200LEV_OK;
    if (dst == nullptr || src == nullptr || dst_pitch < width_bytes || src_pitch < width_bytes) {
        lev_set_error("b200lev_copy2d_async: bad pointers or pitches");
        return B200LEV_ERR_ARG;
    }
#ifdef B200LEV_EMU
    (void)to_device;
    (void)stream;
    for (size_t r = 0; r < height; ++r)
        memcpy((char*)dst + r * dst_pitch, (const char*)src + r * src_pitch, width_bytes);
    return B200LEV_OK;
#else
    cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, height,
                      to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                      (cudaStream_t)stream);
    return lev_check_cuda("cudaMemcpy2DAsync");
#endif
}

extern "C" int b200lev_device_count(void) {
#ifdef B200LEV_EMU
    return 1;
#else
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
#endif
}

static int lev_check_tokens(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                            const b200lev_opts_t* o) {
    if (!ref || !hyp || !o) {
        lev_set_error("NULL argument");
        return B200LEV_ERR_ARG;
    }
    if (ref->T < 0 || hyp->T < 0 || ref->N < 0 || hyp->N < 0) {
        lev_set_error("negative dimension");
        return B200LEV_ERR_ARG;
    }
    if (o->ref_group < 1 || ref->N * o->ref_group != hyp->N) {
        // SM:191-194
        lev_set_error("ref has batch size %lld, but hyp has %lld",
                      (long long)(ref->N * (o->ref_group < 1 ? 1 : o->ref_group)), (long long)hyp->N);
        return B200LEV_ERR_ARG;
    }
    if ((ref->N * ref->T > 0 && !ref->data) || (hyp->N * hyp->T > 0 && !hyp->data)) {
        lev_set_error("NULL token data");
        return B200LEV_ERR_ARG;
    }
    if (ref->T >= (1 << 24) || hyp->T >= (1 << 24) || hyp->N >= ((int64_t)1 << 31)) {
        lev_set_error("dimension too large");
        return B200LEV_ERR_UNSUPPORTED;
    }
    return B200LEV_OK;
}

extern "C" size_t b200lev_workspace_bytes(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                                          int32_t kind, int32_t exclude_last) {
    if (!ref || !hyp) return 0;
    return lev_layout(ref, hyp, kind, exclude_last).bytes + 128;
}

static inline char* lev_ws_base(const void* ws) {
    return (char*)(((uintptr_t)ws + 127) & ~(uintptr_t)127);
}

extern "C" const int32_t* b200lev_workspace_ref_lens(const b200lev_tokens_t* ref,
                                                     const b200lev_tokens_t* hyp, const void* ws) {
    LevLayout L = lev_layout(ref, hyp, 0, 0);
    return (const int32_t*)(lev_ws_base(ws) + L.off_ref_len);
}
extern "C" const int32_t* b200lev_workspace_hyp_lens(const b200lev_tokens_t* ref,
                                                     const b200lev_tokens_t* hyp, const void* ws) {
    LevLayout L = lev_layout(ref, hyp, 0, 0);
    return (const int32_t*)(lev_ws_base(ws) + L.off_hyp_len);
}

// Fills the cost fields of `p`.  Returns count_mode / float_path through the out-params.
static void lev_classify_costs(const b200lev_opts_t* o, int64_t R, int64_t H, LevParams* p,
                               bool* count_mode, bool* float_path) {
    float ins = o->ins_cost, del = o->del_cost, sub = o->sub_cost;
    float mult = 1.0f;
    bool cm = o->return_mistakes != 0;
    if (ins == del && del == sub && sub > 0.0f) {  // SM:168-174
        if (!cm) mult = ins;
        ins = del = sub = 1.0f;
        cm = false;
    }
    // integer path iff every cost is an integer and every reachable fp32 value of the
    // reference is exact (< 2^24), so int32 arithmetic reproduces it bit for bit
    const float amax = fmaxf(fabsf(ins), fmaxf(fabsf(del), fabsf(sub)));
    const bool integral = floorf(ins) == ins && floorf(del) == del && floorf(sub) == sub &&
                          (double)amax * (double)(R + H + 2) < 16777216.0;
    p->ins_f = ins;
    p->del_f = del;
    p->sub_f = sub;
    p->ins_i = integral ? (int)ins : 0;
    p->del_i = integral ? (int)del : 0;
    p->sub_i = integral ? (int)sub : 0;
    p->mult = mult;
    *count_mode = cm;
    *float_path = !integral;
}

// pack both token tensors into the workspace and fill the common fields of `p`
static int lev_prepare(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                       const b200lev_opts_t* o, const LevLayout& L, char* ws, int32_t* flags,
                       cudaStream_t st, LevParams* p, bool do_pack = true, int bv_check = 0) {
    int32_t* ref_tok = (int32_t*)(ws + L.off_ref_tok);
    int32_t* hyp_tok = (int32_t*)(ws + L.off_hyp_tok);
    int32_t* ref_len = (int32_t*)(ws + L.off_ref_len);
    int32_t* hyp_len = (int32_t*)(ws + L.off_hyp_len);
    // K0 leaves what the DP kernels need to know about the tokens (wider than int32?
    // value range) in 4 state words of the workspace; the caller's flags get the same
    // warning bits
    int32_t* state = (int32_t*)(ws + L.off_flags);
    if (do_pack) {
        // the group kernel's (class, length) histogram is built by the hypothesis pass
        // whenever the shapes make that kernel eligible (lev_group.cu decides later)
        const int G = lev_group_eligible(L.R, L.H, L.P);
        const size_t clear = sizeof(int32_t) * (size_t)(4 + (G ? L.nbins : 0));
        // (with bv_check the bit-vector launcher cleared them before its kernels ran)
        if (!bv_check && cudaMemsetAsync(state, 0, clear, st) != cudaSuccess) return lev_check_cuda("memset");
        lev_prof_begin(LEV_PROF_PACK_REF, st);
        int rc = lev_launch_pack(ref, o->has_eos, o->eos, o->include_eos, ref_tok, L.Rp, nullptr, 0,
                                 ref_len, flags, state, B200LEV_FLAG_REF_NO_EOS, nullptr, 1, 0,
                                 nullptr, bv_check, st, (int32_t*)(ws + L.off_split_ref));
        lev_prof_end(LEV_PROF_PACK_REF, st);
        if (rc) return rc;
        lev_prof_begin(LEV_PROF_PACK_HYP, st);
        rc = lev_launch_pack(hyp, o->has_eos, o->eos, o->include_eos, hyp_tok, L.Hp,
                             (uint16_t*)(ws + L.off_hyp_tok16), L.Hp16, hyp_len, flags, state,
                             B200LEV_FLAG_HYP_NO_EOS, ref_len, o->ref_group, G,
                             G ? (int*)(ws + L.off_ghist) : nullptr, bv_check, st,
                             (int32_t*)(ws + L.off_split_hyp));
        lev_prof_end(LEV_PROF_PACK_HYP, st);
        if (rc) return rc;
    }
    memset(p, 0, sizeof(*p));
    p->ref_tok = ref_tok;
    p->hyp_tok = hyp_tok;
    p->hyp_tok16 = (const uint16_t*)(ws + L.off_hyp_tok16);
    p->Hp16 = L.Hp16;
    p->ref_len = ref_len;
    p->hyp_len = hyp_len;
    p->Rp = L.Rp;
    p->Hp = L.Hp;
    p->R = (int)L.R;
    p->H = (int)L.H;
    p->P = (int)L.P;
    p->ref_group = o->ref_group;
    p->norm = o->norm;
    p->exclude_last = o->exclude_last;
    p->padding = (float)o->padding;
    p->flags = flags;
    p->wide_flag = state;
    p->bv_check = bv_check;
    p->ref_raw = ref->data;
    p->hyp_raw = hyp->data;
    p->ref_st = ref->stride_t;
    p->ref_sn = ref->stride_n;
    p->hyp_st = hyp->stride_t;
    p->hyp_sn = hyp->stride_n;
    p->ref_eb = ref->elem_bytes;
    p->hyp_eb = hyp->elem_bytes;
    p->ghist = (int*)(ws + L.off_ghist);
    p->gcursor = (int*)(ws + L.off_gcursor);
    p->gmeta = (int*)(ws + L.off_gmeta);
    p->slots = (int4*)(ws + L.off_slots);
    p->gmeta_order = (int*)(ws + L.off_slots);
    p->nbins = (int)L.nbins;
    p->raw32 = L.off_raw ? (int32_t*)(ws + L.off_raw) : nullptr;
    p->raw16 = (unsigned short*)p->raw32;
    p->Hr = L.Hr;
    p->Hr16 = L.Hr16;
    return B200LEV_OK;
}

// Unit costs, short references, sequence-first tensors: the bit-vector path (lev_bitvec.cu)
// takes the whole call -- lengths, warnings and DP -- straight from the raw tokens.
// Returns 0 if it does not apply, 1 if it ran unconditionally (forced mode: the call is
// Device-selected mode forks: once the uid pre-pass has finished the veto is final, so the
// wavefront path's stand-by chain (seven launches that exit at once when the bit-vector
// kernels took the call) runs on a side stream NEXT TO the bit-vector DP kernel instead of
// behind it, and the caller's stream joins at the end.  In the vetoed case the DP kernel
// exits at once and the chain does the work.  Nothing the chain touches before its kernels
// have checked the veto overlaps what the DP kernel reads (lev_group.cu clears its tables in
// a kernel that stands by too).  One side stream and two events per device, made once.
// (the side stream and its events are shared by every caller on the device: calls that fork
// are enqueued one at a time)
static std::mutex g_fork_mu;
struct LevFork {
    cudaStream_t side;
    void* after_uid;
    void* joined;
};
#ifndef B200LEV_EMU
static bool lev_fork_get(LevFork* f) {
    static LevFork cache[64];
    static bool have[64];
    if (const char* e = getenv("B200LEV_FORK"))
        if (atoi(e) == 0) return false;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    if (!have[dev]) {
        cudaEvent_t a, b;
        cudaStream_t s2;
        if (cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&a, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&b, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        cache[dev].side = s2;
        cache[dev].after_uid = (void*)a;
        cache[dev].joined = (void*)b;
        have[dev] = true;
    }
    *f = cache[dev];
    return true;
}
// the chain goes to the side stream once the uid kernel is done; returns the stream to use
static cudaStream_t lev_fork_begin(const LevFork& f, bool on) {
    if (!on) return nullptr;
    if (cudaStreamWaitEvent(f.side, (cudaEvent_t)f.after_uid, 0) != cudaSuccess) return nullptr;
    return f.side;
}
static int lev_fork_join(const LevFork& f, cudaStream_t st) {
    if (cudaEventRecord((cudaEvent_t)f.joined, f.side) != cudaSuccess ||
        cudaStreamWaitEvent(st, (cudaEvent_t)f.joined, 0) != cudaSuccess)
        return lev_check_cuda("fork join");
    return B200LEV_OK;
}
#else
static bool lev_fork_get(LevFork*) { return false; }
static cudaStream_t lev_fork_begin(const LevFork&, bool) { return nullptr; }
static int lev_fork_join(const LevFork&, cudaStream_t) { return B200LEV_OK; }
#endif

// done), 2 if its kernels were enqueued in device-selected mode (the caller goes on to
// enqueue the wavefront path with bv_check set; the state words are already cleared),
// < 0 on error.
static int lev_try_bitvec(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                          const b200lev_opts_t* o, const LevLayout& L, char* ws, int32_t* flags,
                          cudaStream_t st, int mode, float* out, int64_t out_si, int64_t out_sn,
                          int Hout, LevFork* fork, bool* forked, double* acc = nullptr,
                          bool* sums_done = nullptr) {
    *forked = false;
    if (!L.off_bv_ref) return 0;
    LevParams tmp;
    memset(&tmp, 0, sizeof(tmp));
    bool cm, fp;
    lev_classify_costs(o, L.R, L.H, &tmp, &cm, &fp);
    bool short_form = false, grouped = false;
    if (!lev_bitvec_eligible(ref, hyp, mode, cm, fp, tmp.ins_i, tmp.del_i, tmp.sub_i, out_sn, o->ref_group,
                             &short_form, &grouped))
        return 0;
    // the short-reference kernel takes any batch, the fused kernel any batch whose references the
    // caller declared shared: nothing to select on the device, no chain behind them
    const bool forced = lev_bitvec_mode() == 1 || short_form || grouped;
    int32_t* state = (int32_t*)(ws + L.off_flags);
    if (!forced) {
        // device-selected: only where the wavefront fallback is the group path, whose kernels
        // know how to stand by; the state words are cleared here, once, for both paths
        const int G = lev_group_eligible(L.R, L.H, L.P);
        if (G == 0) return 0;
        const size_t clear = sizeof(int32_t) * (size_t)(4 + L.nbins);
        if (cudaMemsetAsync(state, 0, clear, st) != cudaSuccess) return lev_check_cuda("memset");
    }
    *forked = !forced && lev_fork_get(fork);
    const int rc = lev_bitvec_launch(ref, hyp, o, mode, tmp.mult, (int32_t*)(ws + L.off_ref_len),
                                     (int32_t*)(ws + L.off_hyp_len), ws + L.off_bv_ref,
                                     ws + L.off_hyp_tok, ws + L.off_slots, forced ? nullptr : state,
                                     flags, out, out_si, Hout, st, *forked ? fork->after_uid : nullptr,
                                     short_form, acc);
    if (rc) return rc;
    if (sums_done != nullptr)
        *sums_done = short_form && mode == LEV_MODE_FINAL && acc != nullptr && !(getenv("B200LEV_BVS_SUMS") && atoi(getenv("B200LEV_BVS_SUMS")) == 0);
    return forced ? 1 : 2;
}

extern "C" int b200lev_err_sum(const float* er, const int32_t* ref_lens, int64_t P, int32_t ref_group,
                               double* acc, void* stream);

static int lev_final_impl(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                          const b200lev_opts_t* opts, float* out, void* workspace,
                          size_t workspace_bytes, int32_t* flags, void* stream, bool do_pack,
                          double* acc = nullptr) {
    int rc = lev_check_tokens(ref, hyp, opts);
    if (rc) return rc;
    if (hyp->N == 0) return B200LEV_OK;
    if (!out || !workspace) {
        lev_set_error("NULL output/workspace");
        return B200LEV_ERR_ARG;
    }
    const LevLayout L = lev_layout(ref, hyp, 0, 0);
    if (workspace_bytes < L.bytes + 128) {
        lev_set_error("workspace too small: %zu < %zu", workspace_bytes, L.bytes + 128);
        return B200LEV_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    LevParams p;
    b200lev_opts_t o = *opts;
    o.exclude_last = 0;  // SM:165
    int bv = 0;
    LevFork fork;
    bool forked = false;
    std::lock_guard<std::mutex> fork_lock(g_fork_mu);
    if (do_pack) {
        bool sums_done = false;
        bv = lev_try_bitvec(ref, hyp, &o, L, lev_ws_base(workspace), flags, st, LEV_MODE_FINAL, out,
                            0, 1, 0, &fork, &forked, acc, &sums_done);
        if (bv < 0) return bv;
        if (bv == 1) {
            if (acc != nullptr && !sums_done)
                return b200lev_err_sum(out, (const int32_t*)(lev_ws_base(workspace) + L.off_ref_len), hyp->N,
                                       o.ref_group, acc, stream);
            return B200LEV_OK;
        }
    }
    cudaStream_t side = lev_fork_begin(fork, forked && bv == 2);
    cudaStream_t cs = side ? side : st;  // the wavefront chain's stream
    rc = lev_prepare(ref, hyp, &o, L, lev_ws_base(workspace), flags, cs, &p, do_pack, bv == 2);
    if (rc == B200LEV_OK) {
        bool cm, fp;
        lev_classify_costs(&o, L.R, L.H, &p, &cm, &fp);
        p.out = out;
        rc = lev_launch_dp(p, LEV_MODE_FINAL, cm, fp, cs);
    }
    if (side) {
        const int jr = lev_fork_join(fork, st);
        if (rc == B200LEV_OK) rc = jr;
    }
    if (rc == B200LEV_OK && acc != nullptr)
        rc = b200lev_err_sum(out, (const int32_t*)(lev_ws_base(workspace) + L.off_ref_len), hyp->N,
                             o.ref_group, acc, stream);
    return rc;
}

extern "C" int b200lev_final_sums(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                                  const b200lev_opts_t* opts, float* out, void* workspace,
                                  size_t workspace_bytes, int32_t* flags, double* acc, void* stream) {
    if (!acc) {
        lev_set_error("NULL accumulator");
        return B200LEV_ERR_ARG;
    }
    if (hyp && hyp->N == 0) return B200LEV_OK;
    return lev_final_impl(ref, hyp, opts, out, workspace, workspace_bytes, flags, stream, true, acc);
}

extern "C" int b200lev_final(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                             const b200lev_opts_t* opts, float* out, void* workspace,
                             size_t workspace_bytes, int32_t* flags, void* stream) {
    return lev_final_impl(ref, hyp, opts, out, workspace, workspace_bytes, flags, stream, true);
}

extern "C" int b200lev_final_packed(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                                    const b200lev_opts_t* opts, float* out, void* workspace,
                                    size_t workspace_bytes, int32_t* flags, void* stream) {
    return lev_final_impl(ref, hyp, opts, out, workspace, workspace_bytes, flags, stream, false);
}

extern "C" int b200lev_pack(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                            const b200lev_opts_t* opts, void* workspace, size_t workspace_bytes,
                            int32_t* flags, void* stream) {
    int rc = lev_check_tokens(ref, hyp, opts);
    if (rc) return rc;
    if (hyp->N == 0) return B200LEV_OK;
    const LevLayout L = lev_layout(ref, hyp, 0, 0);
    if (!workspace || workspace_bytes < L.bytes + 128) {
        lev_set_error("workspace too small: %zu < %zu", workspace_bytes, L.bytes + 128);
        return B200LEV_ERR_WORKSPACE;
    }
    LevParams p;
    return lev_prepare(ref, hyp, opts, L, lev_ws_base(workspace), flags, (cudaStream_t)stream, &p, true);
}

static int lev_prefix_impl(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                           const b200lev_opts_t* opts, float* out, int64_t out_stride_i,
                           int64_t out_stride_n, void* workspace, size_t workspace_bytes,
                           int32_t* flags, void* stream, bool do_pack) {
    int rc = lev_check_tokens(ref, hyp, opts);
    if (rc) return rc;
    const LevLayout L = lev_layout(ref, hyp, 2, opts->exclude_last);
    if (hyp->N == 0 || L.Hout <= 0) return B200LEV_OK;
    if (!out || !workspace) {
        lev_set_error("NULL output/workspace");
        return B200LEV_ERR_ARG;
    }
    if (workspace_bytes < L.bytes + 128) {
        lev_set_error("workspace too small: %zu < %zu", workspace_bytes, L.bytes + 128);
        return B200LEV_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    LevParams p;
    int bv = 0;
    LevFork fork;
    bool forked = false;
    std::lock_guard<std::mutex> fork_lock(g_fork_mu);
    if (do_pack) {
        bv = lev_try_bitvec(ref, hyp, opts, L, lev_ws_base(workspace), flags, st, LEV_MODE_PREFIX,
                            out, out_stride_i, out_stride_n, (int)L.Hout, &fork, &forked);
        if (bv < 0) return bv;
        if (bv == 1) return B200LEV_OK;
    }
    cudaStream_t side = lev_fork_begin(fork, forked && bv == 2);
    cudaStream_t cs = side ? side : st;  // the wavefront chain's stream
    rc = lev_prepare(ref, hyp, opts, L, lev_ws_base(workspace), flags, cs, &p, do_pack, bv == 2);
    if (rc == B200LEV_OK) {
        bool cm, fp;
        lev_classify_costs(opts, L.R, L.H, &p, &cm, &fp);
        p.out = out;
        p.out_si = out_stride_i;
        p.out_sn = out_stride_n;
        p.Hout = (int)L.Hout;
        rc = lev_launch_dp(p, LEV_MODE_PREFIX, cm, fp, cs);
    }
    if (side) {
        const int jr = lev_fork_join(fork, st);
        if (rc == B200LEV_OK) rc = jr;
    }
    return rc;
}

extern "C" int b200lev_prefix(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                              const b200lev_opts_t* opts, float* out, int64_t out_stride_i,
                              int64_t out_stride_n, void* workspace, size_t workspace_bytes,
                              int32_t* flags, void* stream) {
    return lev_prefix_impl(ref, hyp, opts, out, out_stride_i, out_stride_n, workspace,
                           workspace_bytes, flags, stream, true);
}

extern "C" int b200lev_prefix_packed(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                                     const b200lev_opts_t* opts, float* out, int64_t out_stride_i,
                                     int64_t out_stride_n, void* workspace, size_t workspace_bytes,
                                     int32_t* flags, void* stream) {
    return lev_prefix_impl(ref, hyp, opts, out, out_stride_i, out_stride_n, workspace,
                           workspace_bytes, flags, stream, false);
}

extern "C" int b200lev_completion_count(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                                        const b200lev_opts_t* opts, void* workspace,
                                        size_t workspace_bytes, int32_t* umax, int32_t* flags,
                                        void* stream) {
    int rc = lev_check_tokens(ref, hyp, opts);
    if (rc) return rc;
    if (!umax) {
        lev_set_error("NULL umax");
        return B200LEV_ERR_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(umax, 0, sizeof(int32_t), st) != cudaSuccess) return lev_check_cuda("memset");
    if (hyp->N == 0) return B200LEV_OK;
    if (!workspace) {
        lev_set_error("NULL workspace");
        return B200LEV_ERR_ARG;
    }
    const LevLayout L = lev_layout(ref, hyp, 1, opts->exclude_last);
    if (workspace_bytes < L.bytes + 128) {
        lev_set_error("workspace too small: %zu < %zu", workspace_bytes, L.bytes + 128);
        return B200LEV_ERR_WORKSPACE;
    }
    char* ws = lev_ws_base(workspace);
    LevParams p;
    b200lev_opts_t o = *opts;
    o.return_mistakes = 0;  // SM:479-491: the mask is always taken on the cost row
    o.norm = 0;
    rc = lev_prepare(ref, hyp, &o, L, ws, flags, st, &p);
    if (rc) return rc;
    bool cm, fp;
    lev_classify_costs(&o, L.R, L.H, &p, &cm, &fp);
    int32_t* uid = (int32_t*)(ws + L.off_uid);
    int64_t* dtok = (int64_t*)(ws + L.off_dtok);
    int32_t* ndist = (int32_t*)(ws + L.off_ndist);
    uint32_t* dbits = (uint32_t*)(ws + L.off_dbits);
    rc = lev_launch_uid(ref, p.ref_tok, p.wide_flag, p.ref_len, uid, dtok, ndist, L.Rp, st);
    if (rc) return rc;
    if (cudaMemsetAsync(dbits, 0, sizeof(uint32_t) * (size_t)L.Hout * L.P * L.Wd, st) != cudaSuccess)
        return lev_check_cuda("memset");
    p.uid = uid;
    p.ndist = ndist;
    p.dbits = dbits;
    p.Wd = (int)L.Wd;
    p.umax = umax;
    p.Hout = (int)L.Hout;
    return lev_launch_dp(p, LEV_MODE_MASK, false, fp, st);
}

extern "C" int b200lev_completion_fill(const b200lev_tokens_t* ref, const b200lev_tokens_t* hyp,
                                       const b200lev_opts_t* opts, const void* workspace,
                                       size_t workspace_bytes, int64_t U, int64_t* out,
                                       int64_t out_stride_i, int64_t out_stride_n, void* stream) {
    int rc = lev_check_tokens(ref, hyp, opts);
    if (rc) return rc;
    if (hyp->N == 0 || U <= 0) return B200LEV_OK;
    const LevLayout L = lev_layout(ref, hyp, 1, opts->exclude_last);
    if (!workspace || !out || workspace_bytes < L.bytes + 128) {
        lev_set_error("bad workspace/output");
        return B200LEV_ERR_WORKSPACE;
    }
    const char* ws = lev_ws_base(workspace);
    return lev_launch_completion_fill((const uint32_t*)(ws + L.off_dbits),
                                      (const int64_t*)(ws + L.off_dtok), (const int*)(ws + L.off_flags), L.Rp,
                                      L.Hout, L.P, L.Wd,
                                      opts->ref_group, U, opts->padding, out, out_stride_i,
                                      out_stride_n, (cudaStream_t)stream);
}
