// C-ABI entry points of librssm_rollout.so (declared in include/rssm_rollout.h).
// Validates sizes, builds kernel arguments and the weight-gradient job lists, launches on the
// caller's stream.  No hidden state; errors are reported through a thread-local message.
#include <stdlib.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "kernels.h"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    return fail("%s: %s", what, cudaGetErrorString(e));
}

bool class_size_ok(int k) { return k == 2 || k == 4 || k == 8 || k == 16; }

int check_mrssm(const RssmMrssmDims* d) {
    if (!d) return fail("dims is NULL");
    if (d->B < 1 || d->T < 1) return fail("B and T must be >= 1 (got B=%d T=%d)", d->B, d->T);
    if (d->D != 32 || d->H != 32 || d->E != 64)
        return fail("unsupported sizes D=%d H=%d E=%d: this build instantiates deter=hidden=32, embed=64", d->D, d->H, d->E);
    if (d->C * d->K != 16 || !class_size_ok(d->K))
        return fail("unsupported distribution_config (class=%d, category=%d): need class*category = 16, class in {2,4,8,16}", d->K,
                    d->C);
    if (d->A < 2 || d->A > 8 || (d->A & 1)) return fail("unsupported action_size %d: need an even size in 2..8", d->A);
    if (d->precision != RSSM_PRECISION_FP32 && d->precision != RSSM_PRECISION_BF16) return fail("bad precision %d", d->precision);
    return 0;
}

int check_mtrssm(const RssmMtrssmDims* d) {
    if (!d) return fail("dims is NULL");
    if (d->B < 1 || d->T < 1) return fail("B and T must be >= 1 (got B=%d T=%d)", d->B, d->T);
    if (d->HD != 32 || d->LD != 32 || d->HH != 32 || d->HR != 32 || d->E != 64)
        return fail("unsupported sizes hd=%d ld=%d head=%d rep=%d E=%d: this build instantiates 32/32/32/32/64", d->HD, d->LD, d->HH,
                    d->HR, d->E);
    if (d->CL * d->KL != 16 || d->CH * d->KH != 16 || !class_size_ok(d->KL) || !class_size_ok(d->KH))
        return fail("unsupported l_dist/h_dist (class,category) = (%d,%d)/(%d,%d): need class*category = 16", d->KL, d->CL, d->KH,
                    d->CH);
    if (d->A < 2 || d->A > 8 || (d->A & 1)) return fail("unsupported action_size %d: need an even size in 2..8", d->A);
    if (!(d->l_tau > 1.f) || !(d->h_tau > 1.f)) return fail("tau must be greater than 1.0 (l_tau=%g h_tau=%g)", d->l_tau, d->h_tau);
    if (d->precision != RSSM_PRECISION_FP32 && d->precision != RSSM_PRECISION_BF16 && d->precision != RSSM_PRECISION_BF16_FUSED)
        return fail("bad precision %d", d->precision);
    return 0;
}

// kernel precision policy (fp32-parity / bf16) and saved-record row length of an MMTRSSM precision value
int mt_kernel_precision(int precision) { return precision == RSSM_PRECISION_FP32 ? RSSM_PRECISION_FP32 : RSSM_PRECISION_BF16; }
int mt_saved_ld(int) { return MTRSSM_SAVED_FLOATS; }

#define REQUIRE(ptr)                                                           \
    do {                                                                       \
        if ((ptr) == nullptr) return fail("required pointer %s is NULL", #ptr); \
    } while (0)

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// appends a source segment at staged column `dst` (segments must be added in increasing, gap-free dst order).
// `elem` = element size in bytes of the source (4: fp32, 2: bf16); ptr/ptr0 are element pointers of that type.
void add_seg(rssm::WgradMmaArgs& a, int dst, const void* ptr, int ld, int valid, int elem = 4, int shift = 0,
             const void* ptr0 = nullptr, int ld0 = 0) {
    rssm::WgradSeg& s = a.seg[a.nseg++];
    s.ptr = static_cast<const char*>(ptr), s.ptr0 = static_cast<const char*>(ptr0 ? ptr0 : ptr);
    s.ld_bytes = ld * elem, s.ld0_bytes = (ptr0 ? ld0 : ld) * elem;
    s.valid = valid, s.shift = shift;
    s.c4_begin = dst / 4, s.c4_end = s.c4_begin + (valid + 3) / 4;
    if (elem == 2) {
        s.kind = 2;
    } else {
        const bool vec = (valid % 4 == 0) && (s.ld_bytes % 16 == 0) && (s.ld0_bytes % 16 == 0) && aligned16(s.ptr) && aligned16(s.ptr0);
        s.kind = vec ? 0 : 1;
    }
    a.stride = s.c4_end * 4;
}

// element pointer arithmetic on the opaque records (fp32 or bf16 depending on the precision)
const void* rec_at(const void* base, int elem_off, int elem) { return static_cast<const char*>(base) + (size_t)elem_off * elem; }

void set_out(rssm::WgradMmaArgs& a, int id, float* dW, int ldw, int kvalid, float* db0 = nullptr, float* db1 = nullptr) {
    rssm::WgradOut& o = a.out[id];
    o.dW = dW, o.ldw = ldw, o.kvalid = kvalid, o.db0 = db0, o.db1 = db1;
}

}  // namespace

extern "C" {

int rssm_abi_version(void) { return RSSM_ABI_VERSION; }
const char* rssm_last_error(void) { return g_err; }
long long rssm_kernel_launch_count(void) { return g_launches.load(); }

// ---------------------------------------------------------------------------------------------------
// MoPoE-MRSSM
// ---------------------------------------------------------------------------------------------------
static int mrssm_fwd_common(const RssmMrssmDims* d, const RssmMrssmWeights* w, const RssmMrssmInputs* in,
                            const RssmMrssmOutputs* out, void* stream, bool imagine) {
    if (check_mrssm(d)) return 1;
    REQUIRE(w); REQUIRE(in); REQUIRE(out);
    REQUIRE(in->actions); REQUIRE(in->h0); REQUIRE(in->z0); REQUIRE(out->feature); REQUIRE(out->prior_probs);
    REQUIRE(w->asp_w1); REQUIRE(w->w_ih); REQUIRE(w->w_hh); REQUIRE(w->pr_w1); REQUIRE(w->pr_w2);
    if (imagine) {
        REQUIRE(in->u_prior);
    } else {
        REQUIRE(in->embed_a); REQUIRE(in->embed_v); REQUIRE(in->u_post); REQUIRE(out->post_probs); REQUIRE(out->kl);
        REQUIRE(w->au_w1); REQUIRE(w->au_w2); REQUIRE(w->vi_w1); REQUIRE(w->vi_w2);
    }
    rssm::MrssmFwdArgs a{};
    a.B = d->B, a.T = d->T, a.A = d->A, a.K = d->K, a.w = *w;
    a.actions = in->actions, a.embed_a = in->embed_a, a.embed_v = in->embed_v, a.h0 = in->h0, a.z0 = in->z0;
    a.u_post = in->u_post, a.u_prior = in->u_prior;
    a.feature = out->feature, a.prior_probs = out->prior_probs, a.post_probs = out->post_probs;
    a.prior_stoch = out->prior_stoch, a.kl = out->kl, a.saved = imagine ? nullptr : out->saved;
    g_launches.fetch_add(1);
    return check_cuda(rssm::launch_mrssm_fwd(a, d->precision, imagine, static_cast<cudaStream_t>(stream)),
                      imagine ? "mrssm imagine launch" : "mrssm forward launch");
}

int rssm_mrssm_rollout_fwd(const RssmMrssmDims* d, const RssmMrssmWeights* w, const RssmMrssmInputs* in, const RssmMrssmOutputs* out,
                           void* stream) {
    return mrssm_fwd_common(d, w, in, out, stream, false);
}

int rssm_mrssm_imagine_fwd(const RssmMrssmDims* d, const RssmMrssmWeights* w, const RssmMrssmInputs* in, const RssmMrssmOutputs* out,
                           void* stream) {
    return mrssm_fwd_common(d, w, in, out, stream, true);
}

int rssm_mrssm_rollout_bwd(const RssmMrssmDims* d, const RssmMrssmWeights* w, const RssmMrssmInputs* in, const RssmMrssmOutputs* fo,
                           const RssmMrssmUpstream* up, const RssmMrssmInputGrads* gin, const RssmMrssmWeightGrads* gw, void* stream) {
    if (check_mrssm(d)) return 1;
    REQUIRE(w); REQUIRE(in); REQUIRE(fo); REQUIRE(up); REQUIRE(gin);
    REQUIRE(in->actions); REQUIRE(in->embed_a); REQUIRE(in->embed_v); REQUIRE(in->h0); REQUIRE(in->z0);
    REQUIRE(fo->feature); REQUIRE(fo->prior_probs); REQUIRE(fo->post_probs); REQUIRE(fo->saved);
    REQUIRE(up->d_feature); REQUIRE(gin->d_embed_a); REQUIRE(gin->d_embed_v); REQUIRE(gin->d_h0); REQUIRE(gin->d_z0);
    REQUIRE(gin->dpre);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rssm::MrssmBwdArgs a{};
    a.B = d->B, a.T = d->T, a.A = d->A, a.K = d->K, a.kl_wq = up->kl_wq, a.kl_wp = up->kl_wp, a.w = *w;
    a.h0 = in->h0, a.feature = fo->feature, a.prior_probs = fo->prior_probs, a.post_probs = fo->post_probs, a.saved = fo->saved;
    a.d_feature = up->d_feature, a.d_prior_probs = up->d_prior_probs, a.d_post_probs = up->d_post_probs;
    a.d_prior_stoch = up->d_prior_stoch, a.d_kl = up->d_kl;
    a.dpre = gin->dpre, a.d_actions = gin->d_actions, a.d_embed_a = gin->d_embed_a, a.d_embed_v = gin->d_embed_v;
    a.d_h0 = gin->d_h0, a.d_z0 = gin->d_z0;
    g_launches.fetch_add(1);
    if (check_cuda(rssm::launch_mrssm_bwd(a, d->precision, s), "mrssm backward launch")) return 1;
    if (gw == nullptr) return 0;
    return rssm_mrssm_wgrad(d, in, fo, gin->dpre, gw, stream);
}

int rssm_mrssm_wgrad(const RssmMrssmDims* d, const RssmMrssmInputs* in, const RssmMrssmOutputs* fo, const void* dpre,
                     const RssmMrssmWeightGrads* gw, void* stream) {
    if (check_mrssm(d)) return 1;
    REQUIRE(in); REQUIRE(fo); REQUIRE(dpre); REQUIRE(gw);
    REQUIRE(in->actions); REQUIRE(in->embed_a); REQUIRE(in->embed_v); REQUIRE(in->h0); REQUIRE(in->z0);
    REQUIRE(fo->feature); REQUIRE(fo->saved);
    // staged-row layout and part ids: kernels.h (wgl_mr); record offsets: mrssm_kernels.cu (mrs / mrd)
    using namespace rssm::wgl_mr;
    const int A = d->A, F = 48, SV = MRSSM_SAVED_FLOATS, DPF = MRSSM_DPRE_FLOATS;
    const int RE = d->precision == RSSM_PRECISION_BF16 ? 2 : 4;  // record element size
    const void* sv = fo->saved;
    const float* feat = fo->feature;
    rssm::WgradMmaArgs j{};
    j.B = d->B, j.T = d->T;
    add_seg(j, DP, dpre, DPF, DPF, RE);
    add_seg(j, XASPZ, feat + 32, F, 16, 4, 1, in->z0, 16);  // z_prev
    add_seg(j, XASPA, in->actions, A, A);                   // action (+ zero pad to 8)
    add_seg(j, XH1, rec_at(sv, 0, RE), SV, 32, RE);         // asp hidden
    add_seg(j, XX2, rec_at(sv, 32, RE), SV, 32, RE);        // x2 (GRU input)
    add_seg(j, XHP, feat, F, 32, 4, 1, in->h0, 32);         // h_prev
    add_seg(j, XA, feat, F, 32);                            // [h | embed_a]
    add_seg(j, XA + 32, in->embed_a, 64, 64);
    add_seg(j, XV, feat, F, 32);                            // [h | embed_v]
    add_seg(j, XV + 32, in->embed_v, 64, 64);
    add_seg(j, HID, rec_at(sv, 192, RE), SV, 96, RE);       // prior / audio / vision hidden
    if (j.stride != STRIDE) return fail("internal: MRSSM staged row is %d columns, expected %d", j.stride, STRIDE);
    set_out(j, O_ASP1Z, gw->asp_w1 + A, A + 16, 16, gw->asp_b1);
    set_out(j, O_ASP1A, gw->asp_w1, A + 16, A);
    set_out(j, O_ASP2, gw->asp_w2, 32, 32, gw->asp_b2);
    set_out(j, O_IHR, gw->w_ih, 32, 32, gw->b_ih);
    set_out(j, O_IHZ, gw->w_ih + 32 * 32, 32, 32, gw->b_ih + 32);
    set_out(j, O_IHN, gw->w_ih + 64 * 32, 32, 32, gw->b_ih + 64);
    set_out(j, O_HHR, gw->w_hh, 32, 32, gw->b_hh);
    set_out(j, O_HHZ, gw->w_hh + 32 * 32, 32, 32, gw->b_hh + 32);
    set_out(j, O_HHN, gw->w_hh + 64 * 32, 32, 32, gw->b_hh + 64);
    set_out(j, O_P1, gw->pr_w1, 32, 32, gw->pr_b1);
    set_out(j, O_P2, gw->pr_w2, 32, 32, gw->pr_b2);
    set_out(j, O_A1A, gw->au_w1, 96, 96, gw->au_b1);
    set_out(j, O_A1B, gw->au_w1 + 16 * 96, 96, 96, gw->au_b1 + 16);
    set_out(j, O_A2, gw->au_w2, 32, 32, gw->au_b2);
    set_out(j, O_V1A, gw->vi_w1, 96, 96, gw->vi_b1);
    set_out(j, O_V1B_L, gw->vi_w1 + 16 * 96, 96, 48, gw->vi_b1 + 16);
    set_out(j, O_V1B_R, gw->vi_w1 + 16 * 96 + 48, 96, 48);
    set_out(j, O_V2, gw->vi_w2, 32, 32, gw->vi_b2);
    for (int i = 0; i < N_OUT; ++i)
        if (j.out[i].dW == nullptr) return fail("weight-gradient pointer of part %d is NULL", i);
    g_launches.fetch_add(1);
    return check_cuda(rssm::launch_wgrad_mma(j, 1, d->precision, static_cast<cudaStream_t>(stream)), "mrssm wgrad launch");
}

// ---------------------------------------------------------------------------------------------------
// MoPoE-MMTRSSM
// ---------------------------------------------------------------------------------------------------
static int mtrssm_fwd_common(const RssmMtrssmDims* d, const RssmMtrssmWeights* w, const RssmMtrssmInputs* in,
                             const RssmMtrssmOutputs* out, void* stream, bool imagine) {
    if (check_mtrssm(d)) return 1;
    REQUIRE(w); REQUIRE(in); REQUIRE(out);
    REQUIRE(in->actions); REQUIRE(in->deter_h0); REQUIRE(in->deter_l0); REQUIRE(in->hidden_h0); REQUIRE(in->hidden_l0);
    REQUIRE(in->stoch_h0); REQUIRE(in->stoch_l0);
    REQUIRE(out->feature); REQUIRE(out->hidden_h); REQUIRE(out->hidden_l); REQUIRE(out->prior_probs_h); REQUIRE(out->prior_probs_l);
    REQUIRE(w->l_d2h_w); REQUIRE(w->l_in_w); REQUIRE(w->h_d2h_w); REQUIRE(w->h_in_w); REQUIRE(w->lp_w1); REQUIRE(w->hp_w1);
    if ((in->u_prior_l == nullptr) != (in->u_prior_h == nullptr)) return fail("u_prior_l and u_prior_h must both be given or both NULL");
    if (imagine) {
        REQUIRE(in->u_prior_l);
    } else {
        REQUIRE(in->embed_a); REQUIRE(in->embed_v); REQUIRE(in->u_post_l); REQUIRE(in->u_post_h);
        REQUIRE(out->post_probs_h); REQUIRE(out->post_probs_l); REQUIRE(out->kl_l); REQUIRE(out->kl_h);
        REQUIRE(w->hq_w1); REQUIRE(w->au_w1); REQUIRE(w->vi_w1);
        if ((out->prior_stoch_l == nullptr) != (out->prior_stoch_h == nullptr))
            return fail("prior_stoch_l and prior_stoch_h must both be given or both NULL");
    }
    rssm::MtrssmFwdArgs a{};
    a.B = d->B, a.T = d->T, a.A = d->A, a.KL = d->KL, a.KH = d->KH;
    a.inv_tau_l = 1.f / d->l_tau, a.inv_tau_h = 1.f / d->h_tau, a.w = *w;
    a.actions = in->actions, a.embed_a = in->embed_a, a.embed_v = in->embed_v;
    a.deter_h0 = in->deter_h0, a.deter_l0 = in->deter_l0, a.hidden_h0 = in->hidden_h0, a.hidden_l0 = in->hidden_l0;
    a.stoch_h0 = in->stoch_h0, a.stoch_l0 = in->stoch_l0;
    a.u_post_l = in->u_post_l, a.u_post_h = in->u_post_h, a.u_prior_l = in->u_prior_l, a.u_prior_h = in->u_prior_h;
    a.feature = out->feature, a.hidden_h = out->hidden_h, a.hidden_l = out->hidden_l;
    a.prior_probs_h = out->prior_probs_h, a.prior_probs_l = out->prior_probs_l;
    a.post_probs_h = out->post_probs_h, a.post_probs_l = out->post_probs_l;
    a.prior_stoch_h = out->prior_stoch_h, a.prior_stoch_l = out->prior_stoch_l;
    a.kl_l = out->kl_l, a.kl_h = out->kl_h, a.saved = imagine ? nullptr : out->saved;
    a.saved_ld = mt_saved_ld(d->precision);
    g_launches.fetch_add(1);
    return check_cuda(rssm::launch_mtrssm_fwd(a, mt_kernel_precision(d->precision), imagine, static_cast<cudaStream_t>(stream)),
                      imagine ? "mtrssm imagine launch" : "mtrssm forward launch");
}

int rssm_mtrssm_rollout_fwd(const RssmMtrssmDims* d, const RssmMtrssmWeights* w, const RssmMtrssmInputs* in,
                            const RssmMtrssmOutputs* out, void* stream) {
    return mtrssm_fwd_common(d, w, in, out, stream, false);
}

int rssm_mtrssm_imagine_fwd(const RssmMtrssmDims* d, const RssmMtrssmWeights* w, const RssmMtrssmInputs* in,
                            const RssmMtrssmOutputs* out, void* stream) {
    return mtrssm_fwd_common(d, w, in, out, stream, true);
}

int rssm_mtrssm_rollout_bwd(const RssmMtrssmDims* d, const RssmMtrssmWeights* w, const RssmMtrssmInputs* in,
                            const RssmMtrssmOutputs* fo, const RssmMtrssmUpstream* up, const RssmMtrssmInputGrads* gin,
                            const RssmMtrssmWeightGrads* gw, void* stream) {
    if (check_mtrssm(d)) return 1;
    REQUIRE(w); REQUIRE(in); REQUIRE(fo); REQUIRE(up); REQUIRE(gin);
    REQUIRE(in->actions); REQUIRE(in->embed_a); REQUIRE(in->embed_v); REQUIRE(in->deter_h0); REQUIRE(in->deter_l0);
    REQUIRE(in->stoch_h0); REQUIRE(in->stoch_l0);
    REQUIRE(fo->feature); REQUIRE(fo->prior_probs_h); REQUIRE(fo->prior_probs_l); REQUIRE(fo->post_probs_h); REQUIRE(fo->post_probs_l);
    REQUIRE(fo->saved); REQUIRE(up->d_feature); REQUIRE(gin->d_embed_a); REQUIRE(gin->d_embed_v);
    const bool fused = gw != nullptr && d->precision == RSSM_PRECISION_BF16_FUSED;
    if (!fused) REQUIRE(gin->dpre);
    REQUIRE(gin->d_deter_h0); REQUIRE(gin->d_deter_l0); REQUIRE(gin->d_hidden_h0); REQUIRE(gin->d_hidden_l0);
    REQUIRE(gin->d_stoch_h0); REQUIRE(gin->d_stoch_l0);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rssm::MtrssmBwdArgs a{};
    a.B = d->B, a.T = d->T, a.A = d->A, a.KL = d->KL, a.KH = d->KH;
    a.inv_tau_l = 1.f / d->l_tau, a.inv_tau_h = 1.f / d->h_tau, a.kl_wq = up->kl_wq, a.kl_wp = up->kl_wp, a.w = *w;
    a.feature = fo->feature, a.prior_probs_h = fo->prior_probs_h, a.prior_probs_l = fo->prior_probs_l;
    a.post_probs_h = fo->post_probs_h, a.post_probs_l = fo->post_probs_l, a.saved = fo->saved, a.saved_ld = mt_saved_ld(d->precision);
    a.embed_a = in->embed_a, a.embed_v = in->embed_v, a.actions = in->actions;
    a.deter_h0 = in->deter_h0, a.deter_l0 = in->deter_l0, a.stoch_h0 = in->stoch_h0, a.stoch_l0 = in->stoch_l0;
    a.d_feature = up->d_feature, a.d_prior_probs_h = up->d_prior_probs_h, a.d_prior_probs_l = up->d_prior_probs_l;
    a.d_post_probs_h = up->d_post_probs_h, a.d_post_probs_l = up->d_post_probs_l;
    a.d_prior_stoch_h = up->d_prior_stoch_h, a.d_prior_stoch_l = up->d_prior_stoch_l, a.d_kl_l = up->d_kl_l, a.d_kl_h = up->d_kl_h;
    a.dpre = gin->dpre, a.d_actions = gin->d_actions, a.d_embed_a = gin->d_embed_a, a.d_embed_v = gin->d_embed_v;
    a.d_deter_h0 = gin->d_deter_h0, a.d_deter_l0 = gin->d_deter_l0, a.d_hidden_h0 = gin->d_hidden_h0;
    a.d_hidden_l0 = gin->d_hidden_l0, a.d_stoch_h0 = gin->d_stoch_h0, a.d_stoch_l0 = gin->d_stoch_l0;
    if (fused) {
        const float* const* gp = reinterpret_cast<const float* const*>(gw);
        for (size_t i = 0; i < sizeof(RssmMtrssmWeightGrads) / sizeof(float*); ++i)
            if (gp[i] == nullptr) return fail("weight-gradient pointer %d is NULL", (int)i);
        g_launches.fetch_add(1);
        return check_cuda(rssm::launch_mtrssm_bwd_fused(a, *gw, s), "mtrssm fused backward launch");
    }
    g_launches.fetch_add(1);
    if (check_cuda(rssm::launch_mtrssm_bwd(a, mt_kernel_precision(d->precision), s), "mtrssm backward launch")) return 1;
    if (gw == nullptr) return 0;
    return rssm_mtrssm_wgrad(d, in, fo, gin->dpre, gw, stream);
}

int rssm_mtrssm_wgrad(const RssmMtrssmDims* d, const RssmMtrssmInputs* in, const RssmMtrssmOutputs* fo, const void* dpre,
                      const RssmMtrssmWeightGrads* gw, void* stream) {
    if (check_mtrssm(d)) return 1;
    REQUIRE(in); REQUIRE(fo); REQUIRE(dpre); REQUIRE(gw);
    REQUIRE(in->actions); REQUIRE(in->embed_a); REQUIRE(in->embed_v); REQUIRE(in->deter_h0); REQUIRE(in->deter_l0);
    REQUIRE(in->stoch_h0); REQUIRE(in->stoch_l0); REQUIRE(fo->feature); REQUIRE(fo->saved);
    // staged-row layout and part ids: kernels.h (wgl_mt); record offsets: mtrssm_kernels.cu (mts / mtd);
    // feature = [d_h 0 | z_h 32 | d_l 48 | z_l 80]
    using namespace rssm::wgl_mt;
    const int A = d->A, F = 96, DPF = MTRSSM_DPRE_FLOATS, LDIN = A + 32;
    const int SV = mt_saved_ld(d->precision);
    const int RE = d->precision == RSSM_PRECISION_FP32 ? 4 : 2;  // record element size
    const void* sv = fo->saved;
    const float* feat = fo->feature;
    rssm::WgradMmaArgs j{};
    j.B = d->B, j.T = d->T;
    add_seg(j, DP, dpre, DPF, 304, RE);
    add_seg(j, XLD, feat + 48, F, 32, 4, 1, in->deter_l0, 32);       // d_l_prev
    add_seg(j, XLZ, feat + 80, F, 16, 4, 1, in->stoch_l0, 16);       // z_l_prev
    add_seg(j, XLZ + 16, feat + 32, F, 16, 4, 1, in->stoch_h0, 16);  // z_h_prev
    add_seg(j, XLA, in->actions, A, A);                               // action (+ zero pad to 8)
    add_seg(j, XHD, feat + 0, F, 32, 4, 1, in->deter_h0, 32);        // d_h_prev
    add_seg(j, XHI, feat + 32, F, 16, 4, 1, in->stoch_h0, 16);       // z_h_prev
    add_seg(j, XQ, feat + 48, F, 32);                            // [d_l | d_h]
    add_seg(j, XQ + 32, feat + 0, F, 32);
    add_seg(j, XA, feat + 48, F, 32);                            // [d_l | embed_a]
    add_seg(j, XA + 32, in->embed_a, 64, 64);
    add_seg(j, XV, feat + 48, F, 32);                            // [d_l | embed_v]
    add_seg(j, XV + 32, in->embed_v, 64, 64);
    add_seg(j, HID, sv, SV, 160, RE);                            // the five head hiddens
    if (j.stride != STRIDE) return fail("internal: MMTRSSM staged row is %d columns, expected %d", j.stride, STRIDE);
    // l_rnn: pre_l = _d2h(d_l_prev) + _input2h([action | z_l_prev | z_h_prev]); both biases see sum(dpre_l)
    set_out(j, O_LD, gw->l_d2h_w, 32, 32, gw->l_d2h_b, gw->l_in_b);
    set_out(j, O_LIZ, gw->l_in_w + A, LDIN, 32);
    set_out(j, O_LIA, gw->l_in_w, LDIN, A);
    set_out(j, O_HD, gw->h_d2h_w, 32, 32, gw->h_d2h_b, gw->h_in_b);
    set_out(j, O_HI, gw->h_in_w, 16, 16);
    set_out(j, O_LP1, gw->lp_w1, 32, 32, gw->lp_b1);
    set_out(j, O_LP2, gw->lp_w2, 32, 32, gw->lp_b2);
    set_out(j, O_HP1, gw->hp_w1, 32, 32, gw->hp_b1);
    set_out(j, O_HP2, gw->hp_w2, 32, 32, gw->hp_b2);
    set_out(j, O_HQ1, gw->hq_w1, 64, 64, gw->hq_b1);
    set_out(j, O_HQ2, gw->hq_w2, 32, 32, gw->hq_b2);
    set_out(j, O_A1A, gw->au_w1, 96, 96, gw->au_b1);
    set_out(j, O_A1B, gw->au_w1 + 16 * 96, 96, 96, gw->au_b1 + 16);
    set_out(j, O_A2, gw->au_w2, 32, 32, gw->au_b2);
    set_out(j, O_V1A, gw->vi_w1, 96, 96, gw->vi_b1);
    set_out(j, O_V1B, gw->vi_w1 + 16 * 96, 96, 96, gw->vi_b1 + 16);
    set_out(j, O_V2, gw->vi_w2, 32, 32, gw->vi_b2);
    for (int i = 0; i < N_OUT; ++i)
        if (j.out[i].dW == nullptr) return fail("weight-gradient pointer of part %d is NULL", i);
    g_launches.fetch_add(1);
    // bf16 records with the padded row lengths: slab-staged kernel (one bulk copy per source tensor and 32-row block)
    const bool slab_ok = d->precision == RSSM_PRECISION_BF16 && aligned16(dpre) && aligned16(sv) && aligned16(feat) &&
                         aligned16(in->embed_a) && aligned16(in->embed_v) && aligned16(in->deter_l0) && aligned16(in->deter_h0) &&
                         aligned16(in->stoch_l0) && aligned16(in->stoch_h0) && getenv("RSSM_WGRAD_GENERIC") == nullptr;
    if (slab_ok) {
        rssm::WgradMtSlabArgs k{};
        k.B = d->B, k.T = d->T, k.A = A;
        k.dpre = static_cast<const __nv_bfloat16*>(dpre), k.saved = static_cast<const __nv_bfloat16*>(sv);
        k.feature = feat, k.embed_a = in->embed_a, k.embed_v = in->embed_v, k.actions = in->actions;
        k.deter_l0 = in->deter_l0, k.deter_h0 = in->deter_h0, k.stoch_l0 = in->stoch_l0, k.stoch_h0 = in->stoch_h0;
        for (int i = 0; i < N_OUT; ++i) k.out[i] = j.out[i];
        return check_cuda(rssm::launch_wgrad_mt_slab(k, static_cast<cudaStream_t>(stream)), "mtrssm wgrad (slab) launch");
    }
    return check_cuda(rssm::launch_wgrad_mma(j, 0, mt_kernel_precision(d->precision), static_cast<cudaStream_t>(stream)), "mtrssm wgrad launch");
}

}  // extern "C"
