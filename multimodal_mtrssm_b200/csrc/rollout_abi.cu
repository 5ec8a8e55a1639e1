// C-ABI entry points of librssm_rollout.so (declared in include/rssm_rollout.h).
// Validates sizes, builds kernel arguments and the weight-gradient job lists, launches on the
// caller's stream.  No hidden state; errors are reported through a thread-local message.
#include <atomic>
#include <cstdarg>
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
    if (d->precision != RSSM_PRECISION_FP32 && d->precision != RSSM_PRECISION_BF16) return fail("bad precision %d", d->precision);
    return 0;
}

#define REQUIRE(ptr)                                                           \
    do {                                                                       \
        if ((ptr) == nullptr) return fail("required pointer %s is NULL", #ptr); \
    } while (0)

void add_job(rssm::WgradArgs& a, const float* dY, int ldy, int N, const float* X, int ldx, int K, float* dW, int ldw, float* db,
             int shift = 0, const float* X0 = nullptr, int ldx0 = 0) {
    rssm::WgradJob& j = a.jobs[a.njobs++];
    j.dY = dY, j.ldy = ldy, j.N = N, j.X = X, j.ldx = ldx, j.K = K, j.dW = dW, j.ldw = ldw, j.db = db;
    j.shift = shift, j.X0 = X0 ? X0 : X, j.ldx0 = X0 ? ldx0 : ldx;
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

    // weight gradients: dW = dpre^T . layer input  (record offsets: mrssm_kernels.cu, namespaces mrs / mrd)
    const int A = d->A, F = 48, SV = MRSSM_SAVED_FLOATS, DP = MRSSM_DPRE_FLOATS;
    const float *dp = gin->dpre, *sv = fo->saved, *feat = fo->feature;
    rssm::WgradArgs j{};
    j.B = d->B, j.T = d->T;
    // action_state_projector.0 : input [action | z_prev]
    add_job(j, dp + 0, DP, 32, in->actions, A, A, gw->asp_w1, A + 16, gw->asp_b1);
    add_job(j, dp + 0, DP, 32, feat + 32, F, 16, gw->asp_w1 + A, A + 16, nullptr, 1, in->z0, 16);
    // action_state_projector.2 : input asp hidden
    add_job(j, dp + 32, DP, 32, sv + 0, SV, 32, gw->asp_w2, 32, gw->asp_b2);
    // GRU: weight_ih <- [dpre_r,dpre_z,dpre_n] x x2 ; weight_hh <- [dpre_r,dpre_z | d h_n] x h_prev
    add_job(j, dp + 64, DP, 96, sv + 32, SV, 32, gw->w_ih, 32, gw->b_ih);
    add_job(j, dp + 64, DP, 64, feat, F, 32, gw->w_hh, 32, gw->b_hh, 1, in->h0, 32);
    add_job(j, dp + 160, DP, 32, feat, F, 32, gw->w_hh + 64 * 32, 32, gw->b_hh + 64, 1, in->h0, 32);
    // prior head
    add_job(j, dp + 192, DP, 32, feat, F, 32, gw->pr_w1, 32, gw->pr_b1);
    add_job(j, dp + 224, DP, 16, sv + 192, SV, 32, gw->pr_w2, 32, gw->pr_b2);
    // audio / vision heads : input [deter | embed]
    add_job(j, dp + 240, DP, 32, feat, F, 32, gw->au_w1, 96, gw->au_b1);
    add_job(j, dp + 240, DP, 32, in->embed_a, 64, 64, gw->au_w1 + 32, 96, nullptr);
    add_job(j, dp + 272, DP, 16, sv + 224, SV, 32, gw->au_w2, 32, gw->au_b2);
    add_job(j, dp + 288, DP, 32, feat, F, 32, gw->vi_w1, 96, gw->vi_b1);
    add_job(j, dp + 288, DP, 32, in->embed_v, 64, 64, gw->vi_w1 + 32, 96, nullptr);
    add_job(j, dp + 320, DP, 16, sv + 256, SV, 32, gw->vi_w2, 32, gw->vi_b2);
    for (int i = 0; i < j.njobs; ++i)
        if (j.jobs[i].dW == nullptr) return fail("weight-gradient pointer of job %d is NULL", i);
    g_launches.fetch_add(1);
    return check_cuda(rssm::launch_wgrad(j, s), "mrssm wgrad launch");
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
    g_launches.fetch_add(1);
    return check_cuda(rssm::launch_mtrssm_fwd(a, d->precision, imagine, static_cast<cudaStream_t>(stream)),
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
    REQUIRE(fo->saved); REQUIRE(up->d_feature); REQUIRE(gin->d_embed_a); REQUIRE(gin->d_embed_v); REQUIRE(gin->dpre);
    REQUIRE(gin->d_deter_h0); REQUIRE(gin->d_deter_l0); REQUIRE(gin->d_hidden_h0); REQUIRE(gin->d_hidden_l0);
    REQUIRE(gin->d_stoch_h0); REQUIRE(gin->d_stoch_l0);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rssm::MtrssmBwdArgs a{};
    a.B = d->B, a.T = d->T, a.A = d->A, a.KL = d->KL, a.KH = d->KH;
    a.inv_tau_l = 1.f / d->l_tau, a.inv_tau_h = 1.f / d->h_tau, a.kl_wq = up->kl_wq, a.kl_wp = up->kl_wp, a.w = *w;
    a.feature = fo->feature, a.prior_probs_h = fo->prior_probs_h, a.prior_probs_l = fo->prior_probs_l;
    a.post_probs_h = fo->post_probs_h, a.post_probs_l = fo->post_probs_l, a.saved = fo->saved;
    a.d_feature = up->d_feature, a.d_prior_probs_h = up->d_prior_probs_h, a.d_prior_probs_l = up->d_prior_probs_l;
    a.d_post_probs_h = up->d_post_probs_h, a.d_post_probs_l = up->d_post_probs_l;
    a.d_prior_stoch_h = up->d_prior_stoch_h, a.d_prior_stoch_l = up->d_prior_stoch_l, a.d_kl_l = up->d_kl_l, a.d_kl_h = up->d_kl_h;
    a.dpre = gin->dpre, a.d_actions = gin->d_actions, a.d_embed_a = gin->d_embed_a, a.d_embed_v = gin->d_embed_v;
    a.d_deter_h0 = gin->d_deter_h0, a.d_deter_l0 = gin->d_deter_l0, a.d_hidden_h0 = gin->d_hidden_h0;
    a.d_hidden_l0 = gin->d_hidden_l0, a.d_stoch_h0 = gin->d_stoch_h0, a.d_stoch_l0 = gin->d_stoch_l0;
    g_launches.fetch_add(1);
    if (check_cuda(rssm::launch_mtrssm_bwd(a, d->precision, s), "mtrssm backward launch")) return 1;
    if (gw == nullptr) return 0;

    // record offsets: mtrssm_kernels.cu, namespaces mts / mtd; feature = [d_h 0 | z_h 32 | d_l 48 | z_l 80]
    const int A = d->A, F = 96, SV = MTRSSM_SAVED_FLOATS, DP = MTRSSM_DPRE_FLOATS, LDIN = A + 32;
    const float *dp = gin->dpre, *sv = fo->saved, *feat = fo->feature;
    rssm::WgradArgs j{};
    j.B = d->B, j.T = d->T;
    // l_rnn: pre_l = _d2h(d_l_prev) + _input2h([action | z_l_prev | z_h_prev]); both biases see sum(dpre_l)
    add_job(j, dp + 0, DP, 32, feat + 48, F, 32, gw->l_d2h_w, 32, gw->l_d2h_b, 1, in->deter_l0, 32);
    add_job(j, dp + 0, DP, 32, in->actions, A, A, gw->l_in_w, LDIN, gw->l_in_b);
    add_job(j, dp + 0, DP, 32, feat + 80, F, 16, gw->l_in_w + A, LDIN, nullptr, 1, in->stoch_l0, 16);
    add_job(j, dp + 0, DP, 32, feat + 32, F, 16, gw->l_in_w + A + 16, LDIN, nullptr, 1, in->stoch_h0, 16);
    // h_rnn: pre_h = _d2h(d_h_prev) + _input2h(z_h_prev)
    add_job(j, dp + 32, DP, 32, feat + 0, F, 32, gw->h_d2h_w, 32, gw->h_d2h_b, 1, in->deter_h0, 32);
    add_job(j, dp + 32, DP, 32, feat + 32, F, 16, gw->h_in_w, 16, gw->h_in_b, 1, in->stoch_h0, 16);
    // l_prior / h_prior / h_posterior
    add_job(j, dp + 64, DP, 32, feat + 48, F, 32, gw->lp_w1, 32, gw->lp_b1);
    add_job(j, dp + 96, DP, 16, sv + 0, SV, 32, gw->lp_w2, 32, gw->lp_b2);
    add_job(j, dp + 112, DP, 32, feat + 0, F, 32, gw->hp_w1, 32, gw->hp_b1);
    add_job(j, dp + 144, DP, 16, sv + 32, SV, 32, gw->hp_w2, 32, gw->hp_b2);
    add_job(j, dp + 160, DP, 32, feat + 48, F, 32, gw->hq_w1, 64, gw->hq_b1);
    add_job(j, dp + 160, DP, 32, feat + 0, F, 32, gw->hq_w1 + 32, 64, nullptr);
    add_job(j, dp + 192, DP, 16, sv + 64, SV, 32, gw->hq_w2, 32, gw->hq_b2);
    // audio / vision heads on [d_l | embed]
    add_job(j, dp + 208, DP, 32, feat + 48, F, 32, gw->au_w1, 96, gw->au_b1);
    add_job(j, dp + 208, DP, 32, in->embed_a, 64, 64, gw->au_w1 + 32, 96, nullptr);
    add_job(j, dp + 240, DP, 16, sv + 96, SV, 32, gw->au_w2, 32, gw->au_b2);
    add_job(j, dp + 256, DP, 32, feat + 48, F, 32, gw->vi_w1, 96, gw->vi_b1);
    add_job(j, dp + 256, DP, 32, in->embed_v, 64, 64, gw->vi_w1 + 32, 96, nullptr);
    add_job(j, dp + 288, DP, 16, sv + 128, SV, 32, gw->vi_w2, 32, gw->vi_b2);
    for (int i = 0; i < j.njobs; ++i)
        if (j.jobs[i].dW == nullptr) return fail("weight-gradient pointer of job %d is NULL", i);
    g_launches.fetch_add(1);
    return check_cuda(rssm::launch_wgrad(j, s), "mtrssm wgrad launch");
}

}  // extern "C"
