// Internal launch interface between the C-ABI layer (rollout_abi.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>

#include "../../include/rssm_rollout.h"

namespace rssm {

// ---- MoPoE-MRSSM ----------------------------------------------------------------------------------
struct MrssmFwdArgs {
    int B, T, A, K;
    RssmMrssmWeights w;
    const float *actions, *embed_a, *embed_v, *h0, *z0, *u_post, *u_prior;
    float *feature, *prior_probs, *post_probs, *prior_stoch, *kl, *saved;
};

struct MrssmBwdArgs {
    int B, T, A, K;
    float kl_wq, kl_wp;
    RssmMrssmWeights w;
    const float *h0, *feature, *prior_probs, *post_probs, *saved;
    const float *d_feature, *d_prior_probs, *d_post_probs, *d_prior_stoch, *d_kl;
    float *dpre, *d_actions, *d_embed_a, *d_embed_v, *d_h0, *d_z0;
};

cudaError_t launch_mrssm_fwd(const MrssmFwdArgs& a, int precision, bool imagine, cudaStream_t s);
cudaError_t launch_mrssm_bwd(const MrssmBwdArgs& a, int precision, cudaStream_t s);

// ---- MoPoE-MMTRSSM ---------------------------------------------------------------------------------
struct MtrssmFwdArgs {
    int B, T, A, KL, KH;
    float inv_tau_l, inv_tau_h;
    RssmMtrssmWeights w;
    const float *actions, *embed_a, *embed_v;
    const float *deter_h0, *deter_l0, *hidden_h0, *hidden_l0, *stoch_h0, *stoch_l0;
    const float *u_post_l, *u_post_h, *u_prior_l, *u_prior_h;
    float *feature, *hidden_h, *hidden_l;
    float *prior_probs_h, *prior_probs_l, *post_probs_h, *post_probs_l, *prior_stoch_h, *prior_stoch_l;
    float *kl_l, *kl_h, *saved;
};

struct MtrssmBwdArgs {
    int B, T, A, KL, KH;
    float inv_tau_l, inv_tau_h, kl_wq, kl_wp;
    RssmMtrssmWeights w;
    const float *feature, *prior_probs_h, *prior_probs_l, *post_probs_h, *post_probs_l, *saved;
    const float *d_feature, *d_prior_probs_h, *d_prior_probs_l, *d_post_probs_h, *d_post_probs_l;
    const float *d_prior_stoch_h, *d_prior_stoch_l, *d_kl_l, *d_kl_h;
    float *dpre, *d_actions, *d_embed_a, *d_embed_v;
    float *d_deter_h0, *d_deter_l0, *d_hidden_h0, *d_hidden_l0, *d_stoch_h0, *d_stoch_l0;
};

cudaError_t launch_mtrssm_fwd(const MtrssmFwdArgs& a, int precision, bool imagine, cudaStream_t s);
cudaError_t launch_mtrssm_bwd(const MtrssmBwdArgs& a, int precision, cudaStream_t s);

// ---- batched weight gradient: dW[n][col0 + k] += sum_rows dY[row][n] * X[row'][k] ------------------
// rows = (b,t), b < B, t < T.  X row for (b,t): shift == 0 -> X[(b*T + t) * ldx]; shift == 1 -> the
// PREVIOUS step's row, X[(b*T + t - 1) * ldx] for t > 0 and X0[b * ldx0] for t == 0.
struct WgradJob {
    const float* dY;
    const float* X;
    const float* X0;
    float* dW;
    float* db;  // optional: db[n] += sum_rows dY[row][n]
    int ldy, ldx, ldx0, ldw, N, K, shift;
};
constexpr int MAX_WGRAD_JOBS = 24;
struct WgradArgs {
    int B, T, njobs;
    WgradJob jobs[MAX_WGRAD_JOBS];
};
cudaError_t launch_wgrad(const WgradArgs& a, cudaStream_t s);

}  // namespace rssm
