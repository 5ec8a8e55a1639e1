// Internal launch interface between the C-ABI layer (rollout_abi.cu) and the kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/rssm_rollout.h"

namespace rssm {

// ---- MoPoE-MRSSM ----------------------------------------------------------------------------------
struct MrssmFwdArgs {
    int B, T, A, K;
    RssmMrssmWeights w;
    const float *actions, *embed_a, *embed_v, *h0, *z0, *u_post, *u_prior;
    float *feature, *prior_probs, *post_probs, *prior_stoch, *kl;
    void* saved;  // Rec<NS>::T [B,T,MRSSM_SAVED_FLOATS]
};

struct MrssmBwdArgs {
    int B, T, A, K;
    float kl_wq, kl_wp;
    RssmMrssmWeights w;
    const float *h0, *feature, *prior_probs, *post_probs;
    const void* saved;
    const float *d_feature, *d_prior_probs, *d_post_probs, *d_prior_stoch, *d_kl;
    void* dpre;
    float *d_actions, *d_embed_a, *d_embed_v, *d_h0, *d_z0;
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
    float *kl_l, *kl_h;
    void* saved;
    int saved_ld;    // elements per (b,t) row of `saved`
};

struct MtrssmBwdArgs {
    int B, T, A, KL, KH;
    float inv_tau_l, inv_tau_h, kl_wq, kl_wp;
    RssmMtrssmWeights w;
    const float *feature, *prior_probs_h, *prior_probs_l, *post_probs_h, *post_probs_l;
    // forward inputs read by the fused backward only (X operands of the weight-gradient MMAs)
    const float *embed_a, *embed_v, *actions, *deter_h0, *deter_l0, *stoch_h0, *stoch_l0;
    const void* saved;
    int saved_ld;  // elements per (b,t) row of `saved`
    const float *d_feature, *d_prior_probs_h, *d_prior_probs_l, *d_post_probs_h, *d_post_probs_l;
    const float *d_prior_stoch_h, *d_prior_stoch_l, *d_kl_l, *d_kl_h;
    void* dpre;
    float *d_actions, *d_embed_a, *d_embed_v;
    float *d_deter_h0, *d_deter_l0, *d_hidden_h0, *d_hidden_l0, *d_stoch_h0, *d_stoch_l0;
};

cudaError_t launch_mtrssm_fwd(const MtrssmFwdArgs& a, int precision, bool imagine, cudaStream_t s);
cudaError_t launch_mtrssm_bwd(const MtrssmBwdArgs& a, int precision, cudaStream_t s);
// bf16 path: BPTT + weight gradients in one kernel (tcgen05 / TMEM accumulators); ADDS into g (mtrssm_fused_bwd.cu)
cudaError_t launch_mtrssm_bwd_fused(const MtrssmBwdArgs& a, const RssmMtrssmWeightGrads& g, cudaStream_t s);

// ---- batched weight gradients on the tensor cores (wgrad_kernel.cu) -----------------------------------------------
// A staged shared-memory row holds, per (b,t), the dpre record followed by every layer's input, as bf16 columns.
struct WgradSeg {       // one source column range copied into the staged row, in chunks of 4 elements
    const char* ptr;    // row (b,t) at ptr + (b*T + t) * ld_bytes          (shift == 0)
    const char* ptr0;   // shift == 1: row (b,t-1) for t > 0, ptr0 + b * ld0_bytes for t == 0
    int ld_bytes, ld0_bytes;
    int c4_begin, c4_end;  // staged 4-element chunk range [begin, end) of this segment (staged column = 4 * chunk)
    int valid;             // real source elements; the rest of the range is zero padding
    int shift;
    int kind;              // 0: fp32, 16-byte loads; 1: fp32, guarded scalar loads; 2: bf16, 8-byte copies
};
struct WgradOut {  // destination of one part: dW[n][k] at dW + n*ldw + k for k < kvalid; up to two bias vectors
    float* dW;
    float* db0;
    float* db1;
    int ldw, kvalid;
};
constexpr int MAX_WGRAD_SEGS = 16, MAX_WGRAD_OUTS = 24;
struct WgradMmaArgs {
    int B, T, nseg, stride;  // stride: staged row length in bf16 elements = 4 * (last c4_end)
    WgradSeg seg[MAX_WGRAD_SEGS];
    WgradOut out[MAX_WGRAD_OUTS];
};
// model: 0 = MMTRSSM, 1 = MRSSM
cudaError_t launch_wgrad_mma(const WgradMmaArgs& a, int model, int precision, cudaStream_t s);

// MMTRSSM, bf16 records of MTRSSM_DPRE_FLOATS / MTRSSM_SAVED_FLOATS elements per row: slab-staged variant (one bulk copy per
// source tensor and 32-row block); same outputs (`out`, indexed by wgl_mt part ids) as the generic kernel
struct WgradMtSlabArgs {
    int B, T, A;
    const __nv_bfloat16 *dpre, *saved;
    const float *feature, *embed_a, *embed_v, *actions;
    const float *deter_l0, *deter_h0, *stoch_l0, *stoch_h0;
    WgradOut out[MAX_WGRAD_OUTS];
};
cudaError_t launch_wgrad_mt_slab(const WgradMtSlabArgs& a, cudaStream_t s);

// staged-row column layouts (bf16 elements) and part ids
namespace wgl_mt {
constexpr int DP = 0, XLD = 304, XLZ = 336, XLA = 368, XHD = 376, XHI = 408, XQ = 424, XA = 488, XV = 584, HID = 680, STRIDE = 840;
enum { O_LD, O_LIZ, O_LIA, O_HD, O_HI, O_LP1, O_LP2, O_HP1, O_HP2, O_HQ1, O_HQ2, O_A1A, O_A1B, O_A2, O_V1A, O_V1B, O_V2, N_OUT };
}  // namespace wgl_mt
namespace wgl_mr {
constexpr int DP = 0, XASPZ = 336, XASPA = 352, XH1 = 360, XX2 = 392, XHP = 424, XA = 456, XV = 552, HID = 648, STRIDE = 744;
enum { O_ASP1Z, O_ASP1A, O_ASP2, O_IHR, O_IHZ, O_IHN, O_HHR, O_HHZ, O_HHN, O_P1, O_P2, O_A1A, O_A1B, O_A2, O_V1A, O_V1B_L, O_V1B_R, O_V2, N_OUT };
}  // namespace wgl_mr
static_assert((wgl_mt::STRIDE * 2 / 16) % 2 == 1 && (wgl_mr::STRIDE * 2 / 16) % 2 == 1, "row stride must be an odd number of 16B chunks (ldmatrix bank-conflict free)");

}  // namespace rssm
