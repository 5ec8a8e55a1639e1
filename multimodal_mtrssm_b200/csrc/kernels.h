// Internal launch interface between the C-ABI layer (rollout_abi.cu) and the kernels.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/rssm_rollout.h"

namespace rssm {

// ---- MoPoE-MRSSM ----------------------------------------------------------------------------------
struct MrssmFwdArgs {
    int B, T, A, K;
    int unimodal;  // posterior = the audio head alone, no fusion (BaseRSSM.rollout_representation)
    RssmMrssmWeights w;
    const float *actions, *embed_a, *embed_v, *h0, *z0, *u_post, *u_prior;
    float *feature, *prior_probs, *post_probs, *prior_stoch, *kl;
    void* saved;  // Rec<NS>::T [B,T,MRSSM_SAVED_FLOATS]
};

struct MrssmBwdArgs {
    int B, T, A, K;
    int unimodal;
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


// ---- wide MoPoE-MRSSM (deter = hidden = D, D % 64 == 0; persistent tcgen05 kernels: wide_common.cuh) --------------------
struct MrssmWideFwdArgs {
    int B, T, A, K, D;  // B = sequences of this launch group (<= NBB * 128)
    int NBB, NSL;       // batch blocks of the group, slices (D / 32); grid = NBB * NSL
    int imagine, desc_swap;
    long long plane_stride, t_stride, emb_t_stride;  // record / packed-embedding strides in elements (t_stride 0: keep one step)
    RssmMrssmWeights w;                               // fp32 originals (biases, first projector layer, logit layers)
    const __nv_bfloat16 *pW2, *pWhh, *pWih, *pWhd, *pWae, *pWve;  // packed bf16 weight slices
    const float *actions, *h0, *z0, *u_post, *u_prior;           // first row of the group
    const __nv_bfloat16 *emb_a, *emb_v;                          // packed [T][blocks][8][128][8], first block of the group
    float *feature, *prior_probs, *post_probs, *prior_stoch, *kl;
    __nv_bfloat16 *rec, *h0p;  // record planes [t][plane][blocks][D/8][128][8]; packed h0 [NBB][D/8][128][8]
    float* part;               // partial logits [NBB*128][NSL][48]
    float* logits;             // [rows][T][32] audio / vision logits for the backward (NULL when not saving)
    unsigned long long* timing;  // debug: globaltimer at the end of each phase of the first steps (CTA 0), or NULL
    int exp;                     // debug (RSSM_WIDE_EXP): 1 = skip the operand copies, 2 = skip the MMAs (timing experiments, wrong results)
    int cs;                      // CTAs per cluster along the slice index (1, 2 or 4): activation chunks are multicast inside a cluster
    unsigned* bar;
    int* status;
};
cudaError_t launch_mrssm_wide_fwd(const MrssmWideFwdArgs& a, cudaStream_t s);
cudaError_t launch_wide_persistent(const void* kernel, void* args, int grid, int cs, size_t smem, cudaStream_t s);
size_t mrssm_wide_fwd_smem(int D);


struct MrssmWideBwdArgs {
    int B, T, A, K, D, NBB, NSL;
    float kl_wq, kl_wp;
    long long plane_stride, t_stride, dt_stride;  // elements: one plane, forward record step, gradient-plane step
    long long dlg_off, xin_off;                   // element offsets of the narrow planes inside a gradient-plane step (group-relative)
    RssmMrssmWeights w;
    const __nv_bfloat16 *pW1x, *pWhdT, *pWgT, *pWihTn, *pWhhTn, *pW2T;  // transposed packed weight images
    const __nv_bfloat16* rec;  // forward record planes (first block of the group)
    __nv_bfloat16* drec;       // gradient planes (first block of the group)
    const float* stat;         // per-row statistics of the pre-pass [t][blocks][32][128][4], first block of the group
    long long stat_t_stride;   // floats
    const float *feature, *h0;
    const float* d_feature;
    float *d_actions, *d_h0, *d_z0;
    unsigned* bar;
    int* status;
    unsigned long long* timing;  // debug, as in the forward
    int exp;
    int cs;  // as in the forward
};
cudaError_t launch_mrssm_wide_bwd(const MrssmWideBwdArgs& a, cudaStream_t s);

// pre-pass of the backward over all (b,t) rows (padded to whole 128-row blocks): the part of the row-wise distribution
// backward that does not depend on the carried gradient, and the [z_{t-1} | a_t] operand plane
struct WideRowstatArgs {
    int B, T, A, K, D, NBBT;
    float kl_wq, kl_wp;
    long long dt_stride;  // elements between steps of the xin plane
    const float *logits, *feature, *prior_probs, *post_probs, *z0, *actions;
    const float *d_feature, *d_prior_probs, *d_post_probs, *d_prior_stoch, *d_kl;
    float* stat;
    __nv_bfloat16* xin;   // step 0, block 0 of the narrow [32]-feature plane
};
cudaError_t launch_wide_bwd_rowstat(const WideRowstatArgs& a, cudaStream_t s);
cudaError_t launch_wide_pack_bwd_weights(const MrssmWideBwdArgs& a, cudaStream_t s);

// one output tile of the weight-gradient contraction over (b,t) rows: dW[m * sm + n * sn] += sum_rows Y[row][m] X[row][n]
struct WideWgradTile {
    const __nv_bfloat16 *y, *x, *x0;  // first element of the tile's 128 Y features / N X features in block 0 of step 0
    long long y_tstride, y_bstride, x_tstride, x_bstride;
    int x_shift;                      // 1: X of step t-1 (x0 for t = 0)
    int N, mvalid, nvalid;
    float *dW, *db;                   // db (optional): += sum_rows Y[row][m]
    long long sm, sn;
};
// The tile table travels to the device as a KERNEL PARAMETER of a tiny upload kernel (captured by value under CUDA-graph capture;
// no pageable host-to-device copy, no implicit stream synchronisation, nothing a later call could change behind a captured graph).
constexpr int MAX_WIDE_WGRAD_TILES = 256;  // 256 x 104 B = 26.6 KB < the 32,764-byte kernel parameter limit
struct WideWgradTileTable {
    int n;
    WideWgradTile t[MAX_WIDE_WGRAD_TILES];
};
cudaError_t launch_wide_wgrad_tiles_upload(const WideWgradTileTable& table, WideWgradTile* dst, cudaStream_t s);
cudaError_t launch_wide_wgrad(const WideWgradTile* tiles_dev, int ntiles, int nsplit, int T, int NBBT, const __nv_bfloat16* ones, cudaStream_t s);

struct WideDembedArgs {
    int B, T, D, NBBT;
    long long dt_stride;
    const __nv_bfloat16 *dah, *dvh, *pWaeT, *pWveT;
    float *d_embed_a, *d_embed_v;
};
cudaError_t launch_wide_dembed(const WideDembedArgs& a, cudaStream_t s);
cudaError_t launch_wide_pack_dembed(const float* au_w1, const float* vi_w1, int D, __nv_bfloat16* dstA, __nv_bfloat16* dstV, __nv_bfloat16* ones,
                                    cudaStream_t s);

// packs fp32 [out,in] weights into per-slice bf16 operand images [slice][K/64][8][32*nparts][8]:
// row (part p, unit u) of slice s = src[p] + (s*32+u) * ld[p] + coloff + k
struct WidePackJob {
    const float* src[3];
    int ld[3];
    int coloff, nparts, K;
    __nv_bfloat16* dst;
};
constexpr int MAX_WIDE_PACK_JOBS = 8;
struct WidePackJobs {
    WidePackJob job[MAX_WIDE_PACK_JOBS];
    int njobs, NSL;
};
cudaError_t launch_wide_pack_weights(const WidePackJobs& jobs, cudaStream_t s);
// fp32 rows src[(b*T + t) * ld + coloff + f], f < F (F % 8 == 0)  ->  packed bf16 dst[t][blocks][F/8][128][8]
cudaError_t launch_wide_pack_rows(const float* src, int B, int T, int F, int ld, int coloff, __nv_bfloat16* dst, int blocks,
                                  cudaStream_t s);

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
    int obs_projected;  // 1: embed_a / embed_v are the pre-multiplied first-layer partials [B,T,32] (include/rssm_rollout.h)
    int ld_feature;  // 0 = dense outputs; MTRSSM_ROW_PITCH = grouped rows (RssmMtrssmOutputs.ld_*: all pitches 256, kl 2)
    int rec_tiled;      // 1: `saved` is TILE-BLOCKED: [ceil(B/16)][T][26 chunks][16 rows][8 bf16] -- a tile-step's record is one contiguous
                        // 6.5 KB block already in the tcgen05 operand-image order (RSSM_PRECISION_BF16_FUSED); 0: [B][T][saved_ld]
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
    const float *d_hidden_h, *d_hidden_l;  // upstream gradients of the MTRNN.hidden outputs, dense [B,T,32]; may be null
    void* dpre;
    float *d_actions, *d_embed_a, *d_embed_v;
    float *d_deter_h0, *d_deter_l0, *d_hidden_h0, *d_hidden_l0, *d_stoch_h0, *d_stoch_l0;
    int obs_projected;  // 1: d_embed_a / d_embed_v are [B,T,32] = d of the pre-multiplied partials; no embedding operand images
    int rec_tiled;      // as in MtrssmFwdArgs (the fused backward stages the tile-blocked record as contiguous runs)
    int ld_feature;  // 0 = dense forward outputs; MTRSSM_ROW_PITCH = grouped rows
};

cudaError_t launch_mtrssm_fwd(const MtrssmFwdArgs& a, int precision, bool imagine, cudaStream_t s);
cudaError_t launch_mtrssm_fwd2(const MtrssmFwdArgs& a, cudaStream_t s);  // bf16 policy, two warps per tile
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

// ---- Gaussian reconstruction likelihood (likelihood_kernel.cu) ---------------------------------------
struct NllSeg {
    const void* prediction;  // pred_dtype elements
    const float* target;
    size_t n;                // elements
    double sq_coeff;         // 0.5 / (scale^2 * n_batch)
    double constant;         // n_event * (log(scale) + 0.5 log(2 pi))
    float* loss;             // forward: device scalar out
    const float* d_loss;     // backward: device scalar in (NULL = 1)
    void* d_prediction;      // backward out, pred_dtype elements
    float* d_target;         // backward out, may be NULL
};
struct NllArgs {
    NllSeg seg[RSSM_NLL_MAX_SEGMENTS];
    int nseg;
    int pred_dtype;  // RSSM_DTYPE_*
    double* partials;    // [nseg][ctas]
    unsigned* tickets;   // [RSSM_NLL_MAX_SEGMENTS], zero between launches
};
int nll_ctas_per_segment(int nseg);
cudaError_t launch_gaussian_nll(const NllArgs& a, bool backward, cudaStream_t s);

// ---- one-shot peer-memory allreduce of the gradient bucket (p2p_allreduce.cu) ----------------------------------------------
constexpr int P2P_MAX_RANKS = 8;
struct P2pAllreduceArgs {
    const float* data[P2P_MAX_RANKS];  // every rank's bucket of this slot (peer-mapped; [rank] = the local one)
    uint32_t* flags[P2P_MAX_RANKS];    // every rank's flag rows [2 slots][P2P_MAX_RANKS] (peer-mapped)
    float* out;                        // local result (mean over ranks), n floats
    uint32_t* status;                  // local word, set to 1 on timeout
    size_t n;
    int world, rank, slot;
    uint32_t epoch;
    long long timeout_cycles;
};
cudaError_t launch_p2p_allreduce_mean(const P2pAllreduceArgs& a, cudaStream_t s);

}  // namespace rssm
