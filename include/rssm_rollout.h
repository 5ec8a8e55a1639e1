/* rssm_rollout.h -- C ABI of the B200 latent-rollout library (librssm_rollout.so).
 *
 * The reference (Mamo1031/Multimodal-MTRSSM) is pure Python and has NO FFI for this path; its boundary
 * is a set of Python methods.  Each entry point below replaces the body of one of them (paths relative
 * to /root/reference/src/multimodal_rssm/models/):
 *
 *   rssm_mrssm_rollout_fwd   MoPoE_MRSSM.rollout_representation   mrssm/mopoe_mrssm/core.py:184-260
 *                            (after the encoders, :215-216; includes Transition.forward networks.py:151-173,
 *                             the posterior heads :62-84, the MoPoE fusion :112-163,241-251, the samples
 *                             state.py:17 and the per-(b,t) KL terms of core.py:212-216)
 *                            with dims.unimodal = 1: BaseRSSM.rollout_representation  core.py:137-168 (single
 *                             Representation.forward networks.py:70-84, no fusion)
 *   rssm_mrssm_rollout_bwd   autograd (BPTT) of the above; rssm_mrssm_wgrad = its weight-gradient half
 *   rssm_mrssm_imagine_fwd   BaseRSSM.rollout_transition           core.py:170-185
 *   rssm_mtrssm_rollout_fwd  MoPoE_MMTRSSM.rollout_representation  mmtrssm/mopoe_mmtrssm/core.py:364-494
 *                            (MTRNN :40-61, lower prior :263-287, heads :241-261, fusion :436-455,
 *                             higher prior/posterior :289-319, samples :456,:464 and mmtrssm/state.py:48-49,
 *                             KL terms of :589-600)
 *   rssm_mtrssm_rollout_bwd  autograd (BPTT) of the above
 *   rssm_mtrssm_imagine_fwd  MoPoE_MMTRSSM.rollout_transition      mmtrssm/mopoe_mmtrssm/core.py:496-544
 *   rssm_gaussian_nll_fwd    likelihood (objective.py:7-23) for all modalities of compute_reconstruction_loss
 *                            (mrssm/mopoe_mrssm/core.py:294-303) in one launch; rssm_gaussian_nll_bwd = its autograd
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 (8-byte aligned); the library borrows it for the
 *     duration of the launch on `stream` (a cudaStream_t passed as void*; NULL = default stream) and keeps
 *     no state between calls (re-entrant across streams; MTRNN.hidden is an explicit input/output).
 *   - weights use PyTorch layouts ([out, in] row-major Linear / GRUCell weights), read in place.
 *   - categorical noise is caller-supplied: uniforms u in [0,1), one per (b, t, group); the draw is
 *     idx = min(K-1, #{k : cdf_k <= u}).
 *   - return value: 0 = ok, non-zero = error; rssm_last_error() returns a thread-local message.
 *   - supported sizes (anything else returns an error, never a fallback):
 *       default family: deter = hidden = 32, stochastic size C*K = 16 with K in {2,4,8,16}, embed = 64, action even, 2..8.
 *       wide family (MoPoE-MRSSM only, RSSM_PRECISION_BF16 only): deter = hidden = D in {128, 256, 384, 512}
 *       (BASELINE.json cfg3 "hidden 512"), same stochastic / embed sizes, action 1..8.  The wide family needs a caller-
 *       provided workspace and a larger saved record: size them with rssm_mrssm_workspace_bytes / rssm_mrssm_saved_bytes.
 *   - precision: RSSM_PRECISION_FP32 = every contraction as a 3-way bf16 split on the tensor cores with fp32
 *     accumulation (fp32-level accuracy); RSSM_PRECISION_BF16 = single bf16 operands, fp32 accumulation,
 *     fp32 state/epilogues.
 */
#ifndef RSSM_ROLLOUT_H_
#define RSSM_ROLLOUT_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSSM_ABI_VERSION 5
#define RSSM_PRECISION_FP32 0
#define RSSM_PRECISION_BF16 1
/* bf16 tensor-core path whose backward is ONE kernel: BPTT + weight-gradient contractions on tcgen05 with TMEM accumulators
   (MMTRSSM only).  Same forward, same records, same numerics as RSSM_PRECISION_BF16. */
#define RSSM_PRECISION_BF16_FUSED 2

/* ELEMENTS per (b,t) of the opaque records exchanged between fwd, bwd and wgrad.  Element type: fp32 with
   RSSM_PRECISION_FP32, bf16 with RSSM_PRECISION_BF16 (so bytes = elements * 4 or * 2). */
#define MRSSM_SAVED_FLOATS 320
#define MRSSM_DPRE_FLOATS 336
#define MTRSSM_SAVED_FLOATS 208 /* 192 used; bf16 rows of 416 bytes: whole 32-byte sectors, and a row pitch that lets the  */
#define MTRSSM_DPRE_FLOATS 304  /* weight-gradient kernel read bulk-copied 32-row slabs with ldmatrix in place              */
/* (the fused backward reads the same saved record as the two-kernel backward) */
#define MTRSSM_SAVED_BF16 MTRSSM_SAVED_FLOATS
#define MTRSSM_ROW_PITCH 256   /* floats per (b,t) row of the grouped outputs (RssmMtrssmOutputs.ld_*) */

/* ---- MoPoE-MRSSM -------------------------------------------------------------------------------------- */
typedef struct {
    int B, T;       /* sequences, steps */
    int A, E, D, H; /* action, obs-embed, deterministic, MLP hidden sizes */
    int C, K;       /* category_size (groups) x class_size (classes per group) */
    int precision;  /* RSSM_PRECISION_* */
    int unimodal;   /* 0 = MoPoE_MRSSM.rollout_representation.  1 = BaseRSSM.rollout_representation (core.py:137-168): the posterior
                       is Representation.forward (networks.py:70-84) of ONE head -- the au_* weights on embed_a -- with no fusion;
                       embed_v / vi_* are not used by the result (pass any valid buffers, e.g. the audio ones; their gradients
                       come back zero).  Default-size family only. */
} RssmMrssmDims;

/* state_dict names in comments */
typedef struct {
    const float *asp_w1, *asp_b1, *asp_w2, *asp_b2; /* transition.action_state_projector.{0,2}.{weight,bias}  [H,A+S] [H] [H,H] [H] */
    const float *w_ih, *w_hh, *b_ih, *b_hh;         /* transition.rnn_cell.{weight_ih,weight_hh,bias_ih,bias_hh}   [3D,H] [3D,D] [3D] [3D] */
    const float *pr_w1, *pr_b1, *pr_w2, *pr_b2;     /* transition.rnn_to_prior_projector.{0,2}.*                   [H,D] [H] [S,H] [S] */
    const float *au_w1, *au_b1, *au_w2, *au_b2;     /* audio_representation.rnn_to_post_projector.{0,2}.*          [H,D+E] [H] [S,H] [S] */
    const float *vi_w1, *vi_b1, *vi_w2, *vi_b2;     /* vision_representation.rnn_to_post_projector.{0,2}.*         [H,D+E] [H] [S,H] [S] */
} RssmMrssmWeights;

typedef struct { /* same shapes; ACCUMULATED INTO (caller zero-fills) */
    float *asp_w1, *asp_b1, *asp_w2, *asp_b2;
    float *w_ih, *w_hh, *b_ih, *b_hh;
    float *pr_w1, *pr_b1, *pr_w2, *pr_b2;
    float *au_w1, *au_b1, *au_w2, *au_b2;
    float *vi_w1, *vi_b1, *vi_w2, *vi_b2;
} RssmMrssmWeightGrads;

typedef struct {
    const float *actions;  /* [B,T,A] */
    const float *embed_a;  /* [B,T,E] audio encoder output */
    const float *embed_v;  /* [B,T,E] vision encoder output */
    const float *h0;       /* [B,D]   prev_state.deter */
    const float *z0;       /* [B,S]   prev_state.stoch */
    const float *u_post;   /* [B,T,C] uniforms for the mixed-posterior draw */
    const float *u_prior;  /* [B,T,C] uniforms for the prior State's own draw; NULL = do not draw */
} RssmMrssmInputs;

typedef struct {
    float *feature;      /* [B,T,D+S] posterior.feature = [deter | stoch]            (state.py:18) */
    float *prior_probs;  /* [B,T,C,K] prior.distribution                                          */
    float *post_probs;   /* [B,T,C,K] posterior.distribution                                      */
    float *prior_stoch;  /* [B,T,S]   prior.stoch (one-hot); may be NULL                          */
    float *kl;           /* [B,T]     sum_c KL(post_c || prior_c), before mean / balancing / coeff */
    void *saved;         /* [B,T,MRSSM_SAVED_FLOATS] record elements for the backward; NULL = inference
                            (wide family: rssm_mrssm_saved_bytes(dims) bytes, opaque) */
    void *workspace;     /* wide family only: rssm_mrssm_workspace_bytes(dims, 0) bytes of scratch, 256-byte aligned */
    size_t workspace_bytes;
} RssmMrssmOutputs;

typedef struct {           /* upstream gradients; any pointer may be NULL (= zero) except d_feature */
    const float *d_feature;     /* [B,T,D+S] */
    const float *d_prior_probs; /* [B,T,C,K] */
    const float *d_post_probs;  /* [B,T,C,K] */
    const float *d_prior_stoch; /* [B,T,S]   */
    const float *d_kl;          /* [B,T]     */
    float kl_wq, kl_wp;         /* weights of the gradient paths of `kl` into posterior / prior:
                                   (1,1) plain KL; (1-a, a) KL balancing with a = 0.8 */
} RssmMrssmUpstream;

typedef struct {
    float *d_actions; /* [B,T,A]; may be NULL */
    float *d_embed_a; /* [B,T,E] */
    float *d_embed_v; /* [B,T,E] */
    float *d_h0;      /* [B,D]   */
    float *d_z0;      /* [B,S]   */
    void *dpre;       /* [B,T,MRSSM_DPRE_FLOATS] record elements, workspace (wide family: unused, may be NULL) */
    void *workspace;  /* wide family only: rssm_mrssm_workspace_bytes(dims, 1) bytes of scratch, 256-byte aligned */
    size_t workspace_bytes;
} RssmMrssmInputGrads;

/* bytes of the opaque saved record (default family: B*T*MRSSM_SAVED_FLOATS elements) and of the scratch workspace a call
   needs (pass 0 = forward / imagination, 1 = backward; 0 bytes for the default family).  0 is also returned for unsupported
   dims (the launch entry points report the error). */
size_t rssm_mrssm_saved_bytes(const RssmMrssmDims *dims);
size_t rssm_mrssm_workspace_bytes(const RssmMrssmDims *dims, int pass);
int rssm_mrssm_rollout_fwd(const RssmMrssmDims *dims, const RssmMrssmWeights *w, const RssmMrssmInputs *in,
                           const RssmMrssmOutputs *out, void *stream);
int rssm_mrssm_rollout_bwd(const RssmMrssmDims *dims, const RssmMrssmWeights *w, const RssmMrssmInputs *in,
                           const RssmMrssmOutputs *fwd_out, const RssmMrssmUpstream *up, const RssmMrssmInputGrads *gin,
                           const RssmMrssmWeightGrads *gw, void *stream);
/* weight gradients only (the second half of rssm_mrssm_rollout_bwd, which calls it when gw != NULL):
   dW += dpre^T . layer inputs over all (b,t).  `dpre` is the workspace the backward filled. */
int rssm_mrssm_wgrad(const RssmMrssmDims *dims, const RssmMrssmInputs *in, const RssmMrssmOutputs *fwd_out, const void *dpre,
                     const RssmMrssmWeightGrads *gw, void *stream);
/* imagination: out->feature [B,T,D+S] = [deter | prior sample], out->prior_probs; in->u_prior required;
   in->embed_*, in->u_post, out->post_probs/prior_stoch/kl/saved ignored */
int rssm_mrssm_imagine_fwd(const RssmMrssmDims *dims, const RssmMrssmWeights *w, const RssmMrssmInputs *in,
                           const RssmMrssmOutputs *out, void *stream);

/* ---- MoPoE-MMTRSSM ------------------------------------------------------------------------------------- */
typedef struct {
    int B, T;
    int A, E, HD, LD, HH, HR; /* action, embed, higher/lower deter, head hidden, representation hidden */
    int CL, KL, CH, KH;       /* l_dist / h_dist: category_size x class_size */
    float l_tau, h_tau;
    int precision;
    int obs_projected; /* 0: embed_a / embed_v are the encoder outputs [B,T,E] and the modality heads' first layer multiplies them by
                          W1[:, LD:] inside the step (mopoe_mmtrssm/core.py:259-260).  1 (ABI v5; RSSM_PRECISION_BF16_FUSED only):
                          embed_a / embed_v hold the PRE-MULTIPLIED partials P = e . W1[:, LD:]^T, [B,T,HR] (no bias) -- SURVEY.md
                          §8 f2: one big GEMM before the loop, or the encoder's last Linear with the merged weight; the kernels add
                          them to the first-layer accumulators, d_embed_a / d_embed_v are then [B,T,HR] = d P, and the weight-
                          gradient columns au_w1[:, LD:] / vi_w1[:, LD:] are NOT touched (they belong to the caller's GEMM). */
} RssmMtrssmDims;

typedef struct {
    const float *l_d2h_w, *l_d2h_b, *l_in_w, *l_in_b; /* l_rnn._d2h.{weight,bias} [LD,LD] [LD]; l_rnn._input2h.* [LD,A+LS+HS] [LD] */
    const float *h_d2h_w, *h_d2h_b, *h_in_w, *h_in_b; /* h_rnn._d2h.* [HD,HD] [HD]; h_rnn._input2h.* [HD,HS] [HD] */
    const float *lp_w1, *lp_b1, *lp_w2, *lp_b2;       /* l_prior.{0,2}.*      [HH,LD] [HH] [LS,HH] [LS] */
    const float *hp_w1, *hp_b1, *hp_w2, *hp_b2;       /* h_prior.{0,2}.*      [HH,HD] [HH] [HS,HH] [HS] */
    const float *hq_w1, *hq_b1, *hq_w2, *hq_b2;       /* h_posterior.{0,2}.*  [HH,LD+HD] [HH] [HS,HH] [HS] */
    const float *au_w1, *au_b1, *au_w2, *au_b2;       /* audio_representation.rnn_to_post_projector.{0,2}.* [HR,LD+E] [HR] [LS,HR] [LS] */
    const float *vi_w1, *vi_b1, *vi_w2, *vi_b2;       /* vision_representation.rnn_to_post_projector.{0,2}.* */
} RssmMtrssmWeights;

typedef struct { /* ACCUMULATED INTO (caller zero-fills) */
    float *l_d2h_w, *l_d2h_b, *l_in_w, *l_in_b;
    float *h_d2h_w, *h_d2h_b, *h_in_w, *h_in_b;
    float *lp_w1, *lp_b1, *lp_w2, *lp_b2;
    float *hp_w1, *hp_b1, *hp_w2, *hp_b2;
    float *hq_w1, *hq_b1, *hq_w2, *hq_b2;
    float *au_w1, *au_b1, *au_w2, *au_b2;
    float *vi_w1, *vi_b1, *vi_w2, *vi_b2;
} RssmMtrssmWeightGrads;

typedef struct {
    const float *actions, *embed_a, *embed_v;  /* [B,T,A] [B,T,E] [B,T,E] */
    const float *deter_h0, *deter_l0;          /* [B,HD] [B,LD]  prev_state.deter_*  */
    const float *hidden_h0, *hidden_l0;        /* [B,HD] [B,LD]  prev_state.hidden_* (MTRNN.hidden) */
    const float *stoch_h0, *stoch_l0;          /* [B,HS] [B,LS]  prev_state.stoch_*  */
    const float *u_post_l, *u_post_h;          /* [B,T,CL] [B,T,CH] */
    const float *u_prior_l, *u_prior_h;        /* [B,T,CL] [B,T,CH]; both NULL = do not draw */
} RssmMtrssmInputs;

typedef struct {
    float *feature;                         /* [B,T,HD+HS+LD+LS] = [deter_h | stoch_h | deter_l | stoch_l] (mmtrssm/state.py:51) */
    float *hidden_h, *hidden_l;             /* [B,T,HD] [B,T,LD] MTRNN.hidden after the update */
    float *prior_probs_h, *prior_probs_l;   /* [B,T,CH,KH] [B,T,CL,KL] */
    float *post_probs_h, *post_probs_l;
    float *prior_stoch_h, *prior_stoch_l;   /* [B,T,HS] [B,T,LS]; may be NULL */
    float *kl_l, *kl_h;                     /* [B,T] each */
    void *saved;                            /* opaque record for the backward; NULL = inference.  RSSM_PRECISION_FP32 / _BF16:
                                               [B,T,MTRSSM_SAVED_FLOATS] fp32 / bf16 elements.  RSSM_PRECISION_BF16_FUSED: bf16, TILE-BLOCKED
                                               [ceil(B/16)][T][MTRSSM_SAVED_FLOATS/8][16][8] = ceil(B/16)*16 * T * MTRSSM_SAVED_FLOATS
                                               elements (size it for B rounded up to a multiple of 16; 16-byte aligned) */
    /* Row pitches in floats (ABI v5).  All 0 = DENSE outputs: every tensor above contiguous [B,T,w].  Or the GROUPED row of a (b,t):
       one [B,T,MTRSSM_ROW_PITCH] buffer
       [feature 0:96 | hidden_h 96:128 | hidden_l 128:160 | prior_probs_h 160:176 | prior_probs_l 176:192 | post_probs_h 192:208 |
       post_probs_l 208:224 | prior_stoch_h 224:240 | prior_stoch_l 240:256] -- every pointer above points INTO that buffer (16-byte
       aligned, any column order) and ld_feature = ld_hidden = ld_probs = ld_stoch = MTRSSM_ROW_PITCH -- plus one [B,T,2] buffer
       [kl_l | kl_h] with ld_kl = 2.  A row-step then writes eight whole 128-byte lines of one 1 KB row instead of ten partial-line
       segments (forward 0.56 -> 0.49 ms at the bench batch; profiles/r2_p_grouped_rows_ab.txt).  The pitches are compile-time
       constants of the kernels, so no other value is accepted.  Grouped rows are honoured by rssm_mtrssm_rollout_fwd and the
       fused rssm_mtrssm_rollout_bwd under RSSM_PRECISION_BF16_FUSED (the layout the Python ops use); every other entry point /
       policy needs 0. */
    int ld_feature, ld_hidden, ld_probs, ld_stoch, ld_kl;
} RssmMtrssmOutputs;

typedef struct {
    const float *d_feature;                          /* required */
    const float *d_prior_probs_h, *d_prior_probs_l;  /* optional (NULL = zero) ... */
    const float *d_post_probs_h, *d_post_probs_l;
    const float *d_prior_stoch_h, *d_prior_stoch_l;
    const float *d_kl_l, *d_kl_h;
    float kl_wq, kl_wp;
    const float *d_hidden_h, *d_hidden_l;            /* [B,T,HD] [B,T,LD], dense; optional: gradients into the MTRNN.hidden outputs
                                                        (the reference's autograd carries them, mmtrssm/mopoe_mmtrssm/core.py:472-473) */
} RssmMtrssmUpstream;

typedef struct {
    float *d_actions;                       /* may be NULL */
    float *d_embed_a, *d_embed_v;
    float *d_deter_h0, *d_deter_l0, *d_hidden_h0, *d_hidden_l0, *d_stoch_h0, *d_stoch_l0;
    void *dpre;                             /* [B,T,MTRSSM_DPRE_FLOATS] record elements, workspace; may be NULL when
                                               rssm_mtrssm_rollout_bwd runs fused (RSSM_PRECISION_BF16_FUSED and gw != NULL) */
} RssmMtrssmInputGrads;

int rssm_mtrssm_rollout_fwd(const RssmMtrssmDims *dims, const RssmMtrssmWeights *w, const RssmMtrssmInputs *in,
                            const RssmMtrssmOutputs *out, void *stream);
/* gw == NULL: data gradients only (fills `dpre` for a later rssm_mtrssm_wgrad).  gw != NULL: also ADDS the weight gradients
   into gw -- with RSSM_PRECISION_BF16_FUSED in ONE kernel (BPTT + weight-gradient contractions on tcgen05 / TMEM, no dpre
   round trip), otherwise as backward kernel + rssm_mtrssm_wgrad. */
int rssm_mtrssm_rollout_bwd(const RssmMtrssmDims *dims, const RssmMtrssmWeights *w, const RssmMtrssmInputs *in,
                            const RssmMtrssmOutputs *fwd_out, const RssmMtrssmUpstream *up, const RssmMtrssmInputGrads *gin,
                            const RssmMtrssmWeightGrads *gw, void *stream);
int rssm_mtrssm_wgrad(const RssmMtrssmDims *dims, const RssmMtrssmInputs *in, const RssmMtrssmOutputs *fwd_out, const void *dpre,
                      const RssmMtrssmWeightGrads *gw, void *stream);
/* imagination: feature = [deter_h | prior sample h | deter_l | prior sample l], hidden_*, prior_probs_*;
   u_prior_* required */
int rssm_mtrssm_imagine_fwd(const RssmMtrssmDims *dims, const RssmMtrssmWeights *w, const RssmMtrssmInputs *in,
                            const RssmMtrssmOutputs *out, void *stream);

/* ---- reconstruction likelihood ------------------------------------------------------------------------- */
/* Replaces the body of `likelihood` (objective.py:7-23) as `compute_reconstruction_loss` calls it once per modality
   (mrssm/mopoe_mrssm/core.py:294-303):
     loss = -mean over the batch dims of Independent(Normal(prediction, scale), event_ndims).log_prob(target)
          = 0.5 / (scale^2 * n_batch) * sum_i (target_i - prediction_i)^2 + n_event * (log(scale) + 0.5 log(2 pi)),
   n_event = n_elems / n_batch.  Up to RSSM_NLL_MAX_SEGMENTS (prediction, target) pairs -- the modalities -- per launch.
   prediction may be fp32 / bf16 / fp16 (decoder output under autocast), target and loss are fp32.  Pointers must be 16-byte
   aligned.  The result is bit-reproducible (fixed summation order for a given size). */
#define RSSM_NLL_MAX_SEGMENTS 4
#define RSSM_DTYPE_F32 0
#define RSSM_DTYPE_BF16 1
#define RSSM_DTYPE_F16 2
typedef struct {
    const void *prediction;  /* [n_elems] of pred_dtype */
    const float *target;     /* [n_elems] */
    size_t n_elems, n_batch; /* n_batch = product of the batch dims (the mean's denominator) */
    float scale;
    float *loss;             /* fwd out: device scalar */
    const float *d_loss;     /* bwd in: device scalar upstream gradient; NULL = 1 */
    void *d_prediction;      /* bwd out: [n_elems] of pred_dtype */
    float *d_target;         /* bwd out: [n_elems]; may be NULL */
} RssmNllPair;
/* bytes of zero-filled scratch rssm_gaussian_nll_fwd needs; a call leaves it zero-filled where it must be (reusable without
   a memset by later calls on the SAME stream; concurrent streams need their own) */
size_t rssm_gaussian_nll_workspace_bytes(void);
int rssm_gaussian_nll_fwd(const RssmNllPair *pairs, int n_pairs, int pred_dtype, void *workspace, size_t workspace_bytes, void *stream);
int rssm_gaussian_nll_bwd(const RssmNllPair *pairs, int n_pairs, int pred_dtype, void *stream);

/* ---- data parallelism: one-shot mean-allreduce of the gradient bucket over peer memory ----------------------------------
   SURVEY.md 8(e): the batch-sharded rollout exchanges ONE small bucket per step (the weight gradients, 66 KB at the default sizes;
   the reference: Lightning DDP's NCCL allreduce, mopoe_mmtrssm/configs/default.yaml trainer.strategy).  One process per GPU of one
   NVLink / NVSwitch box, at most RSSM_P2P_MAX_RANKS ranks.  Every rank
     1. rssm_p2p_alloc()s a region (cudaMalloc, zero-filled): [flags: 2 slots x RSSM_P2P_MAX_RANKS u32 | pad to 256 B |
        bucket slot 0: n floats | bucket slot 1: n floats], sized by rssm_p2p_region_bytes(n);
     2. rssm_p2p_export()s a 64-byte handle, exchanges the handles with its peers by any host-side means (the Python host:
        torch.distributed.all_gather_object), and rssm_p2p_import()s each peer's handle (cudaIpc; peer access is enabled lazily);
     3. after its backward calls rssm_p2p_allreduce_mean(): `src` (this rank's gradients, n floats of ordinary device memory) is
        copied into bucket slot (step & 1) of its OWN region (a 66 KB device copy on the same stream) -- or src == NULL and the
        caller has filled that slot in place -- then out[0:n] = mean over ranks of that slot's buckets, read straight from peer memory by one kernel (flag exchange with
        release / acquire at system scope, identical summation order on every rank).  `out` may be `src` (in place).
   `step` counts the calls (0, 1, 2, ...) and must advance by one per call on every rank; the slots alternate so that a rank may
   refill slot (step & 1) as soon as its call of step + 1 has completed on its stream.  A peer that never arrives makes the kernel
   give up after `timeout_ms` (the result is then not written and rssm_p2p_status() returns 1 after synchronisation). */
#define RSSM_P2P_MAX_RANKS 8
typedef struct {
    int world, rank;
    void *regions[RSSM_P2P_MAX_RANKS]; /* device pointers of every rank's region as mapped in THIS process ([rank] = own) */
    size_t n;                          /* floats per bucket */
} RssmP2pComm;
size_t rssm_p2p_region_bytes(size_t n);
int rssm_p2p_alloc(size_t bytes, void **region);
int rssm_p2p_free(void *region);
int rssm_p2p_export(void *region, unsigned char handle[64]);
int rssm_p2p_import(const unsigned char handle[64], void **peer_region);
int rssm_p2p_close(void *peer_region);
float *rssm_p2p_bucket(void *region, size_t n, int slot);   /* address of bucket `slot` inside a region */
int rssm_p2p_allreduce_mean(const RssmP2pComm *comm, long long step, const float *src, float *out, int timeout_ms, void *stream);
int rssm_p2p_status(const RssmP2pComm *comm);              /* 0 = ok, 1 = a call timed out (reads the status word; synchronises) */

/* ---- misc ------------------------------------------------------------------------------------------------ */
int rssm_abi_version(void);
const char *rssm_last_error(void);
/* number of kernels launched by this library in this process (all threads) */
long long rssm_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RSSM_ROLLOUT_H_ */
