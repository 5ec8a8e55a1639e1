// MoPoE-MMTRSSM (fast/slow multi-timescale) latent rollout: persistent forward (+ imagination) and
// fused BPTT backward kernels.
//
// Reference semantics (paths relative to /root/reference/src/multimodal_rssm/models/mmtrssm/):
//   forward  : MoPoE_MMTRSSM.rollout_representation     mopoe_mmtrssm/core.py:364-494
//              MTRNN._compute_mtrnn (leaky integrator)    mopoe_mmtrssm/core.py:40-61
//              _compute_lower_prior                       mopoe_mmtrssm/core.py:263-287
//              _compute_lower_posterior_with_logits       mopoe_mmtrssm/core.py:241-261
//              MoPoE fusion (inline)                      mopoe_mmtrssm/core.py:436-455
//              _compute_higher_prior_posterior            mopoe_mmtrssm/core.py:289-319
//              MTState.__init__ (samples, feature order)  state.py:19-51
//   imagine  : MoPoE_MMTRSSM.rollout_transition          mopoe_mmtrssm/core.py:496-544
//   backward : autograd of the above
//
// MTRNN.hidden (a mutable module attribute in the reference, :38,:59) is a functional input/output
// here.  `l_posterior` and the dummy `transition` are never evaluated by the reference rollout
// (:405-490) and have no kernel-side counterpart.
//
// Fixed sizes of this instantiation: hd = ld = 32, hs = ls = 16, head hidden 32, E = 64, A <= 8.
// feature layout (state.py:51): [deter_h 0:32 | stoch_h 32:48 | deter_l 48:80 | stoch_l 80:96].
#include <stdlib.h>

#include <type_traits>

#include "frag.cuh"
#include "kernels.h"
#include "mtrssm_common.cuh"

// -DFZ_TIMING: warp 0 of CTA 0 stamps clock64() at the phase boundaries of every forward step (see mtrssm_fused_bwd.cu)
#ifdef FZ_TIMING
#include <stdio.h>
__device__ long long fw_dbg[1024][8];
#define FW_TS(i)                                                                             \
    do {                                                                                     \
        if (lane == 0 && blockIdx.x == 0 && warp == 0 && t < 1024) fw_dbg[t][i] = clock64(); \
    } while (0)
#else
#define FW_TS(i) do {} while (0)
#endif

namespace rssm {

// hidden -> ELU -> logits for a 32-wide hidden layer whose pre-activation is already accumulated
template <int NS>
__device__ __forceinline__ void head_l2(float (&acc)[4][4], float (&logits)[2][4], const float* bias2, const uint2* w2,
                                        typename Rec<NS>::T* svA, typename Rec<NS>::T* svB, int sv_off, const Rows& r, int lane) {
    map_c<4>(acc, EluOp<NS == 1>{});
    if (svA) store_rec<4>(acc, svA + sv_off, svB + sv_off, r);
    AFrag<NS, 2> f1;
    to_afrag<NS, 2>(f1, acc);
    init_bias<2>(logits, bias2, r.t);
    gemm<NS, 2, 2>(logits, f1, w2, lane);
}

// ================================================================================================
// forward
// ================================================================================================
#ifndef RSSM_MIN_CTAS
#define RSSM_MIN_CTAS 1
#endif
template <int NS, int KL, int KH, bool IMAGINE>
__global__ void __launch_bounds__(128, NS == 1 ? RSSM_MIN_CTAS : 1) mtrssm_fwd_kernel(const MtrssmFwdArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2* W = reinterpret_cast<uint2*>(smem_raw);
    float* bias = reinterpret_cast<float*>(W + (size_t)NS * mt::FWD_TILES * 32);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int A = p.A;
    __shared__ PackTable tb;
    if (tid == 0) {
        using namespace mt;
        const int ldin = A + 32;
        tb.nblocks = tb.ntiles = 0;
        pack_add(tb, false, wblk<NS>(W, L_D2H), p.w.l_d2h_w, 32, 0, 0, 32, 32, 2, 4);
        pack_add(tb, false, wblk<NS>(W, L_IN_ZL), p.w.l_in_w, ldin, 0, A, 16, 32, 1, 4);
        pack_add(tb, false, wblk<NS>(W, L_IN_ZH), p.w.l_in_w, ldin, 0, A + 16, 16, 32, 1, 4);
        pack_add(tb, false, wblk<NS>(W, L_IN_A), p.w.l_in_w, ldin, 0, 0, A, 32, 1, 4);
        pack_add(tb, false, wblk<NS>(W, H_D2H), p.w.h_d2h_w, 32, 0, 0, 32, 32, 2, 4);
        pack_add(tb, false, wblk<NS>(W, H_IN), p.w.h_in_w, 16, 0, 0, 16, 32, 1, 4);
        pack_add(tb, false, wblk<NS>(W, LP1), p.w.lp_w1, 32, 0, 0, 32, 32, 2, 4);
        pack_add(tb, false, wblk<NS>(W, LP2), p.w.lp_w2, 32, 0, 0, 32, 16, 2, 2);
        pack_add(tb, false, wblk<NS>(W, HP1), p.w.hp_w1, 32, 0, 0, 32, 32, 2, 4);
        pack_add(tb, false, wblk<NS>(W, HP2), p.w.hp_w2, 32, 0, 0, 32, 16, 2, 2);
        if (!IMAGINE) {
            pack_add(tb, false, wblk<NS>(W, HQ1L), p.w.hq_w1, 64, 0, 0, 32, 32, 2, 4);
            pack_add(tb, false, wblk<NS>(W, HQ1H), p.w.hq_w1, 64, 0, 32, 32, 32, 2, 4);
            pack_add(tb, false, wblk<NS>(W, HQ2), p.w.hq_w2, 32, 0, 0, 32, 16, 2, 2);
            pack_add(tb, false, wblk<NS>(W, A1H), p.w.au_w1, 96, 0, 0, 32, 32, 2, 4);
            pack_add(tb, false, wblk<NS>(W, A1E), p.w.au_w1, 96, 0, 32, 64, 32, 4, 4);
            pack_add(tb, false, wblk<NS>(W, A2), p.w.au_w2, 32, 0, 0, 32, 16, 2, 2);
            pack_add(tb, false, wblk<NS>(W, V1H), p.w.vi_w1, 96, 0, 0, 32, 32, 2, 4);
            pack_add(tb, false, wblk<NS>(W, V1E), p.w.vi_w1, 96, 0, 32, 64, 32, 4, 4);
            pack_add(tb, false, wblk<NS>(W, V2), p.w.vi_w2, 32, 0, 0, 32, 16, 2, 2);
        }
    }
    {  // the biases, while thread 0 fills the table
        using namespace mt;
        for (int i = tid; i < FWD_BIAS; i += nthr) {
            float v;
            if (i < B_H) v = p.w.l_d2h_b[i] + p.w.l_in_b[i];
            else if (i < B_LP1) v = p.w.h_d2h_b[i - B_H] + p.w.h_in_b[i - B_H];
            else if (i < B_LP2) v = p.w.lp_b1[i - B_LP1];
            else if (i < B_HP1) v = p.w.lp_b2[i - B_LP2];
            else if (i < B_HP2) v = p.w.hp_b1[i - B_HP1];
            else if (i < B_HQ1) v = p.w.hp_b2[i - B_HP2];
            else if (IMAGINE) v = 0.f;
            else if (i < B_HQ2) v = p.w.hq_b1[i - B_HQ1];
            else if (i < B_A1) v = p.w.hq_b2[i - B_HQ2];
            else if (i < B_A2) v = p.w.au_b1[i - B_A1];
            else if (i < B_V1) v = p.w.au_b2[i - B_A2];
            else if (i < B_V2) v = p.w.vi_b1[i - B_V1];
            else v = p.w.vi_b2[i - B_V2];
            bias[i] = v;
        }
    }
    __syncthreads();
    pack_run<NS>(tb, tid, nthr);
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int row0 = (blockIdx.x * (nthr >> 5) + warp) * 16;
    if (row0 >= p.B) return;
    const Rows r = make_rows(row0, p.B, lane);
    const int T = p.T;
    constexpr int CL = 16 / KL, CH = 16 / KH, F = 96;
    const float keep_l = 1.f - p.inv_tau_l, keep_h = 1.f - p.inv_tau_h;
    using RT = typename Rec<NS>::T;
    RT* saved = reinterpret_cast<RT*>(p.saved);
    // per-warp double-buffered input staging (see frag.cuh, namespace stg)
    // (bf16 path only: on the fp32-parity path the 3-way split weights need the shared memory)
    constexpr bool STAGED = !IMAGINE && NS == 1;
    float* stage_base = bias + mt::FWD_BIAS + warp * 2 * stg::FLOATS;
    if (STAGED) {
        for (int i = lane; i < 2 * stg::FLOATS; i += 32) stage_base[i] = 0.f;  // action pad columns stay zero
        __syncwarp();
        stage_inputs(stage_base, p.embed_a, p.embed_v, p.actions, A, p.u_post_l, CL, p.u_post_h, CH, row0, p.B, T, 0, lane);
    }

    // carried state
    float ul[4][4], uh[4][4];
    load_c<4>(ul, p.hidden_l0 + (size_t)r.rA * 32, p.hidden_l0 + (size_t)r.rB * 32, r.t);
    load_c<4>(uh, p.hidden_h0 + (size_t)r.rA * 32, p.hidden_h0 + (size_t)r.rB * 32, r.t);
    AFrag<NS, 2> dlf, dhf;
    AFrag<NS, 1> zlf, zhf;
    {
        float c[4][4];
        load_c<4>(c, p.deter_l0 + (size_t)r.rA * 32, p.deter_l0 + (size_t)r.rB * 32, r.t);
        to_afrag<NS, 2>(dlf, c);
        load_c<4>(c, p.deter_h0 + (size_t)r.rA * 32, p.deter_h0 + (size_t)r.rB * 32, r.t);
        to_afrag<NS, 2>(dhf, c);
        float z[2][4];
        load_c<2>(z, p.stoch_l0 + (size_t)r.rA * 16, p.stoch_l0 + (size_t)r.rB * 16, r.t);
        to_afrag<NS, 1>(zlf, z);
        load_c<2>(z, p.stoch_h0 + (size_t)r.rA * 16, p.stoch_h0 + (size_t)r.rB * 16, r.t);
        to_afrag<NS, 1>(zhf, z);
    }

    for (int t = 0; t < T; ++t) {
#ifdef RSSM_EXP_STORE_L2  // timing experiment only (wrong results): all per-step stores land in an L2-resident 40 MB window
        const size_t iA = (size_t)(r.rA & 1023) * T + t, iB = (size_t)(r.rB & 1023) * T + t;
#else
        const size_t iA = (size_t)r.rA * T + t, iB = (size_t)r.rB * T + t;
#endif
        RT* svA = saved ? saved + iA * p.saved_ld : nullptr;
        RT* svB = saved ? saved + iB * p.saved_ld : nullptr;
        const float* stage = stage_base + (t & 1) * stg::FLOATS;
        FW_TS(0);
        if (STAGED) {
            cp_async_wait_all();  // this step's inputs have landed ...
            __syncwarp();         // ... for every lane, and every lane is done reading the other stage
            if (t + 1 < T)
                stage_inputs(stage_base + ((t + 1) & 1) * stg::FLOATS, p.embed_a, p.embed_v, p.actions, A, p.u_post_l, CL, p.u_post_h,
                             CH, row0, p.B, T, t + 1, lane);
        }

        FW_TS(1);
        // ---- two leaky-integrator cells (mopoe_mmtrssm/core.py:59-60), both from the PREVIOUS state ----
        {
            float pl[4][4], ph[4][4];
            init_bias<4>(pl, bias + mt::B_L, r.t);
            AFrag<NS, 1> fa;
            if (STAGED) load_a_staged_act<NS>(fa, stage + stg::ACT, r.g, r.t);
            else load_a_global<NS, 1>(fa, p.actions + iA * A, p.actions + iB * A, r.t, A);
            gemm<NS, 2, 4>(pl, dlf, wblk<NS>(W, mt::L_D2H), lane);
            gemm<NS, 1, 4>(pl, zlf, wblk<NS>(W, mt::L_IN_ZL), lane);
            gemm<NS, 1, 4>(pl, zhf, wblk<NS>(W, mt::L_IN_ZH), lane);
            gemm<NS, 1, 4>(pl, fa, wblk<NS>(W, mt::L_IN_A), lane);
            init_bias<4>(ph, bias + mt::B_H, r.t);
            gemm<NS, 2, 4>(ph, dhf, wblk<NS>(W, mt::H_D2H), lane);
            gemm<NS, 1, 4>(ph, zhf, wblk<NS>(W, mt::H_IN), lane);
            float dl[4][4], dh[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    ul[nt][j] = keep_l * ul[nt][j] + pl[nt][j] * p.inv_tau_l;
                    uh[nt][j] = keep_h * uh[nt][j] + ph[nt][j] * p.inv_tau_h;
                    dl[nt][j] = Math<NS == 1>::tanh(ul[nt][j]);
                    dh[nt][j] = Math<NS == 1>::tanh(uh[nt][j]);
                }
            store_c<4>(dh, p.feature + iA * F, p.feature + iB * F, r);
            store_c<4>(dl, p.feature + iA * F + 48, p.feature + iB * F + 48, r);
            store_c<4>(uh, p.hidden_h + iA * 32, p.hidden_h + iB * 32, r);
            store_c<4>(ul, p.hidden_l + iA * 32, p.hidden_l + iB * 32, r);
            to_afrag<NS, 2>(dlf, dl);
            to_afrag<NS, 2>(dhf, dh);
        }
        FW_TS(2);
        // ---- priors (:285-286, :311-312) ------------------------------------------------------------------
        float ppl[2][4], pph[2][4];
        {
            float acc[4][4], lg[2][4];
            init_bias<4>(acc, bias + mt::B_LP1, r.t);
            gemm<NS, 2, 4>(acc, dlf, wblk<NS>(W, mt::LP1), lane);
            head_l2<NS>(acc, lg, bias + mt::B_LP2, wblk<NS>(W, mt::LP2), svA, svB, mts::LP_HID, r, lane);
            softmax_groups<KL, NS == 1>(lg, ppl);
            store_c<2>(ppl, p.prior_probs_l + iA * 16, p.prior_probs_l + iB * 16, r);
            init_bias<4>(acc, bias + mt::B_HP1, r.t);
            gemm<NS, 2, 4>(acc, dhf, wblk<NS>(W, mt::HP1), lane);
            head_l2<NS>(acc, lg, bias + mt::B_HP2, wblk<NS>(W, mt::HP2), svA, svB, mts::HP_HID, r, lane);
            softmax_groups<KH, NS == 1>(lg, pph);
            store_c<2>(pph, p.prior_probs_h + iA * 16, p.prior_probs_h + iB * 16, r);
        }
        FW_TS(3);
        if (p.u_prior_l != nullptr) {  // prior MTState ctor draws h then l (:467-474 -> state.py:48-49)
            float zh[2][4], zl[2][4];
            sample_onehot<KH>(pph, p.u_prior_h + iA * CH, p.u_prior_h + iB * CH, zh, lane);
            sample_onehot<KL>(ppl, p.u_prior_l + iA * CL, p.u_prior_l + iB * CL, zl, lane);
            if (IMAGINE) {  // prev_state = prior_state (:542)
                store_c<2>(zh, p.feature + iA * F + 32, p.feature + iB * F + 32, r);
                store_c<2>(zl, p.feature + iA * F + 80, p.feature + iB * F + 80, r);
                to_afrag<NS, 1>(zhf, zh);
                to_afrag<NS, 1>(zlf, zl);
            } else if (p.prior_stoch_l != nullptr) {
                store_c<2>(zh, p.prior_stoch_h + iA * 16, p.prior_stoch_h + iB * 16, r);
                store_c<2>(zl, p.prior_stoch_l + iA * 16, p.prior_stoch_l + iB * 16, r);
            }
        }
        FW_TS(4);
        if constexpr (!IMAGINE) {
        // ---- lower posterior: modality heads on d_l (:422-433), MoPoE fusion (:436-455), sample (:456) ----
        float la[2][4], lv[2][4];
#ifdef RSSM_EXP_SMALL_BODY  // timing experiment only (wrong results): evaluate ONE modality head, half the code
        zero_c<2>(lv);
#pragma unroll
        for (int m = 0; m < 1; ++m) {
#else
#pragma unroll
        for (int m = 0; m < 2; ++m) {
#endif
            float acc[4][4];
            init_bias<4>(acc, bias + (m == 0 ? mt::B_A1 : mt::B_V1), r.t);
            AFrag<NS, 4> fe;
            if (STAGED) load_a_staged64<NS>(fe, stage + (m == 0 ? stg::EA : stg::EV), r.g, r.t);
            else load_a_global<NS, 4>(fe, (m == 0 ? p.embed_a : p.embed_v) + iA * 64, (m == 0 ? p.embed_a : p.embed_v) + iB * 64, r.t, 64);
            gemm<NS, 2, 4>(acc, dlf, wblk<NS>(W, m == 0 ? mt::A1H : mt::V1H), lane);
            gemm<NS, 4, 4>(acc, fe, wblk<NS>(W, m == 0 ? mt::A1E : mt::V1E), lane);
            float (&lg)[2][4] = m == 0 ? la : lv;
            head_l2<NS>(acc, lg, bias + (m == 0 ? mt::B_A2 : mt::B_V2), wblk<NS>(W, m == 0 ? mt::A2 : mt::V2), svA, svB,
                        m == 0 ? mts::A_HID : mts::V_HID, r, lane);
            if (svA) store_rec<2>(lg, svA + (m == 0 ? mts::LA : mts::LV), svB + (m == 0 ? mts::LA : mts::LV), r);
        }
        FW_TS(5);
        {
            float q[2][4], zs[2][4];
            if constexpr (NS == 1) {
                mopoe_posterior_fast<KL>(la, lv, q);  // probability-domain MoPoE (frag.cuh)
            } else {
                float lsa[2][4], lsv[2][4], mixed[2][4];
                log_softmax_flat<false>(la, lsa);
                log_softmax_flat<false>(lv, lsv);
                mopoe_mix<false>(lsa, lsv, mixed, nullptr, nullptr);
                softmax_groups<KL, false>(mixed, q);
            }
            store_c<2>(q, p.post_probs_l + iA * 16, p.post_probs_l + iB * 16, r);
            if (STAGED) sample_onehot<KL>(q, stage + stg::U0 + r.g * 8, stage + stg::U0 + (r.g + 8) * 8, zs, lane);
            else sample_onehot<KL>(q, p.u_post_l + iA * CL, p.u_post_l + iB * CL, zs, lane);
            store_c<2>(zs, p.feature + iA * F + 80, p.feature + iB * F + 80, r);
            to_afrag<NS, 1>(zlf, zs);
            float kl[2];
            kl_rows<NS == 1>(q, ppl, kl);
            if (r.t == 0) {
                if (r.vA) p.kl_l[iA] = kl[0];
                if (r.vB) p.kl_l[iB] = kl[1];
            }
        }
        FW_TS(6);
        // ---- higher posterior on [d_l ; d_h] (:315-317), sample (:464) -------------------------------------
        {
            float acc[4][4], lg[2][4], q[2][4], zs[2][4];
            init_bias<4>(acc, bias + mt::B_HQ1, r.t);
            gemm<NS, 2, 4>(acc, dlf, wblk<NS>(W, mt::HQ1L), lane);
            gemm<NS, 2, 4>(acc, dhf, wblk<NS>(W, mt::HQ1H), lane);
            head_l2<NS>(acc, lg, bias + mt::B_HQ2, wblk<NS>(W, mt::HQ2), svA, svB, mts::HQ_HID, r, lane);
            softmax_groups<KH, NS == 1>(lg, q);
            store_c<2>(q, p.post_probs_h + iA * 16, p.post_probs_h + iB * 16, r);
            if (STAGED) sample_onehot<KH>(q, stage + stg::U1 + r.g * 8, stage + stg::U1 + (r.g + 8) * 8, zs, lane);
            else sample_onehot<KH>(q, p.u_post_h + iA * CH, p.u_post_h + iB * CH, zs, lane);
            store_c<2>(zs, p.feature + iA * F + 32, p.feature + iB * F + 32, r);
            to_afrag<NS, 1>(zhf, zs);
            float kl[2];
            kl_rows<NS == 1>(q, pph, kl);
            if (r.t == 0) {
                if (r.vA) p.kl_h[iA] = kl[0];
                if (r.vB) p.kl_h[iB] = kl[1];
            }
        }
        FW_TS(7);
        }  // !IMAGINE
    }
}

// ================================================================================================
// backward
// ================================================================================================
template <int NS, int KL, int KH>
__global__ void __launch_bounds__(NS == 1 ? 256 : 128, 1) mtrssm_bwd_kernel(const MtrssmBwdArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2* W = reinterpret_cast<uint2*>(smem_raw);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int A = p.A;
    __shared__ PackTable tb;
    if (tid == 0) {
        using namespace mt;
        const int ldin = A + 32;
        tb.nblocks = tb.ntiles = 0;
        pack_add(tb, true, wblk<NS>(W, T_A2), p.w.au_w2, 32, 0, 0, 16, 32, 1, 4);
        pack_add(tb, true, wblk<NS>(W, T_V2), p.w.vi_w2, 32, 0, 0, 16, 32, 1, 4);
        pack_add(tb, true, wblk<NS>(W, T_LP2), p.w.lp_w2, 32, 0, 0, 16, 32, 1, 4);
        pack_add(tb, true, wblk<NS>(W, T_HP2), p.w.hp_w2, 32, 0, 0, 16, 32, 1, 4);
        pack_add(tb, true, wblk<NS>(W, T_HQ2), p.w.hq_w2, 32, 0, 0, 16, 32, 1, 4);
        pack_add(tb, true, wblk<NS>(W, T_A1H), p.w.au_w1, 96, 0, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblk<NS>(W, T_V1H), p.w.vi_w1, 96, 0, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblk<NS>(W, T_LP1), p.w.lp_w1, 32, 0, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblk<NS>(W, T_HP1), p.w.hp_w1, 32, 0, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblk<NS>(W, T_HQ1L), p.w.hq_w1, 64, 0, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblk<NS>(W, T_HQ1H), p.w.hq_w1, 64, 0, 32, 32, 32, 2, 4);
        pack_add(tb, true, wblk<NS>(W, T_A1E), p.w.au_w1, 96, 0, 32, 32, 64, 2, 8);
        pack_add(tb, true, wblk<NS>(W, T_V1E), p.w.vi_w1, 96, 0, 32, 32, 64, 2, 8);
        pack_add(tb, true, wblk<NS>(W, T_L_D2H), p.w.l_d2h_w, 32, 0, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblk<NS>(W, T_H_D2H), p.w.h_d2h_w, 32, 0, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblk<NS>(W, T_L_IN_ZL), p.w.l_in_w, ldin, 0, A, 32, 16, 2, 2);
        pack_add(tb, true, wblk<NS>(W, T_L_IN_ZH), p.w.l_in_w, ldin, 0, A + 16, 32, 16, 2, 2);
        pack_add(tb, true, wblk<NS>(W, T_H_IN), p.w.h_in_w, 16, 0, 0, 32, 16, 2, 2);
        pack_add(tb, true, wblk<NS>(W, T_L_IN_A), p.w.l_in_w, ldin, 0, 0, 32, A, 2, 2);
    }
    __syncthreads();
    pack_run<NS>(tb, tid, nthr);
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int row0 = (blockIdx.x * (nthr >> 5) + warp) * 16;
    if (row0 >= p.B) return;
    const Rows r = make_rows(row0, p.B, lane);
    const int T = p.T;
    constexpr int F = 96;
    const float keep_l = 1.f - p.inv_tau_l, keep_h = 1.f - p.inv_tau_h;
    using RT = typename Rec<NS>::T;
    const RT* saved = reinterpret_cast<const RT*>(p.saved);
    RT* dpre = reinterpret_cast<RT*>(p.dpre);
    // per-warp staging of this kernel's per-step inputs (bf16 path; see namespace bst)
    constexpr bool STAGED = NS == 1;
    float* st = reinterpret_cast<float*>(W + (size_t)NS * mt::BWD_TILES * 32) + warp * bst::WORDS;
    __shared__ __align__(8) uint64_t bars_all[8][bst::NBAR];
    uint64_t* bars = bars_all[warp];
    uint32_t ph_df = 0, ph_sv = 0, ph_ft = 0;  // mbarrier phase parities
    const char* sv_bytes = reinterpret_cast<const char*>(p.saved);
    if constexpr (STAGED) {
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < bst::NBAR; ++i) mbar_init(&bars[i], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        bulk_rows(st + bst::DF, bst::DF_LD, reinterpret_cast<const char*>(p.d_feature), 384, bst::DF_BYTES, row0, p.B, T, T - 1, &bars[bst::BAR_DF], lane);
        bulk_rows(st + bst::SV, bst::SV_LD, sv_bytes, (size_t)p.saved_ld * 2, bst::SV_BYTES, row0, p.B, T, T - 1, &bars[bst::BAR_SV], lane);
        bulk_rows(st + bst::FT, bst::FT_LD, reinterpret_cast<const char*>(p.feature), 384, bst::FT_BYTES, row0, p.B, T, T - 1, &bars[bst::BAR_FT], lane);
        bstage_pr(st, p, row0, T - 1, lane);
    }

    // upstream KL gradients: lane t of a quad fetches ONE of the quad's four values (0: kl_h row A, 1: kl_h row B,
    // 2: kl_l row A, 3: kl_l row B) one step ahead; the step gathers them with quad shuffles
    const float* dkl_src = (r.t < 2) ? p.d_kl_h : p.d_kl_l;
    const size_t dkl_row = (size_t)((r.t & 1) ? r.rB : r.rA) * T;
    float dkl_next = dkl_src != nullptr ? dkl_src[dkl_row + T - 1] : 0.f;
    // carried gradients (w.r.t. the state handed from step t to step t+1)
    float ddl[4][4], ddh[4][4], dul[4][4], duh[4][4], dzl[2][4], dzh[2][4];
    zero_c<4>(ddl), zero_c<4>(ddh), zero_c<4>(dul), zero_c<4>(duh), zero_c<2>(dzl), zero_c<2>(dzh);

    for (int t = T - 1; t >= 0; --t) {
        const size_t iA = (size_t)r.rA * T + t, iB = (size_t)r.rB * T + t;
        const RT* svA = saved + iA * p.saved_ld;
        const RT* svB = saved + iB * p.saved_ld;
        RT* dpA = dpre + iA * MTRSSM_DPRE_FLOATS;
        RT* dpB = dpre + iB * MTRSSM_DPRE_FLOATS;

        const float dkl_cur = dkl_next;
        if (t > 0 && dkl_src != nullptr) dkl_next = dkl_src[dkl_row + t - 1];
        float hid[4][4];  // a head's saved hidden, fetched from the staged record or from global memory
        auto load_hid = [&](int off) {
            if constexpr (STAGED) load_staged_rec<4>(hid, st + bst::SV, off, r.g, r.t);
            else load_rec<4>(hid, svA + off, svB + off, r.t);
        };
        if constexpr (STAGED) {
            mbar_wait(&bars[bst::BAR_DF], ph_df), ph_df ^= 1;  // d_feature(t) has landed
            float g4[4][4], g2[2][4];
            load_staged<4, false>(g4, st + bst::DF, bst::DF_LD, 0, r.g, r.t);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) ddh[nt][j] += g4[nt][j];
            load_staged<2, false>(g2, st + bst::DF, bst::DF_LD, 32, r.g, r.t);  // straight-through: d stoch -> d probs
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) dzh[nt][j] += g2[nt][j];
            load_staged<4, false>(g4, st + bst::DF, bst::DF_LD, 48, r.g, r.t);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) ddl[nt][j] += g4[nt][j];
            load_staged<2, false>(g2, st + bst::DF, bst::DF_LD, 80, r.g, r.t);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) dzl[nt][j] += g2[nt][j];
            __syncwarp();  // every lane is done with DF: refill it for the next (earlier) step
            if (t > 0)
                bulk_rows(st + bst::DF, bst::DF_LD, reinterpret_cast<const char*>(p.d_feature), 384, bst::DF_BYTES, row0, p.B, T, t - 1,
                          &bars[bst::BAR_DF], lane);
            cp_async_wait_all();  // the probability rows of step t ...
            __syncwarp();
            mbar_wait(&bars[bst::BAR_SV], ph_sv), ph_sv ^= 1;  // ... and its saved record have landed
        } else {
            add_global<4>(ddh, p.d_feature, iA * F, iB * F, r.t);
            add_global<2>(dzh, p.d_feature, iA * F + 32, iB * F + 32, r.t);  // straight-through: d stoch -> d probs
            add_global<4>(ddl, p.d_feature, iA * F + 48, iB * F + 48, r.t);
            add_global<2>(dzl, p.d_feature, iA * F + 80, iB * F + 80, r.t);
        }

        // ---- higher layer: posterior + prior heads ------------------------------------------------------
        {
            float q[2][4], pp[2][4], dpp[2][4];
            if constexpr (STAGED) {
                load_staged<2, true>(q, st + bst::PR, 64, 0, r.g, r.t);
                load_staged<2, true>(pp, st + bst::PR, 64, 32, r.g, r.t);
            } else {
                load_c<2>(q, p.post_probs_h + iA * 16, p.post_probs_h + iB * 16, r.t);
                load_c<2>(pp, p.prior_probs_h + iA * 16, p.prior_probs_h + iB * 16, r.t);
            }
            add_global<2>(dzh, p.d_post_probs_h, iA * 16, iB * 16, r.t);
            zero_c<2>(dpp);
            add_global<2>(dpp, p.d_prior_probs_h, iA * 16, iB * 16, r.t);
            add_global<2>(dpp, p.d_prior_stoch_h, iA * 16, iB * 16, r.t);
            if (p.d_kl_h != nullptr) {
                const float dkl[2] = {__shfl_sync(FULL, dkl_cur, (lane & ~3) + 0), __shfl_sync(FULL, dkl_cur, (lane & ~3) + 1)};
                kl_rows_bwd<NS == 1>(q, pp, dkl, p.kl_wq, p.kl_wp, dzh, dpp);
            }
            float dlg[2][4];
            AFrag<NS, 2> f1;
            softmax_groups_bwd<KH>(q, dzh, dlg);
            load_hid(mts::HQ_HID);
            head_bwd<NS>(dlg, wblk<NS>(W, mt::T_HQ2), hid, dpA, dpB, mtd::HQL, mtd::HQ1, f1, r, lane);
            gemm<NS, 2, 4>(ddl, f1, wblk<NS>(W, mt::T_HQ1L), lane);
            gemm<NS, 2, 4>(ddh, f1, wblk<NS>(W, mt::T_HQ1H), lane);
            softmax_groups_bwd<KH>(pp, dpp, dlg);
            load_hid(mts::HP_HID);
            head_bwd<NS>(dlg, wblk<NS>(W, mt::T_HP2), hid, dpA, dpB, mtd::HPL, mtd::HP1, f1, r, lane);
            gemm<NS, 2, 4>(ddh, f1, wblk<NS>(W, mt::T_HP1), lane);
        }
        // ---- lower layer: MoPoE posterior + prior head ------------------------------------------------------
        {
            float q[2][4], pp[2][4], dpp[2][4];
            if constexpr (STAGED) {
                load_staged<2, true>(q, st + bst::PR, 64, 16, r.g, r.t);
                load_staged<2, true>(pp, st + bst::PR, 64, 48, r.g, r.t);
                __syncwarp();  // every lane is done with PR: refill it for the next (earlier) step
                bstage_pr(st, p, row0, t - 1, lane);
            } else {
                load_c<2>(q, p.post_probs_l + iA * 16, p.post_probs_l + iB * 16, r.t);
                load_c<2>(pp, p.prior_probs_l + iA * 16, p.prior_probs_l + iB * 16, r.t);
            }
            add_global<2>(dzl, p.d_post_probs_l, iA * 16, iB * 16, r.t);
            zero_c<2>(dpp);
            add_global<2>(dpp, p.d_prior_probs_l, iA * 16, iB * 16, r.t);
            add_global<2>(dpp, p.d_prior_stoch_l, iA * 16, iB * 16, r.t);
            if (p.d_kl_l != nullptr) {
                const float dkl[2] = {__shfl_sync(FULL, dkl_cur, (lane & ~3) + 2), __shfl_sync(FULL, dkl_cur, (lane & ~3) + 3)};
                kl_rows_bwd<NS == 1>(q, pp, dkl, p.kl_wq, p.kl_wp, dzl, dpp);
            }
            float dla[2][4], dlv[2][4];
            {
                float dm[2][4], la[2][4], lv[2][4], lsa[2][4], lsv[2][4], mixed[2][4], ra[2][4], rv[2][4];
                softmax_groups_bwd<KL>(q, dzl, dm);
                if constexpr (STAGED) {
                    load_staged_rec<2>(la, st + bst::SV, mts::LA, r.g, r.t);
                    load_staged_rec<2>(lv, st + bst::SV, mts::LV, r.g, r.t);
                } else {
                    load_rec<2>(la, svA + mts::LA, svB + mts::LA, r.t);
                    load_rec<2>(lv, svA + mts::LV, svB + mts::LV, r.t);
                }
                log_softmax_flat<NS == 1>(la, lsa);
                log_softmax_flat<NS == 1>(lv, lsv);
                mopoe_mix<NS == 1>(lsa, lsv, mixed, ra, rv);
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        ra[nt][j] *= dm[nt][j];
                        rv[nt][j] *= dm[nt][j];
                    }
                log_softmax_flat_bwd<NS == 1>(lsa, ra, dla);
                log_softmax_flat_bwd<NS == 1>(lsv, rv, dlv);
            }
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                AFrag<NS, 2> f1;
                load_hid(m == 0 ? mts::A_HID : mts::V_HID);
                head_bwd<NS>(m == 0 ? dla : dlv, wblk<NS>(W, m == 0 ? mt::T_A2 : mt::T_V2), hid, dpA, dpB, m == 0 ? mtd::LA : mtd::LV,
                             m == 0 ? mtd::A1 : mtd::V1, f1, r, lane);
                gemm<NS, 2, 4>(ddl, f1, wblk<NS>(W, m == 0 ? mt::T_A1H : mt::T_V1H), lane);
                float de[8][4];
                zero_c<8>(de);
                gemm<NS, 2, 8>(de, f1, wblk<NS>(W, m == 0 ? mt::T_A1E : mt::T_V1E), lane);
                float* dE = m == 0 ? p.d_embed_a : p.d_embed_v;
                store_c<8>(de, dE + iA * 64, dE + iB * 64, r);
            }
            float dlg[2][4];
            AFrag<NS, 2> f1;
            softmax_groups_bwd<KL>(pp, dpp, dlg);
            load_hid(mts::LP_HID);
            if constexpr (STAGED) {
                __syncwarp();  // last read of the staged saved record: refill it
                if (t > 0)
                    bulk_rows(st + bst::SV, bst::SV_LD, sv_bytes, (size_t)p.saved_ld * 2, bst::SV_BYTES, row0, p.B, T, t - 1, &bars[bst::BAR_SV], lane);
            }
            head_bwd<NS>(dlg, wblk<NS>(W, mt::T_LP2), hid, dpA, dpB, mtd::LPL, mtd::LP1, f1, r, lane);
            gemm<NS, 2, 4>(ddl, f1, wblk<NS>(W, mt::T_LP1), lane);
        }
        // ---- the two leaky integrators: u = keep*u_prev + pre/tau, d = tanh(u) -------------------------------
        {
            float dh[4][4], dl[4][4], ph[4][4], pl[4][4];
            if constexpr (STAGED) {
                mbar_wait(&bars[bst::BAR_FT], ph_ft), ph_ft ^= 1;  // feature[0:80](t) has landed
                load_staged<4, false>(dh, st + bst::FT, bst::FT_LD, 0, r.g, r.t);
                load_staged<4, false>(dl, st + bst::FT, bst::FT_LD, 48, r.g, r.t);
                __syncwarp();
                if (t > 0)
                    bulk_rows(st + bst::FT, bst::FT_LD, reinterpret_cast<const char*>(p.feature), 384, bst::FT_BYTES, row0, p.B, T, t - 1,
                              &bars[bst::BAR_FT], lane);
            } else {
                load_c<4>(dh, p.feature + iA * F, p.feature + iB * F, r.t);
                load_c<4>(dl, p.feature + iA * F + 48, p.feature + iB * F + 48, r.t);
            }
            // upstream gradients of the hidden outputs (u_t of the two cells) join the carried d u
            add_global<4>(duh, p.d_hidden_h, iA * 32, iB * 32, r.t);
            add_global<4>(dul, p.d_hidden_l, iA * 32, iB * 32, r.t);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float gh = duh[nt][j] + ddh[nt][j] * (1.f - dh[nt][j] * dh[nt][j]);
                    const float gl = dul[nt][j] + ddl[nt][j] * (1.f - dl[nt][j] * dl[nt][j]);
                    ph[nt][j] = gh * p.inv_tau_h;
                    pl[nt][j] = gl * p.inv_tau_l;
                    duh[nt][j] = gh * keep_h;
                    dul[nt][j] = gl * keep_l;
                }
            store_rec<4>(pl, dpA + mtd::L, dpB + mtd::L, r);
            store_rec<4>(ph, dpA + mtd::H, dpB + mtd::H, r);
            AFrag<NS, 2> fl, fh;
            to_afrag<NS, 2>(fl, pl);
            to_afrag<NS, 2>(fh, ph);
            zero_c<4>(ddl), zero_c<4>(ddh), zero_c<2>(dzl), zero_c<2>(dzh);
            gemm<NS, 2, 4>(ddl, fl, wblk<NS>(W, mt::T_L_D2H), lane);
            gemm<NS, 2, 4>(ddh, fh, wblk<NS>(W, mt::T_H_D2H), lane);
            gemm<NS, 2, 2>(dzl, fl, wblk<NS>(W, mt::T_L_IN_ZL), lane);
            gemm<NS, 2, 2>(dzh, fl, wblk<NS>(W, mt::T_L_IN_ZH), lane);
            gemm<NS, 2, 2>(dzh, fh, wblk<NS>(W, mt::T_H_IN), lane);
            if (p.d_actions != nullptr) {
                float da[2][4];
                zero_c<2>(da);
                gemm<NS, 2, 2>(da, fl, wblk<NS>(W, mt::T_L_IN_A), lane);
                store_c_partial(da, p.d_actions + iA * A, p.d_actions + iB * A, r, A);
            }
        }
    }
    store_c<4>(ddh, p.d_deter_h0 + (size_t)r.rA * 32, p.d_deter_h0 + (size_t)r.rB * 32, r);
    store_c<4>(ddl, p.d_deter_l0 + (size_t)r.rA * 32, p.d_deter_l0 + (size_t)r.rB * 32, r);
    store_c<4>(duh, p.d_hidden_h0 + (size_t)r.rA * 32, p.d_hidden_h0 + (size_t)r.rB * 32, r);
    store_c<4>(dul, p.d_hidden_l0 + (size_t)r.rA * 32, p.d_hidden_l0 + (size_t)r.rB * 32, r);
    store_c<2>(dzh, p.d_stoch_h0 + (size_t)r.rA * 16, p.d_stoch_h0 + (size_t)r.rB * 16, r);
    store_c<2>(dzl, p.d_stoch_l0 + (size_t)r.rA * 16, p.d_stoch_l0 + (size_t)r.rB * 16, r);
}

// ================================================================================================
// launchers
// ================================================================================================
// warps (= 16-sequence tiles) per CTA: spread small batches over all SMs, share one weight copy per SM for large ones
static int pick_warps_per_cta(int B, int max_wpc) {
    const int warps = (B + 15) / 16;
    int wpc = 1;
    while (wpc < max_wpc && warps > 2 * wpc * 148) wpc *= 2;
    return wpc;
}

// smem = smem_cta (weights, biases) + warps per CTA * smem_warp (per-warp staging)
template <typename KernelT, typename ArgsT>
static cudaError_t launch(KernelT kernel, const ArgsT& args, int B, size_t smem_cta, size_t smem_warp, int max_wpc, cudaStream_t stream) {
    const int wpc = pick_warps_per_cta(B, max_wpc);
    size_t smem = smem_cta + wpc * smem_warp;
    if (const char* pad = getenv("RSSM_EXP_SMEM_PAD")) smem += (size_t)atoi(pad);  // occupancy experiment only
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    const int ctas = ((B + 15) / 16 + wpc - 1) / wpc;
    kernel<<<ctas, wpc * 32, smem, stream>>>(args);
#ifdef FZ_TIMING
    if (getenv("RSSM_FZ_TIMING") && std::is_same<ArgsT, MtrssmFwdArgs>::value) {
        static long long h[1024][8];
        cudaStreamSynchronize(stream);
        cudaMemcpyFromSymbol(h, fw_dbg, sizeof(h));
        const int T = args.T < 1024 ? args.T : 1024, lo = T > 8 ? 2 : 0, hi = T > 8 ? T - 3 : T - 1;
        const char* names[8] = {"top", "wait inputs+stage next", "two cells", "prior heads", "prior draws", "audio+vision heads", "fusion+draw+kl",
                                "higher posterior"};
        fprintf(stderr, "[fwd timing B=%d T=%d] warp 0, mean cycles per interval over steps %d..%d:\n", B, args.T, lo, hi);
        for (int i = 1; i < 8; ++i) {
            double sum = 0;
            for (int t = lo; t <= hi; ++t) sum += (double)(h[t][i] - h[t][i - 1]);
            fprintf(stderr, "   %-24s %8.0f\n", names[i], sum / (hi - lo + 1));
        }
        double step = 0;
        for (int t = lo + 1; t <= hi; ++t) step += (double)(h[t][0] - h[t - 1][0]);
        fprintf(stderr, "   %-24s %8.0f\n", "WHOLE STEP", step / (hi - lo));
    }
#endif
    return cudaGetLastError();
}

// supported (class_size_l, class_size_h) pairs; default.yaml is (4, 2)
#ifdef RSSM_EXP_ONLY_DEFAULT
#define MT_DISPATCH(KERNEL, ...)                                                              \
    if (a.KL == 4 && a.KH == 2) return launch(KERNEL<NS, 4, 2 __VA_ARGS__>, a, a.B, smem_cta, smem_warp, max_wpc, s); \
    return cudaErrorInvalidValue;
#else
#define MT_DISPATCH(KERNEL, ...)                                                              \
    if (a.KL == 4 && a.KH == 2) return launch(KERNEL<NS, 4, 2 __VA_ARGS__>, a, a.B, smem_cta, smem_warp, max_wpc, s); \
    if (a.KL == 4 && a.KH == 4) return launch(KERNEL<NS, 4, 4 __VA_ARGS__>, a, a.B, smem_cta, smem_warp, max_wpc, s); \
    if (a.KL == 2 && a.KH == 2) return launch(KERNEL<NS, 2, 2 __VA_ARGS__>, a, a.B, smem_cta, smem_warp, max_wpc, s); \
    if (a.KL == 8 && a.KH == 8) return launch(KERNEL<NS, 8, 8 __VA_ARGS__>, a, a.B, smem_cta, smem_warp, max_wpc, s); \
    if (a.KL == 16 && a.KH == 16) return launch(KERNEL<NS, 16, 16 __VA_ARGS__>, a, a.B, smem_cta, smem_warp, max_wpc, s); \
    return cudaErrorInvalidValue;
#endif

template <int NS, bool IMAGINE>
static cudaError_t launch_mtrssm_fwd_k(const MtrssmFwdArgs& a, cudaStream_t s) {
    // weights + biases (+ per-warp double-buffered input staging)
    const size_t smem_cta = (size_t)NS * mt::FWD_TILES * 32 * sizeof(uint2) + mt::FWD_BIAS * sizeof(float);
    const size_t smem_warp = IMAGINE || NS != 1 ? 0 : 2 * stg::FLOATS * sizeof(float);
    const int max_wpc = 4;
#define COMMA_IMAGINE , IMAGINE
    MT_DISPATCH(mtrssm_fwd_kernel, COMMA_IMAGINE)
#undef COMMA_IMAGINE
}

cudaError_t launch_mtrssm_fwd(const MtrssmFwdArgs& a, int precision, bool imagine, cudaStream_t s) {
    if (precision == RSSM_PRECISION_FP32)
        return imagine ? launch_mtrssm_fwd_k<3, true>(a, s) : launch_mtrssm_fwd_k<3, false>(a, s);
    return imagine ? launch_mtrssm_fwd_k<1, true>(a, s) : launch_mtrssm_fwd_k<1, false>(a, s);
}

template <int NS>
static cudaError_t launch_mtrssm_bwd_k(const MtrssmBwdArgs& a, cudaStream_t s) {
    const size_t smem_cta = (size_t)NS * mt::BWD_TILES * 32 * sizeof(uint2);
    const size_t smem_warp = NS == 1 ? bst::WORDS * sizeof(float) : 0;
    const int max_wpc = NS == 1 ? 8 : 4;  // bf16 path: one CTA of 8 warps per SM (255 registers, 218 KB of shared memory)
    MT_DISPATCH(mtrssm_bwd_kernel, )
}

cudaError_t launch_mtrssm_bwd(const MtrssmBwdArgs& a, int precision, cudaStream_t s) {
    return precision == RSSM_PRECISION_FP32 ? launch_mtrssm_bwd_k<3>(a, s) : launch_mtrssm_bwd_k<1>(a, s);
}

}  // namespace rssm
