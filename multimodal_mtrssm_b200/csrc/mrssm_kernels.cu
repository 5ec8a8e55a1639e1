// MoPoE-MRSSM latent rollout: persistent forward (+ imagination) and fused BPTT backward kernels.
//
// Reference semantics (paths relative to /root/reference/src/multimodal_rssm/models/):
//   forward  : MoPoE_MRSSM.rollout_representation   mrssm/mopoe_mrssm/core.py:184-260
//              Transition.forward                     networks.py:151-173   (GRUCell, gate order r,z,n)
//              _compute_posterior_with_logits         mrssm/mopoe_mrssm/core.py:62-84
//              _moe_fusion_categorical                mrssm/mopoe_mrssm/core.py:112-163
//              State.__init__ (sample, feature cat)   state.py:14-18
//   imagine  : BaseRSSM.rollout_transition            core.py:170-185
//   backward : autograd of the above (BPTT through deter AND the straight-through stoch)
//
// One warp owns 16 sequences and iterates t = 0..T-1 (backward: T-1..0) entirely in registers; see
// frag.cuh for the tile algebra.  Fixed sizes of this instantiation: D = H = 32, S = C*K = 16,
// E = 64, A <= 8 (default.yaml: 32/32/16/64/6).  The C-ABI layer rejects anything else loudly.
//
// Weight gradients are NOT accumulated here: the backward kernel writes the pre-activation
// gradients of every layer ("dpre" record, per (b,t)) and `wgrad_kernel.cu` contracts them with the
// saved layer inputs in one batch-parallel pass (the recurrence only carries data gradients).
#include "frag.cuh"
#include "kernels.h"

namespace rssm {

namespace mr {
// ---- forward weight blocks (tile offsets; a tile = 32 lanes x uint2) -------------------------
constexpr int ASP1Z = 0;             // z_prev (16)   -> asp hidden (32)   KT1 NT4
constexpr int ASP1A = ASP1Z + 4;     // action (<=16) -> asp hidden        KT1 NT4
constexpr int ASP2 = ASP1A + 4;      // asp hidden    -> x2 (32)           KT2 NT4
constexpr int IH_RZ = ASP2 + 8;      // x2 -> r,z pre (64)                 KT2 NT8
constexpr int HH_RZ = IH_RZ + 16;    // h  -> r,z pre (64)                 KT2 NT8
constexpr int IH_N = HH_RZ + 16;     // x2 -> i_n (32)                     KT2 NT4
constexpr int HH_N = IH_N + 8;       // h  -> h_n (32)                     KT2 NT4
constexpr int P1 = HH_N + 8;         // h' -> prior hidden                 KT2 NT4
constexpr int P2 = P1 + 8;           // prior hidden -> prior logits (16)  KT2 NT2
constexpr int A1H = P2 + 4;          // h' -> audio hidden                 KT2 NT4
constexpr int A1E = A1H + 8;         // embed_a (64) -> audio hidden       KT4 NT4
constexpr int A2 = A1E + 16;         // audio hidden -> audio logits       KT2 NT2
constexpr int V1H = A2 + 4;
constexpr int V1E = V1H + 8;
constexpr int V2 = V1E + 16;
constexpr int FWD_TILES = V2 + 4;    // 132
// ---- forward biases (float offsets) ------------------------------------------------------------
constexpr int B_ASP1 = 0, B_ASP2 = 32, B_RZ = 64, B_IN = 128, B_HN = 160, B_P1 = 192, B_P2 = 224, B_A1 = 240, B_A2 = 272,
              B_V1 = 288, B_V2 = 320, FWD_BIAS = 336;
// ---- backward (transposed) weight blocks -------------------------------------------------------
constexpr int T_A2 = 0;              // d audio logits (16) -> d audio hidden (32)   KT1 NT4
constexpr int T_V2 = T_A2 + 4;
constexpr int T_P2 = T_V2 + 4;
constexpr int T_A1H = T_P2 + 4;      // dpre audio (32) -> d h' (32)                 KT2 NT4
constexpr int T_V1H = T_A1H + 8;
constexpr int T_P1 = T_V1H + 8;
constexpr int T_A1E = T_P1 + 8;      // dpre audio (32) -> d embed_a (64)            KT2 NT8
constexpr int T_V1E = T_A1E + 16;
constexpr int T_IH_R = T_V1E + 16;   // dpre_r / dpre_z / dpre_n (32 each) -> d x2   KT2 NT4 x3
constexpr int T_IH_Z = T_IH_R + 8;
constexpr int T_IH_N = T_IH_Z + 8;
constexpr int T_HH_R = T_IH_N + 8;   // dpre_r / dpre_z / d h_n -> d h_prev         KT2 NT4 x3
constexpr int T_HH_Z = T_HH_R + 8;
constexpr int T_HH_N = T_HH_Z + 8;
constexpr int T_ASP2 = T_HH_N + 8;   // d x2 -> d asp hidden                         KT2 NT4
constexpr int T_ASP1Z = T_ASP2 + 8;  // dpre asp1 -> d z_prev (16)                   KT2 NT2
constexpr int T_ASP1A = T_ASP1Z + 4; // dpre asp1 -> d action (one 16-block)          KT2 NT2
constexpr int BWD_TILES = T_ASP1A + 4;  // 132
}  // namespace mr

// saved-for-backward record, floats per (b,t)   (kernels.h: MRSSM_SAVED_FLOATS = 320)
namespace mrs {
constexpr int ASP_HID = 0, X2 = 32, R = 64, Z = 96, N = 128, HN = 160, P_HID = 192, A_HID = 224, V_HID = 256, LA = 288, LV = 304;
}
// pre-activation-gradient record, floats per (b,t)   (kernels.h: MRSSM_DPRE_FLOATS = 336)
namespace mrd {
constexpr int ASP1 = 0, X2 = 32, GI = 64 /*r,z,n*/, HN = 160, P1 = 192, PL = 224, A1 = 240, LA = 272, V1 = 288, LV = 320;
}

template <int NS>
__device__ __forceinline__ uint2* wblock(uint2* W, int tile_off) {
    return W + (size_t)NS * tile_off * 32;
}

// ================================================================================================
// forward
// ================================================================================================
template <int NS, int K, bool IMAGINE>
__global__ void __launch_bounds__(128) mrssm_fwd_kernel(const MrssmFwdArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2* W = reinterpret_cast<uint2*>(smem_raw);
    float* bias = reinterpret_cast<float*>(W + (size_t)NS * mr::FWD_TILES * 32);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int A = p.A;
    __shared__ PackTable tb;
    if (tid == 0) {
        using namespace mr;
        const int ldasp = A + 16;
        tb.nblocks = tb.ntiles = 0;
        pack_add(tb, false, wblock<NS>(W, ASP1Z), p.w.asp_w1, ldasp, 0, A, 16, 32, 1, 4);
        pack_add(tb, false, wblock<NS>(W, ASP1A), p.w.asp_w1, ldasp, 0, 0, A, 32, 1, 4);
        pack_add(tb, false, wblock<NS>(W, ASP2), p.w.asp_w2, 32, 0, 0, 32, 32, 2, 4);
        pack_add(tb, false, wblock<NS>(W, IH_RZ), p.w.w_ih, 32, 0, 0, 32, 64, 2, 8);
        pack_add(tb, false, wblock<NS>(W, HH_RZ), p.w.w_hh, 32, 0, 0, 32, 64, 2, 8);
        pack_add(tb, false, wblock<NS>(W, IH_N), p.w.w_ih, 32, 64, 0, 32, 32, 2, 4);
        pack_add(tb, false, wblock<NS>(W, HH_N), p.w.w_hh, 32, 64, 0, 32, 32, 2, 4);
        pack_add(tb, false, wblock<NS>(W, P1), p.w.pr_w1, 32, 0, 0, 32, 32, 2, 4);
        pack_add(tb, false, wblock<NS>(W, P2), p.w.pr_w2, 32, 0, 0, 32, 16, 2, 2);
        if (!IMAGINE) {
            pack_add(tb, false, wblock<NS>(W, A1H), p.w.au_w1, 96, 0, 0, 32, 32, 2, 4);
            pack_add(tb, false, wblock<NS>(W, A1E), p.w.au_w1, 96, 0, 32, 64, 32, 4, 4);
            pack_add(tb, false, wblock<NS>(W, A2), p.w.au_w2, 32, 0, 0, 32, 16, 2, 2);
            pack_add(tb, false, wblock<NS>(W, V1H), p.w.vi_w1, 96, 0, 0, 32, 32, 2, 4);
            pack_add(tb, false, wblock<NS>(W, V1E), p.w.vi_w1, 96, 0, 32, 64, 32, 4, 4);
            pack_add(tb, false, wblock<NS>(W, V2), p.w.vi_w2, 32, 0, 0, 32, 16, 2, 2);
        }
    }
    {  // the biases, while thread 0 fills the table
        using namespace mr;
        for (int i = tid; i < FWD_BIAS; i += nthr) {
            float v;
            if (i < B_ASP2) v = p.w.asp_b1[i - B_ASP1];
            else if (i < B_RZ) v = p.w.asp_b2[i - B_ASP2];
            else if (i < B_IN) v = p.w.b_ih[i - B_RZ] + p.w.b_hh[i - B_RZ];
            else if (i < B_HN) v = p.w.b_ih[64 + i - B_IN];
            else if (i < B_P1) v = p.w.b_hh[64 + i - B_HN];
            else if (i < B_P2) v = p.w.pr_b1[i - B_P1];
            else if (i < B_A1) v = p.w.pr_b2[i - B_P2];
            else if (IMAGINE) v = 0.f;
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
    constexpr int C = 16 / K;
    constexpr int F = 48;
    using RT = typename Rec<NS>::T;
    RT* saved = reinterpret_cast<RT*>(p.saved);
    // per-warp double-buffered input staging (frag.cuh, namespace stg); bf16 path only (smem budget)
    constexpr bool STAGED = !IMAGINE && NS == 1;
    float* stage_base = bias + mr::FWD_BIAS + warp * 2 * stg::FLOATS;
    if (STAGED) {
        for (int i = lane; i < 2 * stg::FLOATS; i += 32) stage_base[i] = 0.f;  // action pad columns stay zero
        __syncwarp();
        stage_inputs(stage_base, p.embed_a, p.embed_v, p.actions, A, p.u_post, C, nullptr, 0, row0, p.B, T, 0, lane);
    }

    // carried state: h (fp32 C tiles + A operand), z_prev (A operand)
    float h[4][4];
    load_c<4>(h, p.h0 + (size_t)r.rA * 32, p.h0 + (size_t)r.rB * 32, r.t);
    AFrag<NS, 2> hf;
    to_afrag<NS, 2>(hf, h);
    AFrag<NS, 1> zf;
    {
        float z0[2][4];
        load_c<2>(z0, p.z0 + (size_t)r.rA * 16, p.z0 + (size_t)r.rB * 16, r.t);
        to_afrag<NS, 1>(zf, z0);
    }

    for (int t = 0; t < T; ++t) {
        const size_t iA = (size_t)r.rA * T + t, iB = (size_t)r.rB * T + t;
        RT* svA = saved ? saved + iA * MRSSM_SAVED_FLOATS : nullptr;
        RT* svB = saved ? saved + iB * MRSSM_SAVED_FLOATS : nullptr;
        const float* stage = stage_base + (t & 1) * stg::FLOATS;
        if (STAGED) {
            cp_async_wait_all();  // this step's inputs have landed ...
            __syncwarp();         // ... for every lane, and every lane is done reading the other stage
            if (t + 1 < T)
                stage_inputs(stage_base + ((t + 1) & 1) * stg::FLOATS, p.embed_a, p.embed_v, p.actions, A, p.u_post, C, nullptr, 0, row0,
                             p.B, T, t + 1, lane);
        }

        // ---- action_state_projector (networks.py:168-169) ----------------------------------------
        AFrag<NS, 2> fx;
        {
            float acc[4][4];
            init_bias<4>(acc, bias + mr::B_ASP1, r.t);
            AFrag<NS, 1> fa;
            if (STAGED) load_a_staged_act<NS>(fa, stage + stg::ACT, r.g, r.t);
            else load_a_global<NS, 1>(fa, p.actions + iA * A, p.actions + iB * A, r.t, A);
            gemm<NS, 1, 4>(acc, zf, wblock<NS>(W, mr::ASP1Z), lane);
            gemm<NS, 1, 4>(acc, fa, wblock<NS>(W, mr::ASP1A), lane);
            map_c<4>(acc, EluOp<NS == 1>{});
            if (svA) store_rec<4>(acc, svA + mrs::ASP_HID, svB + mrs::ASP_HID, r);
            AFrag<NS, 2> f1;
            to_afrag<NS, 2>(f1, acc);
            float x2[4][4];
            init_bias<4>(x2, bias + mr::B_ASP2, r.t);
            gemm<NS, 2, 4>(x2, f1, wblock<NS>(W, mr::ASP2), lane);
            if (svA) store_rec<4>(x2, svA + mrs::X2, svB + mrs::X2, r);
            to_afrag<NS, 2>(fx, x2);
        }
        // ---- GRUCell (networks.py:170) -----------------------------------------------------------
        {
            float grz[8][4], gin[4][4], ghn[4][4];
            init_bias<8>(grz, bias + mr::B_RZ, r.t);
            gemm<NS, 2, 8>(grz, fx, wblock<NS>(W, mr::IH_RZ), lane);
            gemm<NS, 2, 8>(grz, hf, wblock<NS>(W, mr::HH_RZ), lane);
            init_bias<4>(gin, bias + mr::B_IN, r.t);
            gemm<NS, 2, 4>(gin, fx, wblock<NS>(W, mr::IH_N), lane);
            init_bias<4>(ghn, bias + mr::B_HN, r.t);
            gemm<NS, 2, 4>(ghn, hf, wblock<NS>(W, mr::HH_N), lane);
            float rg[4][4], zg[4][4], ng[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    rg[nt][j] = Math<NS == 1>::sigmoid(grz[nt][j]);
                    zg[nt][j] = Math<NS == 1>::sigmoid(grz[4 + nt][j]);
                    ng[nt][j] = Math<NS == 1>::tanh(gin[nt][j] + rg[nt][j] * ghn[nt][j]);
                    h[nt][j] = (h[nt][j] - ng[nt][j]) * zg[nt][j] + ng[nt][j];
                }
            if (svA) {
                store_rec<4>(rg, svA + mrs::R, svB + mrs::R, r);
                store_rec<4>(zg, svA + mrs::Z, svB + mrs::Z, r);
                store_rec<4>(ng, svA + mrs::N, svB + mrs::N, r);
                store_rec<4>(ghn, svA + mrs::HN, svB + mrs::HN, r);
            }
            store_c<4>(h, p.feature + iA * F, p.feature + iB * F, r);
            to_afrag<NS, 2>(hf, h);
        }
        // ---- prior head + factory (networks.py:171-172) ------------------------------------------
        float pp[2][4];
        {
            float acc[4][4];
            init_bias<4>(acc, bias + mr::B_P1, r.t);
            gemm<NS, 2, 4>(acc, hf, wblock<NS>(W, mr::P1), lane);
            map_c<4>(acc, EluOp<NS == 1>{});
            if (svA) store_rec<4>(acc, svA + mrs::P_HID, svB + mrs::P_HID, r);
            AFrag<NS, 2> f1;
            to_afrag<NS, 2>(f1, acc);
            float lp[2][4];
            init_bias<2>(lp, bias + mr::B_P2, r.t);
            gemm<NS, 2, 2>(lp, f1, wblock<NS>(W, mr::P2), lane);
            softmax_groups<K, NS == 1>(lp, pp);
            store_c<2>(pp, p.prior_probs + iA * 16, p.prior_probs + iB * 16, r);
        }
        if (p.u_prior != nullptr) {  // State(prior) draws its own sample (networks.py:173 -> state.py:17)
            float zs[2][4];
            sample_onehot<K>(pp, p.u_prior + iA * C, p.u_prior + iB * C, zs, lane);
            if (IMAGINE) {
                store_c<2>(zs, p.feature + iA * F + 32, p.feature + iB * F + 32, r);
                to_afrag<NS, 1>(zf, zs);  // the prior's own sample is fed back (core.py:182-183)
            } else if (p.prior_stoch != nullptr) {
                store_c<2>(zs, p.prior_stoch + iA * 16, p.prior_stoch + iB * 16, r);
            }
        }
        if constexpr (!IMAGINE) {
        // ---- per-modality posterior heads (mopoe_mrssm/core.py:80-81) ----------------------------
        float la[2][4], lv[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const float* emb = m == 0 ? p.embed_a : p.embed_v;
            float acc[4][4];
            init_bias<4>(acc, bias + (m == 0 ? mr::B_A1 : mr::B_V1), r.t);
            AFrag<NS, 4> fe;
            if (STAGED) load_a_staged64<NS>(fe, stage + (m == 0 ? stg::EA : stg::EV), r.g, r.t);
            else load_a_global<NS, 4>(fe, emb + iA * 64, emb + iB * 64, r.t, 64);
            gemm<NS, 2, 4>(acc, hf, wblock<NS>(W, m == 0 ? mr::A1H : mr::V1H), lane);
            gemm<NS, 4, 4>(acc, fe, wblock<NS>(W, m == 0 ? mr::A1E : mr::V1E), lane);
            map_c<4>(acc, EluOp<NS == 1>{});
            if (svA) store_rec<4>(acc, svA + (m == 0 ? mrs::A_HID : mrs::V_HID), svB + (m == 0 ? mrs::A_HID : mrs::V_HID), r);
            AFrag<NS, 2> f1;
            to_afrag<NS, 2>(f1, acc);
            float (&lg)[2][4] = m == 0 ? la : lv;
            init_bias<2>(lg, bias + (m == 0 ? mr::B_A2 : mr::B_V2), r.t);
            gemm<NS, 2, 2>(lg, f1, wblock<NS>(W, m == 0 ? mr::A2 : mr::V2), lane);
            if (svA) store_rec<2>(lg, svA + (m == 0 ? mrs::LA : mrs::LV), svB + (m == 0 ? mrs::LA : mrs::LV), r);
        }
        // ---- MoPoE fusion, factory, sample, KL (mopoe_mrssm/core.py:241-251,135-163) --------------
        {
            float lsa[2][4], lsv[2][4], mixed[2][4], q[2][4], zs[2][4];
            if (p.unimodal) {  // Representation.forward: factory(logits of the one head), networks.py:83
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) mixed[nt][j] = la[nt][j];
            } else if constexpr (NS != 1) {
                log_softmax_flat<false>(la, lsa);
                log_softmax_flat<false>(lv, lsv);
                mopoe_mix<false>(lsa, lsv, mixed, nullptr, nullptr);
            }
            if (NS == 1 && !p.unimodal) mopoe_posterior_fast<K>(la, lv, q);  // probability-domain MoPoE (frag.cuh)
            else softmax_groups<K, NS == 1>(mixed, q);
            store_c<2>(q, p.post_probs + iA * 16, p.post_probs + iB * 16, r);
            if (STAGED) sample_onehot<K>(q, stage + stg::U0 + r.g * 8, stage + stg::U0 + (r.g + 8) * 8, zs, lane);
            else sample_onehot<K>(q, p.u_post + iA * C, p.u_post + iB * C, zs, lane);
            store_c<2>(zs, p.feature + iA * F + 32, p.feature + iB * F + 32, r);
            to_afrag<NS, 1>(zf, zs);  // prev_state = mixed_posterior (mopoe_mrssm/core.py:256)
            float kl[2];
            kl_rows<NS == 1>(q, pp, kl);
            if (r.t == 0) {
                if (r.vA) p.kl[iA] = kl[0];
                if (r.vB) p.kl[iB] = kl[1];
            }
        }
        }  // !IMAGINE
    }
}

// ================================================================================================
// backward (BPTT, data gradients + dpre records)
// ================================================================================================
template <int NS, int K>
__global__ void __launch_bounds__(128) mrssm_bwd_kernel(const MrssmBwdArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2* W = reinterpret_cast<uint2*>(smem_raw);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int A = p.A;
    __shared__ PackTable tb;
    if (tid == 0) {
        using namespace mr;
        const int ldasp = A + 16;
        tb.nblocks = tb.ntiles = 0;
        pack_add(tb, true, wblock<NS>(W, T_A2), p.w.au_w2, 32, 0, 0, 16, 32, 1, 4);
        pack_add(tb, true, wblock<NS>(W, T_V2), p.w.vi_w2, 32, 0, 0, 16, 32, 1, 4);
        pack_add(tb, true, wblock<NS>(W, T_P2), p.w.pr_w2, 32, 0, 0, 16, 32, 1, 4);
        pack_add(tb, true, wblock<NS>(W, T_A1H), p.w.au_w1, 96, 0, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblock<NS>(W, T_V1H), p.w.vi_w1, 96, 0, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblock<NS>(W, T_P1), p.w.pr_w1, 32, 0, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblock<NS>(W, T_A1E), p.w.au_w1, 96, 0, 32, 32, 64, 2, 8);
        pack_add(tb, true, wblock<NS>(W, T_V1E), p.w.vi_w1, 96, 0, 32, 32, 64, 2, 8);
        pack_add(tb, true, wblock<NS>(W, T_IH_R), p.w.w_ih, 32, 0, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblock<NS>(W, T_IH_Z), p.w.w_ih, 32, 32, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblock<NS>(W, T_IH_N), p.w.w_ih, 32, 64, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblock<NS>(W, T_HH_R), p.w.w_hh, 32, 0, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblock<NS>(W, T_HH_Z), p.w.w_hh, 32, 32, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblock<NS>(W, T_HH_N), p.w.w_hh, 32, 64, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblock<NS>(W, T_ASP2), p.w.asp_w2, 32, 0, 0, 32, 32, 2, 4);
        pack_add(tb, true, wblock<NS>(W, T_ASP1Z), p.w.asp_w1, ldasp, 0, A, 32, 16, 2, 2);
        pack_add(tb, true, wblock<NS>(W, T_ASP1A), p.w.asp_w1, ldasp, 0, 0, 32, A, 2, 2);
    }
    __syncthreads();
    pack_run<NS>(tb, tid, nthr);
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int row0 = (blockIdx.x * (nthr >> 5) + warp) * 16;
    if (row0 >= p.B) return;
    const Rows r = make_rows(row0, p.B, lane);
    const int T = p.T;
    constexpr int F = 48;
    using RT = typename Rec<NS>::T;
    const RT* saved = reinterpret_cast<const RT*>(p.saved);
    RT* dpre = reinterpret_cast<RT*>(p.dpre);

    float dh[4][4], dz[2][4];  // carried: d loss / d deter[t], d loss / d post_stoch[t] from step t+1
    zero_c<4>(dh);
    zero_c<2>(dz);

    for (int t = T - 1; t >= 0; --t) {
        const size_t iA = (size_t)r.rA * T + t, iB = (size_t)r.rB * T + t;
        const RT* svA = saved + iA * MRSSM_SAVED_FLOATS;
        const RT* svB = saved + iB * MRSSM_SAVED_FLOATS;
        RT* dpA = dpre + iA * MRSSM_DPRE_FLOATS;
        RT* dpB = dpre + iB * MRSSM_DPRE_FLOATS;

        // upstream gradient on feature = [deter | post_stoch]
        {
            float g[4][4];
            load_c<4>(g, p.d_feature + iA * F, p.d_feature + iB * F, r.t);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) dh[nt][j] += g[nt][j];
        }
        float q[2][4], pp[2][4], dq[2][4], dpp[2][4];
        load_c<2>(q, p.post_probs + iA * 16, p.post_probs + iB * 16, r.t);
        load_c<2>(pp, p.prior_probs + iA * 16, p.prior_probs + iB * 16, r.t);
        {
            float g[2][4];
            load_c<2>(g, p.d_feature + iA * F + 32, p.d_feature + iB * F + 32, r.t);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) dq[nt][j] = g[nt][j] + dz[nt][j];  // straight-through: d stoch -> d probs
        }
        if (p.d_post_probs != nullptr) {
            float g[2][4];
            load_c<2>(g, p.d_post_probs + iA * 16, p.d_post_probs + iB * 16, r.t);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) dq[nt][j] += g[nt][j];
        }
        zero_c<2>(dpp);
        if (p.d_prior_probs != nullptr) load_c<2>(dpp, p.d_prior_probs + iA * 16, p.d_prior_probs + iB * 16, r.t);
        if (p.d_prior_stoch != nullptr) {
            float g[2][4];
            load_c<2>(g, p.d_prior_stoch + iA * 16, p.d_prior_stoch + iB * 16, r.t);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) dpp[nt][j] += g[nt][j];
        }
        if (p.d_kl != nullptr) {
            const float dkl[2] = {p.d_kl[iA], p.d_kl[iB]};
            kl_rows_bwd<NS == 1>(q, pp, dkl, p.kl_wq, p.kl_wp, dq, dpp);
        }

        // ---- posterior: factory softmax -> MoPoE mix -> flat log-softmaxes -> heads -----------------
        float dla[2][4], dlv[2][4];
        {
            float dm[2][4];
            softmax_groups_bwd<K>(q, dq, dm);
            float la[2][4], lv[2][4], lsa[2][4], lsv[2][4], mixed[2][4], ra[2][4], rv[2][4];
            load_rec<2>(la, svA + mrs::LA, svB + mrs::LA, r.t);
            load_rec<2>(lv, svA + mrs::LV, svB + mrs::LV, r.t);
            if (p.unimodal) {  // posterior logits ARE the one head's logits: the vision head gets no gradient
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dla[nt][j] = dm[nt][j], dlv[nt][j] = 0.f;
            } else {
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
        }
        store_rec<2>(dla, dpA + mrd::LA, dpB + mrd::LA, r);
        store_rec<2>(dlv, dpA + mrd::LV, dpB + mrd::LV, r);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            float (&dl)[2][4] = m == 0 ? dla : dlv;
            AFrag<NS, 1> fl;
            to_afrag<NS, 1>(fl, dl);
            float dhid[4][4], hid[4][4];
            zero_c<4>(dhid);
            gemm<NS, 1, 4>(dhid, fl, wblock<NS>(W, m == 0 ? mr::T_A2 : mr::T_V2), lane);
            load_rec<4>(hid, svA + (m == 0 ? mrs::A_HID : mrs::V_HID), svB + (m == 0 ? mrs::A_HID : mrs::V_HID), r.t);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) dhid[nt][j] *= elu_grad_from_out(hid[nt][j]);
            store_rec<4>(dhid, dpA + (m == 0 ? mrd::A1 : mrd::V1), dpB + (m == 0 ? mrd::A1 : mrd::V1), r);
            AFrag<NS, 2> f1;
            to_afrag<NS, 2>(f1, dhid);
            gemm<NS, 2, 4>(dh, f1, wblock<NS>(W, m == 0 ? mr::T_A1H : mr::T_V1H), lane);
            float de[8][4];
            zero_c<8>(de);
            gemm<NS, 2, 8>(de, f1, wblock<NS>(W, m == 0 ? mr::T_A1E : mr::T_V1E), lane);
            float* dE = m == 0 ? p.d_embed_a : p.d_embed_v;
            store_c<8>(de, dE + iA * 64, dE + iB * 64, r);
        }
        // ---- prior head ------------------------------------------------------------------------------
        {
            float dlp[2][4];
            softmax_groups_bwd<K>(pp, dpp, dlp);
            store_rec<2>(dlp, dpA + mrd::PL, dpB + mrd::PL, r);
            AFrag<NS, 1> fl;
            to_afrag<NS, 1>(fl, dlp);
            float dhid[4][4], hid[4][4];
            zero_c<4>(dhid);
            gemm<NS, 1, 4>(dhid, fl, wblock<NS>(W, mr::T_P2), lane);
            load_rec<4>(hid, svA + mrs::P_HID, svB + mrs::P_HID, r.t);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) dhid[nt][j] *= elu_grad_from_out(hid[nt][j]);
            store_rec<4>(dhid, dpA + mrd::P1, dpB + mrd::P1, r);
            AFrag<NS, 2> f1;
            to_afrag<NS, 2>(f1, dhid);
            gemm<NS, 2, 4>(dh, f1, wblock<NS>(W, mr::T_P1), lane);
        }
        // ---- GRUCell -----------------------------------------------------------------------------------
        float dx2[4][4];
        {
            float rg[4][4], zg[4][4], ng[4][4], hn[4][4], hp[4][4];
            load_rec<4>(rg, svA + mrs::R, svB + mrs::R, r.t);
            load_rec<4>(zg, svA + mrs::Z, svB + mrs::Z, r.t);
            load_rec<4>(ng, svA + mrs::N, svB + mrs::N, r.t);
            load_rec<4>(hn, svA + mrs::HN, svB + mrs::HN, r.t);
            if (t > 0) load_c<4>(hp, p.feature + (iA - 1) * F, p.feature + (iB - 1) * F, r.t);
            else load_c<4>(hp, p.h0 + (size_t)r.rA * 32, p.h0 + (size_t)r.rB * 32, r.t);
            float dpr[4][4], dpz[4][4], dpn[4][4], dhn[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float g = dh[nt][j], z = zg[nt][j], n = ng[nt][j], rr = rg[nt][j];
                    const float dn = g * (1.f - z);
                    const float dzg = g * (hp[nt][j] - n);
                    const float dn_pre = dn * (1.f - n * n);
                    dpn[nt][j] = dn_pre;
                    dhn[nt][j] = dn_pre * rr;
                    dpr[nt][j] = dn_pre * hn[nt][j] * rr * (1.f - rr);
                    dpz[nt][j] = dzg * z * (1.f - z);
                    dh[nt][j] = g * z;  // direct path h' <- h_prev
                }
            store_rec<4>(dpr, dpA + mrd::GI, dpB + mrd::GI, r);
            store_rec<4>(dpz, dpA + mrd::GI + 32, dpB + mrd::GI + 32, r);
            store_rec<4>(dpn, dpA + mrd::GI + 64, dpB + mrd::GI + 64, r);
            store_rec<4>(dhn, dpA + mrd::HN, dpB + mrd::HN, r);
            AFrag<NS, 2> fr, fz, fn;
            to_afrag<NS, 2>(fr, dpr);
            to_afrag<NS, 2>(fz, dpz);
            to_afrag<NS, 2>(fn, dpn);
            zero_c<4>(dx2);
            gemm<NS, 2, 4>(dx2, fr, wblock<NS>(W, mr::T_IH_R), lane);
            gemm<NS, 2, 4>(dx2, fz, wblock<NS>(W, mr::T_IH_Z), lane);
            gemm<NS, 2, 4>(dx2, fn, wblock<NS>(W, mr::T_IH_N), lane);
            gemm<NS, 2, 4>(dh, fr, wblock<NS>(W, mr::T_HH_R), lane);
            gemm<NS, 2, 4>(dh, fz, wblock<NS>(W, mr::T_HH_Z), lane);
            to_afrag<NS, 2>(fn, dhn);
            gemm<NS, 2, 4>(dh, fn, wblock<NS>(W, mr::T_HH_N), lane);
        }
        // ---- action_state_projector ------------------------------------------------------------------
        {
            store_rec<4>(dx2, dpA + mrd::X2, dpB + mrd::X2, r);
            AFrag<NS, 2> f2;
            to_afrag<NS, 2>(f2, dx2);
            float dhid[4][4], hid[4][4];
            zero_c<4>(dhid);
            gemm<NS, 2, 4>(dhid, f2, wblock<NS>(W, mr::T_ASP2), lane);
            load_rec<4>(hid, svA + mrs::ASP_HID, svB + mrs::ASP_HID, r.t);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) dhid[nt][j] *= elu_grad_from_out(hid[nt][j]);
            store_rec<4>(dhid, dpA + mrd::ASP1, dpB + mrd::ASP1, r);
            AFrag<NS, 2> f1;
            to_afrag<NS, 2>(f1, dhid);
            zero_c<2>(dz);
            gemm<NS, 2, 2>(dz, f1, wblock<NS>(W, mr::T_ASP1Z), lane);
            if (p.d_actions != nullptr) {
                float da[2][4];
                zero_c<2>(da);
                gemm<NS, 2, 2>(da, f1, wblock<NS>(W, mr::T_ASP1A), lane);
                store_c_partial(da, p.d_actions + iA * A, p.d_actions + iB * A, r, A);
            }
        }
    }
    store_c<4>(dh, p.d_h0 + (size_t)r.rA * 32, p.d_h0 + (size_t)r.rB * 32, r);
    store_c<2>(dz, p.d_z0 + (size_t)r.rA * 16, p.d_z0 + (size_t)r.rB * 16, r);
}

// ================================================================================================
// launchers
// ================================================================================================
static int pick_warps_per_cta(int B) {
    const int warps = (B + 15) / 16;
    if (warps <= 2 * 148) return 1;  // small batches: spread single-warp CTAs over the 148 SMs
    if (warps <= 4 * 148) return 2;
    return 4;
}

template <typename KernelT, typename ArgsT>
static cudaError_t launch(KernelT kernel, const ArgsT& args, int B, size_t smem, cudaStream_t stream) {
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    const int wpc = pick_warps_per_cta(B);
    const int ctas = ((B + 15) / 16 + wpc - 1) / wpc;
    kernel<<<ctas, wpc * 32, smem, stream>>>(args);
    return cudaGetLastError();
}

template <int NS, bool IMAGINE>
static cudaError_t launch_mrssm_fwd_k(const MrssmFwdArgs& a, cudaStream_t s) {
    const size_t smem = (size_t)NS * mr::FWD_TILES * 32 * sizeof(uint2) + mr::FWD_BIAS * sizeof(float) +
                        (IMAGINE || NS != 1 ? 0 : 4 * 2 * stg::FLOATS * sizeof(float));
    switch (a.K) {
        case 2: return launch(mrssm_fwd_kernel<NS, 2, IMAGINE>, a, a.B, smem, s);
        case 4: return launch(mrssm_fwd_kernel<NS, 4, IMAGINE>, a, a.B, smem, s);
        case 8: return launch(mrssm_fwd_kernel<NS, 8, IMAGINE>, a, a.B, smem, s);
        case 16: return launch(mrssm_fwd_kernel<NS, 16, IMAGINE>, a, a.B, smem, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_mrssm_fwd(const MrssmFwdArgs& a, int precision, bool imagine, cudaStream_t s) {
    if (precision == RSSM_PRECISION_FP32)
        return imagine ? launch_mrssm_fwd_k<3, true>(a, s) : launch_mrssm_fwd_k<3, false>(a, s);
    return imagine ? launch_mrssm_fwd_k<1, true>(a, s) : launch_mrssm_fwd_k<1, false>(a, s);
}

template <int NS>
static cudaError_t launch_mrssm_bwd_k(const MrssmBwdArgs& a, cudaStream_t s) {
    const size_t smem = (size_t)NS * mr::BWD_TILES * 32 * sizeof(uint2);
    switch (a.K) {
        case 2: return launch(mrssm_bwd_kernel<NS, 2>, a, a.B, smem, s);
        case 4: return launch(mrssm_bwd_kernel<NS, 4>, a, a.B, smem, s);
        case 8: return launch(mrssm_bwd_kernel<NS, 8>, a, a.B, smem, s);
        case 16: return launch(mrssm_bwd_kernel<NS, 16>, a, a.B, smem, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_mrssm_bwd(const MrssmBwdArgs& a, int precision, cudaStream_t s) {
    return precision == RSSM_PRECISION_FP32 ? launch_mrssm_bwd_k<3>(a, s) : launch_mrssm_bwd_k<1>(a, s);
}

}  // namespace rssm
