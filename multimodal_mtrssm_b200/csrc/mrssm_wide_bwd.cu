// Wide MoPoE-MRSSM rollout, backward (BPTT): one persistent cooperative kernel walks the T steps in reverse and leaves every
// layer's pre-activation gradient in packed bf16 planes; the weight gradients (contractions over all (b,t) rows) and the
// embedding gradients are non-recurrent and run afterwards (mrssm_wide_wgrad.cu).  autograd twin of mrssm_wide_fwd.cu; same
// CTA = (batch block bb, slice s) decomposition, operand layout and barrier scheme (wide_common.cuh).
//
// Every epilogue input (per-row statistics, gate record, head hiddens, h_{t-1}, upstream d h_t) is fetched into shared memory
// by bulk copies / cp.async WHILE the phase's contraction runs; the carried d h lives in shared memory for the whole kernel.
// One step t (reverse) = four phases separated by grid barriers:
//   P1  d[z_t ; a_{t+1}] = dhid1_{t+1} . W1                      (every slice CTA of the block, redundantly)   N = 32, K = D
//       row-wise (thread = row): the carried d z_t through straight-through sample, per-group softmax, MoPoE fusion and flat
//       log-softmax backward (linear in d z_t; everything else of that backward is precomputed for all (b,t) by the pre-pass
//       wide_bwd_rowstat_kernel) -> d logits; d head hidden = d logits . W2 (CUDA cores), * ELU' -> planes DPH / DAH / DVH [slice]
//   P2  d h_t = [DPH | DAH | DVH] . [W_prior | W_audio | W_vision][:, slice] + carry + upstream                  N = 32, K = 3D
//       GRU gate backward -> planes DG_R, DG_Z, DGI_N, DGH_N [slice]; carry = d h_t * z
//   P3  d x2 = [DG_R | DG_Z | DGI_N] . W_ih[:, slice] ; carry += [DG_R | DG_Z | DGH_N] . W_hh[:, slice]          N = 64 / 32, K = 3D
//   P4  d hid1_t = (d x2 . W2[:, slice]) * ELU'                                                                  N = 32, K = D
#include "kernels.h"
#include "wide_common.cuh"

namespace rssm {
namespace wide {

constexpr int BSTAGES = 4;
constexpr int BB_MAX_BYTES = 64 * 64 * 2;
constexpr int BSTAGE_BYTES = A_BYTES + BB_MAX_BYTES;
constexpr int TB_DX = 0, TB_DH = 32, TB_P3 = 64, TB_H1 = 128, BWD_ACC_OFF = 160;  // second issuer's accumulator copies: + BWD_ACC_OFF
constexpr int STAT_BYTES = 32 * BM * 16;            // per-row statistics of one (t, block): [32 float4 columns][128 rows]
constexpr int PIECE_BYTES = BM * 16;                // 8 features x 128 rows of a packed plane
constexpr int STAGE_REGION = STAT_BYTES + 12 * PIECE_BYTES;  // P1: stat + 12 head-hidden pieces; P2: 16 gate pieces + hprev + dfe
constexpr int ROWF = 36;                            // padded fp32 row (32 values) in shared memory: conflict-free float4 access
constexpr int P2_HPREV_OFF = 16 * PIECE_BYTES, P2_DFE_OFF = P2_HPREV_OFF + BM * ROWF * 4;
static_assert(P2_DFE_OFF + BM * ROWF * 4 <= STAGE_REGION, "staging region too small");

struct BwdSmem {
    unsigned char *ring, *stage;
    float *carry, *w2l;
    uint64_t *full, *empty, *accbar, *stgbar, *firstbar;
    uint32_t* tmem_base;
};
__host__ __device__ inline size_t bwd_smem_bytes() {
    return 128 + (size_t)BSTAGES * BSTAGE_BYTES + STAGE_REGION + BM * ROWF * 4 + 3 * 32 * 16 * 4 + (2 * BSTAGES + 3) * 8 + 16;
}
__device__ __forceinline__ BwdSmem carve_bwd(unsigned char* dyn) {
    BwdSmem s;
    unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn) + 127) & ~(uintptr_t)127);
    s.ring = p, p += (size_t)BSTAGES * BSTAGE_BYTES;
    s.stage = p, p += STAGE_REGION;
    s.carry = reinterpret_cast<float*>(p), p += BM * ROWF * 4;
    s.w2l = reinterpret_cast<float*>(p), p += 3 * 32 * 16 * 4;
    s.full = reinterpret_cast<uint64_t*>(p), p += BSTAGES * 8;
    s.empty = reinterpret_cast<uint64_t*>(p), p += BSTAGES * 8;
    s.accbar = reinterpret_cast<uint64_t*>(p), p += 8;
    s.stgbar = reinterpret_cast<uint64_t*>(p), p += 8;
    s.firstbar = reinterpret_cast<uint64_t*>(p), p += 8;
    s.tmem_base = reinterpret_cast<uint32_t*>(p);
    return s;
}

using MB = Math<true>;

__device__ __forceinline__ float hsum16(float v) {
#pragma unroll
    for (int m = 8; m >= 1; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
    return v;
}
__device__ __forceinline__ float hmax16(float v) {
#pragma unroll
    for (int m = 8; m >= 1; m >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, m));
    return v;
}
__device__ __forceinline__ float gsum(float v, int K) {
    for (int m = K >> 1; m >= 1; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
    return v;
}
// in registers: v[i] <- sum of v over i's group of K consecutive entries (K = 2, 4, 8, 16; static indices only)
__device__ __forceinline__ void gsum16(float (&v)[16], int K) {
#pragma unroll
    for (int m = 1; m < 16; m <<= 1) {
        if (m < K) {
            float t[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) t[i] = v[i] + v[i ^ m];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = t[i];
        }
    }
}

// ---- pre-pass: everything of the row-wise distribution backward that does not depend on the carried gradient ----------------
// stat[t][block][c][r][4] (c = value / 4): values 0-15 q, 16-31 ra, 32-47 rv, 48-63 softmax(la), 64-79 softmax(lv),
// 80-95 d lp (complete), 96-111 / 112-127 d la / d lv for a zero carried gradient.  The carried part is linear in d z:
//   dm = q (dz - gsum(q dz));  d la += dm ra - softmax(la) sum(dm ra);  d lv likewise.
// Also writes the [z_{t-1} | a_t | 0] operand plane of the first projector layer's weight gradient.
__global__ void __launch_bounds__(256) wide_bwd_rowstat_kernel(const WideRowstatArgs p) {
    const int lane = threadIdx.x & 31, j = lane & 15, K = p.K, A = p.A, T = p.T, F = p.D + 16;
    const long long items = (long long)p.NBBT * BM * T;
    const long long stride = (long long)gridDim.x * (blockDim.x / 16);
    for (long long base = (long long)blockIdx.x * (blockDim.x / 16) + (threadIdx.x >> 5) * 2; base < items; base += stride) {
        const long long it = base + (lane >> 4);
        const bool in = it < items;
        const int t = (int)((in ? it : 0) % T);
        const int rp = (int)((in ? it : 0) / T);  // padded row
        const bool v = in && rp < p.B;
        const long long bt = (long long)(v ? rp : 0) * T + t;
        const float q = p.post_probs[bt * 16 + j], pp = p.prior_probs[bt * 16 + j];
        const float la = p.logits[bt * 32 + j], lv = p.logits[bt * 32 + 16 + j];
        float dq = p.d_feature[bt * F + p.D + j], dpp = 0.f;
        if (p.d_post_probs != nullptr) dq += p.d_post_probs[bt * 16 + j];
        if (p.d_prior_probs != nullptr) dpp += p.d_prior_probs[bt * 16 + j];
        if (p.d_prior_stoch != nullptr) dpp += p.d_prior_stoch[bt * 16 + j];
        if (p.d_kl != nullptr) {
            const float g = p.d_kl[bt];
            dq += g * p.kl_wq * (clamp_log<true>(q) - clamp_log<true>(pp) + 1.f);
            dpp -= g * p.kl_wp * MB::div(q, fmaxf(pp, 1.1920928955078125e-07f));
        }
        const float dmixed = q * (dq - gsum(q * dq, K));
        const float dlp = pp * (dpp - gsum(pp * dpp, K));
        const float ma = hmax16(la), mv = hmax16(lv);
        const float lsa = la - ma - MB::log(hsum16(MB::exp(la - ma)));
        const float lsv = lv - mv - MB::log(hsum16(MB::exp(lv - mv)));
        const float f = lsa + lsv, mx = fmaxf(lsa, fmaxf(lsv, f));
        const float ea = MB::exp(lsa - mx), ev = MB::exp(lsv - mx), ef = MB::exp(f - mx);
        const float inv = MB::div(1.f, ea + ev + ef);
        const float ra = (ea + ef) * inv, rv = (ev + ef) * inv, sa = MB::exp(lsa), sv = MB::exp(lsv);
        const float dlsa = dmixed * ra, dlsv = dmixed * rv;
        const float dla = dlsa - sa * hsum16(dlsa), dlv = dlsv - sv * hsum16(dlsv);
        if (in) {
            const int bb = rp >> 7, r = rp & 127;
            float* st = p.stat + ((long long)t * p.NBBT + bb) * (32 * BM * 4) + r * 4 + (j & 3);
            const float vals[8] = {q, ra, rv, sa, sv, dlp, dla, dlv};
#pragma unroll
            for (int k = 0; k < 8; ++k) st[(long long)(k * 4 + (j >> 2)) * (BM * 4)] = v ? vals[k] : 0.f;
            __nv_bfloat16* xi = p.xin + (long long)t * p.dt_stride;
            const float zprev = !v ? 0.f : (t == 0 ? p.z0[(long long)rp * 16 + j] : p.feature[(bt - 1) * F + p.D + j]);
            xi[pk_off(bb, r, j, 32)] = __float2bfloat16_rn(zprev);
            xi[pk_off(bb, r, 16 + j, 32)] = __float2bfloat16_rn((v && j < A) ? p.actions[bt * A + j] : 0.f);
        }
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) mrssm_wide_bwd_kernel(const MrssmWideBwdArgs p) {
    extern __shared__ unsigned char smem_dyn[];
    const int D = p.D, KC = D >> 6, NSL = p.NSL, A = p.A, T = p.T, K = p.K, F = D + 16;
    const BwdSmem sm = carve_bwd(smem_dyn);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bb = blockIdx.x / NSL, s = blockIdx.x - bb * NSL;

    if (tid == 0) {
        for (int i = 0; i < BSTAGES; ++i) mbar_init(&sm.full[i], 1), mbar_init(&sm.empty[i], p.cs);  // every CTA of the cluster releases a slot
        mbar_init(sm.accbar, N_ISSUERS), mbar_init(sm.stgbar, 1), mbar_init(sm.firstbar, 1);  // both issuers commit the accumulator barrier
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < 3 * 32 * 16; i += NTHREADS) {  // w2l[h][j][o] = W2_h[o][32 s + j]
        const int h = i / 512, j = (i >> 4) & 31, o = i & 15;
        const float* w2 = h == 0 ? p.w.pr_w2 : (h == 1 ? p.w.au_w2 : p.w.vi_w2);
        sm.w2l[i] = w2[o * D + s * 32 + j];
    }
    for (int i = tid; i < BM * ROWF; i += NTHREADS) sm.carry[i] = 0.f;
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *sm.tmem_base;
    const uint32_t crank = p.cs > 1 ? cluster_rank() : 0;
    const uint16_t cmask = (uint16_t)((1u << p.cs) - 1u);
    if (p.cs > 1) cluster_sync_all();  // every CTA's barriers are initialised before any remote arrive / multicast copy

    Ring ring;
    uint32_t accph = 0, stgph = 0;
    unsigned epoch = 0;
    const uint32_t lboA = BM * 16, sbo = 128;
    const long long blk = (long long)bb * D * BM;

    auto load = [&](const __nv_bfloat16* a_src, const __nv_bfloat16* b_src, uint32_t b_bytes) {
        mbar_wait(&sm.empty[ring.slot], ring.phase ^ 1);
        unsigned char* st = sm.ring + (size_t)ring.slot * BSTAGE_BYTES;
        if (p.exp == 1) {
            mbar_expect_tx(&sm.full[ring.slot], 0);
        } else {
            mbar_expect_tx(&sm.full[ring.slot], A_BYTES + b_bytes);
            if (p.cs == 1) {
                bulk_g2s(st, a_src, A_BYTES, &sm.full[ring.slot]);
            } else {  // my 1/CS of the activation chunk, delivered to every CTA of the cluster
                const uint32_t piece = A_BYTES / p.cs;
                bulk_g2s_mc(st + crank * piece, reinterpret_cast<const unsigned char*>(a_src) + crank * piece, piece, &sm.full[ring.slot], cmask);
            }
            bulk_g2s(st + A_BYTES, b_src, b_bytes, &sm.full[ring.slot]);
        }
        ring.advance(BSTAGES);
    };
    const int issuer = warp == MMA_WARP ? 0 : 1;  // (meaningful in the issuer warps only)
    // chunk c of an accumulation group: issuer c % N_ISSUERS accumulates it into its own copy of the accumulator
    auto mma_chunk = [&](uint32_t tcol, int N, int c) {
        if (N_ISSUERS > 1 && (c & 1) != issuer) {
            ring.advance(BSTAGES);
            return;
        }
        mbar_wait(&sm.full[ring.slot], ring.phase);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sm.ring + (size_t)ring.slot * BSTAGE_BYTES), b0 = a0 + A_BYTES;
        const uint32_t lboB = N * 16;
        const uint32_t idesc = idesc_bf16(N, 0, 0);
        const uint32_t dcol = tmem + tcol + (N_ISSUERS > 1 ? issuer * BWD_ACC_OFF : 0);
        const bool first = c < N_ISSUERS;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
            if (p.exp != 2)
                umma(dcol, smem_desc(a0 + kk * 2 * lboA, lboA, sbo), smem_desc(b0 + kk * 2 * lboB, lboB, sbo), idesc,
                     (first && kk == 0) ? 0u : 1u);
        if (p.cs == 1) umma_commit(&sm.empty[ring.slot]);
        else umma_commit_mc(&sm.empty[ring.slot], cmask);
        ring.advance(BSTAGES);
    };

    const int row = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int half = (warp >> 2) & 1;  // epilogue warps 0-3: column groups 0-1 of the slice, warps 4-7: groups 2-3
    const int grow = bb * BM + row;
    const bool rvalid = grow < p.B;
    float* carry = sm.carry + row * ROWF;  // d h carried to the previous step, this thread's 32 units (shared memory, CTA-private)
    const uint4* stage16 = reinterpret_cast<const uint4*>(sm.stage);
    const float4* stage4 = reinterpret_cast<const float4*>(sm.stage);

    // planes of step t; 2 KB piece (8 features x 128 rows) of this CTA's slice
    auto rec = [&](int t, int plane) { return p.rec + (long long)t * p.t_stride + (long long)plane * p.plane_stride; };
    auto drec = [&](int t, int plane) { return p.drec + (long long)t * p.dt_stride + (long long)plane * p.plane_stride; };
    auto piece = [&](int qd) { return pk_off(bb, 0, s * 32 + qd * 8, D); };

    int tidx = 0;
    auto stamp = [&]() {
        if (p.timing != nullptr && blockIdx.x == 0 && tid == 0 && tidx < 500) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            p.timing[tidx++] = now;
        }
    };
    stamp();
    // P1 runs for t = T-1 .. 0 and once more as "t = -1" to finish d z0 / d a_0 of step 0
    for (int t = T - 1; t >= -1; --t) {
        // =================================== P1 ===================================
        const bool have_next = t < T - 1;  // dhid1_{t+1} exists
        if (warp == PRODUCER_WARP) {
            if (lane == 0) {
                if (t >= 0) {  // epilogue inputs: per-row statistics and this slice's head hiddens
                    mbar_expect_tx(sm.stgbar, STAT_BYTES + 12 * PIECE_BYTES);
                    bulk_g2s(sm.stage, p.stat + ((long long)t * p.stat_t_stride + (long long)bb * (32 * BM * 4)), STAT_BYTES, sm.stgbar);
                    for (int i = 0; i < 12; ++i)
                        bulk_g2s(sm.stage + STAT_BYTES + i * PIECE_BYTES, rec(t, P_PH + (i >> 2)) + piece(i & 3), PIECE_BYTES, sm.stgbar);
                }
                if (have_next) {
                    const __nv_bfloat16* a_src = drec(t + 1, DP_H1) + blk;
                    for (int c = 0; c < KC; ++c) load(a_src + (long long)c * (BM * 64), p.pW1x + (long long)c * (32 * 64), 32 * 64 * 2);
                }
            }
        } else if (warp == MMA_WARP || (N_ISSUERS > 1 && warp == MMA_WARP2)) {
            if (lane == 0 && have_next) {
                for (int c = 0; c < KC; ++c) mma_chunk(TB_DX, 32, c);
                umma_commit(sm.accbar);
            }
        } else {
            // ---- d [z_t ; a_{t+1}] of this thread's row ---------------------------------------------------------------------
            float dz[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) dz[i] = 0.f;
            if (have_next) {
                mbar_wait(sm.accbar, accph), accph ^= 1;
                tc_fence_after();
                float da[16];
                acc_ld16(tlane + TB_DX, BWD_ACC_OFF, dz);
                acc_ld16(tlane + TB_DX + 16, BWD_ACC_OFF, da);
                if (s == 0 && half == 0 && rvalid && p.d_actions != nullptr) {
#pragma unroll
                    for (int a = 0; a < 8; ++a)
                        if (a < A) p.d_actions[((long long)grow * T + (t + 1)) * A + a] = da[a];
                }
            }
            if (t < 0) {
                if (s == 0 && half == 0 && rvalid) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(p.d_z0 + (long long)grow * 16 + i) = make_float4(dz[i], dz[i + 1], dz[i + 2], dz[i + 3]);
                }
                if (rvalid) {
#pragma unroll
                    for (int i = 16 * half; i < 16 * half + 16; i += 4) *reinterpret_cast<float4*>(p.d_h0 + (long long)grow * D + s * 32 + i) = *reinterpret_cast<const float4*>(carry + i);
                }
            } else {
                mbar_wait(sm.stgbar, stgph), stgph ^= 1;
                auto ld16 = [&](int k, float (&v)[16]) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 x = stage4[(k * 4 + c) * BM + row];
                        v[4 * c] = x.x, v[4 * c + 1] = x.y, v[4 * c + 2] = x.z, v[4 * c + 3] = x.w;
                    }
                };
                // carried part of the distribution backward (linear in d z): dm = q (dz - gsum(q dz))
                float dl[48];
                {
                    float q[16], w[16];
                    ld16(0, q);
#pragma unroll
                    for (int i = 0; i < 16; ++i) w[i] = q[i] * dz[i];
                    gsum16(w, K);
#pragma unroll
                    for (int i = 0; i < 16; ++i) dz[i] = q[i] * (dz[i] - w[i]);  // dz now holds dm
                }
                ld16(5, *reinterpret_cast<float(*)[16]>(&dl[0]));
#pragma unroll
                for (int e = 0; e < 2; ++e) {  // audio, vision
                    float rr[16], sx[16], st[16];
                    ld16(1 + e, rr), ld16(3 + e, sx), ld16(6 + e, st);
                    float sum = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) rr[i] *= dz[i], sum += rr[i];
#pragma unroll
                    for (int i = 0; i < 16; ++i) dl[16 + 16 * e + i] = st[i] + rr[i] - sx[i] * sum;
                }
                if (s == 0 && half == 0) {  // d logits operand of the logit layers' weight gradients
                    __nv_bfloat16* dlg = drec(t, 0) + p.dlg_off;
#pragma unroll
                    for (int g = 0; g < 6; ++g) {
                        float v8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v8[i] = dl[g * 8 + i];
                        *reinterpret_cast<uint4*>(dlg + pk_off(bb, row, g * 8, 48)) = pack8(v8);
                    }
                }
                // ---- d head hidden of this row x slice: d logits . W2, * ELU' ------------------------------------------------
#pragma unroll
                for (int h = 0; h < 3; ++h) {
#pragma unroll 1
                    for (int qd = 2 * half; qd < 2 * half + 2; ++qd) {
                        float y[8], g[8];
                        unpack8(stage16[(STAT_BYTES / 16) + (h * 4 + qd) * BM + row], y);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4* w = reinterpret_cast<const float4*>(sm.w2l + (h * 32 + qd * 8 + i) * 16);
                            float acc = 0.f;
#pragma unroll
                            for (int o4 = 0; o4 < 4; ++o4) {
                                const float4 ww = w[o4];
                                acc = fmaf(dl[h * 16 + 4 * o4], ww.x, acc), acc = fmaf(dl[h * 16 + 4 * o4 + 1], ww.y, acc);
                                acc = fmaf(dl[h * 16 + 4 * o4 + 2], ww.z, acc), acc = fmaf(dl[h * 16 + 4 * o4 + 3], ww.w, acc);
                            }
                            g[i] = acc * elu_grad_from_out(y[i]);
                        }
                        *reinterpret_cast<uint4*>(drec(t, DP_PH + h) + piece(qd) + row * 8) = pack8(g);
                    }
                }
            }
        }
        if (t < 0) break;
        stamp();
        grid_sync(p.bar + bb * 64, epoch, p.status, NSL);
        stamp();
        // =================================== P2 ===================================
        if (warp == PRODUCER_WARP) {
            if (lane == 0) {
                mbar_expect_tx(sm.stgbar, 16 * PIECE_BYTES);  // gate record of this slice: r, z, n, hn
                for (int i = 0; i < 16; ++i) bulk_g2s(sm.stage + i * PIECE_BYTES, rec(t, P_R + (i >> 2)) + piece(i & 3), PIECE_BYTES, sm.stgbar);
                for (int c = 0; c < 3 * KC; ++c)
                    load(drec(t, DP_PH + c / KC) + blk + (long long)(c % KC) * (BM * 64), p.pWhdT + ((long long)s * 3 * KC + c) * (32 * 64), 32 * 64 * 2);
            }
        } else if (warp == MMA_WARP || (N_ISSUERS > 1 && warp == MMA_WARP2)) {
            if (lane == 0) {
                for (int c = 0; c < 3 * KC; ++c) mma_chunk(TB_DH, 32, c);
                umma_commit(sm.accbar);
            }
        } else {
            float* hps = reinterpret_cast<float*>(sm.stage + P2_HPREV_OFF) + row * ROWF;
            float* dfs = reinterpret_cast<float*>(sm.stage + P2_DFE_OFF) + row * ROWF;
            if (rvalid) {  // this row's h_{t-1} and upstream d h_t, fetched while the contraction runs
                const float* hprev = ((t == 0) ? p.h0 + (long long)grow * D : p.feature + ((long long)grow * T + (t - 1)) * F) + s * 32;
                const float* dfe = p.d_feature + ((long long)grow * T + t) * F + s * 32;
#pragma unroll
                for (int i = 4 * half; i < 4 * half + 4; ++i) cp_async16(hps + 4 * i, hprev + 4 * i), cp_async16(dfs + 4 * i, dfe + 4 * i);
            }
            cp_async_commit();
            mbar_wait(sm.accbar, accph), accph ^= 1;
            tc_fence_after();
            mbar_wait(sm.stgbar, stgph), stgph ^= 1;
            cp_async_wait_all();
#pragma unroll 1
            for (int qd = 2 * half; qd < 2 * half + 2; ++qd) {
                float dh[8], r[8], z[8], n[8], hn[8], hp[8], g0[8], g1[8], g2[8], g3[8];
                acc_ld8(tlane + TB_DH + qd * 8, BWD_ACC_OFF, dh);
                unpack8(stage16[(0 * 4 + qd) * BM + row], r);
                unpack8(stage16[(1 * 4 + qd) * BM + row], z);
                unpack8(stage16[(2 * 4 + qd) * BM + row], n);
                unpack8(stage16[(3 * 4 + qd) * BM + row], hn);
                if (rvalid) {
                    const float4 a = *reinterpret_cast<const float4*>(hps + qd * 8), b = *reinterpret_cast<const float4*>(hps + qd * 8 + 4);
                    hp[0] = a.x, hp[1] = a.y, hp[2] = a.z, hp[3] = a.w, hp[4] = b.x, hp[5] = b.y, hp[6] = b.z, hp[7] = b.w;
                    const float4 c = *reinterpret_cast<const float4*>(dfs + qd * 8), d = *reinterpret_cast<const float4*>(dfs + qd * 8 + 4);
                    dh[0] += c.x, dh[1] += c.y, dh[2] += c.z, dh[3] += c.w, dh[4] += d.x, dh[5] += d.y, dh[6] += d.z, dh[7] += d.w;
                    const float4 e = *reinterpret_cast<const float4*>(carry + qd * 8), f = *reinterpret_cast<const float4*>(carry + qd * 8 + 4);
                    dh[0] += e.x, dh[1] += e.y, dh[2] += e.z, dh[3] += e.w, dh[4] += f.x, dh[5] += f.y, dh[6] += f.z, dh[7] += f.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) hp[i] = 0.f, dh[i] = 0.f;
                }
                float cr[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float dn = dh[i] * (1.f - z[i]), dzg = dh[i] * (hp[i] - n[i]);
                    const float dnp = dn * (1.f - n[i] * n[i]);
                    const float drp = dnp * hn[i] * r[i] * (1.f - r[i]);
                    g0[i] = drp, g1[i] = dzg * z[i] * (1.f - z[i]), g2[i] = dnp, g3[i] = dnp * r[i];
                    cr[i] = dh[i] * z[i];
                }
                const long long o = piece(qd) + row * 8;
                *reinterpret_cast<uint4*>(drec(t, DP_GR) + o) = pack8(g0);
                *reinterpret_cast<uint4*>(drec(t, DP_GZ) + o) = pack8(g1);
                *reinterpret_cast<uint4*>(drec(t, DP_GIN) + o) = pack8(g2);
                *reinterpret_cast<uint4*>(drec(t, DP_GHN) + o) = pack8(g3);
                *reinterpret_cast<float4*>(carry + qd * 8) = make_float4(cr[0], cr[1], cr[2], cr[3]);
                *reinterpret_cast<float4*>(carry + qd * 8 + 4) = make_float4(cr[4], cr[5], cr[6], cr[7]);
            }
        }
        stamp();
        grid_sync(p.bar + bb * 64, epoch, p.status, NSL);
        stamp();
        // =================================== P3 ===================================
        if (warp == PRODUCER_WARP) {
            if (lane == 0) {
                for (int c = 0; c < 2 * KC; ++c)
                    load(drec(t, DP_GR + c / KC) + blk + (long long)(c % KC) * (BM * 64), p.pWgT + ((long long)s * 2 * KC + c) * (64 * 64), 64 * 64 * 2);
                for (int c = 0; c < KC; ++c) load(drec(t, DP_GIN) + blk + (long long)c * (BM * 64), p.pWihTn + ((long long)s * KC + c) * (32 * 64), 32 * 64 * 2);
                for (int c = 0; c < KC; ++c) load(drec(t, DP_GHN) + blk + (long long)c * (BM * 64), p.pWhhTn + ((long long)s * KC + c) * (32 * 64), 32 * 64 * 2);
            }
        } else if (warp == MMA_WARP || (N_ISSUERS > 1 && warp == MMA_WARP2)) {
            if (lane == 0) {
                for (int c = 0; c < 2 * KC; ++c) mma_chunk(TB_P3, 64, c);
                for (int c = 0; c < KC; ++c) mma_chunk(TB_P3, 32, 2 * KC + c);
                for (int c = 0; c < KC; ++c) mma_chunk(TB_P3 + 32, 32, 3 * KC + c);
                umma_commit(sm.accbar);
            }
        } else {
            mbar_wait(sm.accbar, accph), accph ^= 1;
            tc_fence_after();
#pragma unroll 1
            for (int qd = 2 * half; qd < 2 * half + 2; ++qd) {
                float v[8], c[8];
                acc_ld8(tlane + TB_P3 + qd * 8, BWD_ACC_OFF, v);
                *reinterpret_cast<uint4*>(drec(t, DP_X2) + piece(qd) + row * 8) = pack8(v);
                acc_ld8(tlane + TB_P3 + 32 + qd * 8, BWD_ACC_OFF, c);
                float4 e = *reinterpret_cast<const float4*>(carry + qd * 8), f = *reinterpret_cast<const float4*>(carry + qd * 8 + 4);
                e.x += c[0], e.y += c[1], e.z += c[2], e.w += c[3], f.x += c[4], f.y += c[5], f.z += c[6], f.w += c[7];
                *reinterpret_cast<float4*>(carry + qd * 8) = e, *reinterpret_cast<float4*>(carry + qd * 8 + 4) = f;
            }
        }
        stamp();
        grid_sync(p.bar + bb * 64, epoch, p.status, NSL);
        stamp();
        // =================================== P4 ===================================
        if (warp == PRODUCER_WARP) {
            if (lane == 0) {
                mbar_expect_tx(sm.stgbar, 4 * PIECE_BYTES);  // hid1_t of this slice (ELU')
                for (int i = 0; i < 4; ++i) bulk_g2s(sm.stage + i * PIECE_BYTES, rec(t, P_HID1) + piece(i), PIECE_BYTES, sm.stgbar);
                for (int c = 0; c < KC; ++c) load(drec(t, DP_X2) + blk + (long long)c * (BM * 64), p.pW2T + ((long long)s * KC + c) * (32 * 64), 32 * 64 * 2);
            }
        } else if (warp == MMA_WARP || (N_ISSUERS > 1 && warp == MMA_WARP2)) {
            if (lane == 0) {
                for (int c = 0; c < KC; ++c) mma_chunk(TB_H1, 32, c);
                umma_commit(sm.accbar);
            }
        } else {
            mbar_wait(sm.accbar, accph), accph ^= 1;
            tc_fence_after();
            mbar_wait(sm.stgbar, stgph), stgph ^= 1;
#pragma unroll 1
            for (int qd = 2 * half; qd < 2 * half + 2; ++qd) {
                float v[8], y[8];
                acc_ld8(tlane + TB_H1 + qd * 8, BWD_ACC_OFF, v);
                unpack8(stage16[qd * BM + row], y);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] *= elu_grad_from_out(y[i]);
                *reinterpret_cast<uint4*>(drec(t, DP_H1) + piece(qd) + row * 8) = pack8(v);
            }
        }
        stamp();
        grid_sync(p.bar + bb * 64, epoch, p.status, NSL);
        stamp();
    }

    if (p.cs > 1) {
        // no CTA may leave while a peer can still multicast into its shared memory or arrive on its barriers: take every ring
        // slot once more (= all peers' last commits have arrived here), then meet the cluster
        if (warp == PRODUCER_WARP && lane == 0)
            for (int i = 0; i < BSTAGES; ++i) mbar_wait(&sm.empty[ring.slot], ring.phase ^ 1), ring.advance(BSTAGES);
        __syncthreads();
        cluster_sync_all();
    }
    tc_fence_before();
    __syncthreads();
    __syncwarp();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- transposed weight images for the backward contractions -----------------------------------------------------------------
// B operand row n, column k of job `y` (per slice s unless stated):
//   0 pW1x   [32 x D]   (n, k) = asp_w1[k][A + n] for n < 16 (z), asp_w1[k][n - 16] for 16 <= n < 16 + A (action), else 0   (no slices)
//   1 pWhdT  [32 x 3D]  (u, k) = {pr_w1, au_w1, vi_w1}[k mod D][32 s + u]
//   2 pWgT   [64 x 2D]  (n, k) = n < 32 ? w_ih[k][32 s + n] : w_hh[k][32 s + n - 32]      (r and z gate rows k < 2D)
//   3 pWihTn [32 x D]   (u, k) = w_ih[2D + k][32 s + u]
//   4 pWhhTn [32 x D]   (u, k) = w_hh[2D + k][32 s + u]
//   5 pW2T   [32 x D]   (u, k) = asp_w2[k][32 s + u]
__global__ void wide_pack_bwd_weights_kernel(const MrssmWideBwdArgs p) {
    const int D = p.D, job = blockIdx.y, NSL = job == 0 ? 1 : p.NSL, A = p.A;
    const int N = job == 2 ? 64 : 32, Kt = job == 1 ? 3 * D : (job == 2 ? 2 * D : D), KCt = Kt >> 6;
    __nv_bfloat16* dst = const_cast<__nv_bfloat16*>(job == 0 ? p.pW1x : job == 1 ? p.pWhdT : job == 2 ? p.pWgT : job == 3 ? p.pWihTn : job == 4 ? p.pWhhTn : p.pW2T);
    const long long total = (long long)NSL * KCt * 8 * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i % N), kg = (int)((i / N) % 8), c = (int)((i / (8LL * N)) % KCt), s = (int)(i / (8LL * N * KCt));
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = c * 64 + kg * 8 + e;
            float x;
            switch (job) {
                case 0: x = n < 16 ? p.w.asp_w1[(long long)k * (A + 16) + A + n] : (n < 16 + A ? p.w.asp_w1[(long long)k * (A + 16) + n - 16] : 0.f); break;
                case 1: {
                    const int h = k / D, kk = k - h * D;
                    x = h == 0 ? p.w.pr_w1[(long long)kk * D + s * 32 + n] : (h == 1 ? p.w.au_w1 : p.w.vi_w1)[(long long)kk * (D + 64) + s * 32 + n];
                    break;
                }
                case 2: x = n < 32 ? p.w.w_ih[(long long)k * D + s * 32 + n] : p.w.w_hh[(long long)k * D + s * 32 + n - 32]; break;
                case 3: x = p.w.w_ih[(long long)(2 * D + k) * D + s * 32 + n]; break;
                case 4: x = p.w.w_hh[(long long)(2 * D + k) * D + s * 32 + n]; break;
                default: x = p.w.asp_w2[(long long)k * D + s * 32 + n]; break;
            }
            v[e] = x;
        }
        *reinterpret_cast<uint4*>(dst + i * 8) = pack8(v);
    }
}

}  // namespace wide

cudaError_t launch_wide_bwd_rowstat(const WideRowstatArgs& a, cudaStream_t s) {
    const long long items = (long long)a.NBBT * wide::BM * a.T;
    const long long blocks = (items + 15) / 16;
    wide::wide_bwd_rowstat_kernel<<<(unsigned)(blocks < 8192 ? blocks : 8192), 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_wide_pack_bwd_weights(const MrssmWideBwdArgs& a, cudaStream_t s) {
    wide::wide_pack_bwd_weights_kernel<<<dim3(64, 6), 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_mrssm_wide_bwd(const MrssmWideBwdArgs& a, cudaStream_t s) {
    MrssmWideBwdArgs args = a;
    return launch_wide_persistent((const void*)wide::mrssm_wide_bwd_kernel, &args, a.NBB * a.NSL, a.cs, wide::bwd_smem_bytes(), s);
}

}  // namespace rssm
