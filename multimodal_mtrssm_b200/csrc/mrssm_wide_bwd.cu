// Wide MoPoE-MRSSM rollout, backward (BPTT): one persistent cooperative kernel walks the T steps in reverse and leaves every
// layer's pre-activation gradient in packed bf16 planes; the weight gradients (contractions over all (b,t) rows) and the
// embedding gradients are non-recurrent and run afterwards (mrssm_wide_wgrad.cu).  autograd twin of mrssm_wide_fwd.cu; same
// CTA = (batch block bb, slice s) decomposition, operand layout and barrier scheme (wide_common.cuh).
//
// One step t (reverse) = four phases separated by grid barriers:
//   P1  d[a_{t+1} ; z_t] = dhid1_{t+1} . W1                      (every slice CTA of the block, redundantly)   N = 32, K = D
//       row-wise: d z_t -> straight-through -> d q, KL terms, per-group softmax / MoPoE fusion / flat log-softmax backward
//       -> d logits (prior, audio, vision); d head hidden = d logits . W2 (CUDA cores), * ELU' -> planes DPH / DAH / DVH [slice]
//   P2  d h_t = [DPH | DAH | DVH] . [W_prior | W_audio | W_vision][:, slice] + carry + upstream                  N = 32, K = 3D
//       GRU gate backward -> planes DG_R, DG_Z, DGI_N, DGH_N [slice]; carry = d h_t * z
//   P3  d x2 = [DG_R | DG_Z | DGI_N] . W_ih[:, slice] ; carry += [DG_R | DG_Z | DGH_N] . W_hh[:, slice]          N = 64 / 32, K = 3D
//   P4  d hid1_t = (d x2 . W2[:, slice]) * ELU'                                                                  N = 32, K = D
#include "kernels.h"
#include "wide_common.cuh"

namespace rssm {
namespace wide {

constexpr int BSTAGES = 6;
constexpr int BB_MAX_BYTES = 64 * 64 * 2;
constexpr int BSTAGE_BYTES = A_BYTES + BB_MAX_BYTES;
constexpr int TB_DX = 0, TB_DH = 32, TB_P3 = 64, TB_H1 = 128;

struct BwdSmem {
    unsigned char* ring;
    float *w2l, *dlog, *dzc;
    uint64_t *full, *empty, *accbar;
    uint32_t* tmem_base;
};
__host__ __device__ inline size_t bwd_smem_bytes() {
    return 128 + (size_t)BSTAGES * BSTAGE_BYTES + 3 * 32 * 16 * 4 + BM * 48 * 4 + BM * 16 * 4 + (2 * BSTAGES + 1) * 8 + 16;
}
__device__ __forceinline__ BwdSmem carve_bwd(unsigned char* dyn) {
    BwdSmem s;
    unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn) + 127) & ~(uintptr_t)127);
    s.ring = p, p += (size_t)BSTAGES * BSTAGE_BYTES;
    s.w2l = reinterpret_cast<float*>(p), p += 3 * 32 * 16 * 4;
    s.dlog = reinterpret_cast<float*>(p), p += BM * 48 * 4;
    s.dzc = reinterpret_cast<float*>(p), p += BM * 16 * 4;
    s.full = reinterpret_cast<uint64_t*>(p), p += BSTAGES * 8;
    s.empty = reinterpret_cast<uint64_t*>(p), p += BSTAGES * 8;
    s.accbar = reinterpret_cast<uint64_t*>(p), p += 8;
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

__global__ void __launch_bounds__(NTHREADS, 1) mrssm_wide_bwd_kernel(const MrssmWideBwdArgs p) {
    extern __shared__ unsigned char smem_dyn[];
    const int D = p.D, KC = D >> 6, NSL = p.NSL, A = p.A, T = p.T, K = p.K, F = D + 16;
    const BwdSmem sm = carve_bwd(smem_dyn);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bb = blockIdx.x / NSL, s = blockIdx.x - bb * NSL;

    if (tid == 0) {
        for (int i = 0; i < BSTAGES; ++i) mbar_init(&sm.full[i], 1), mbar_init(&sm.empty[i], 1);
        mbar_init(sm.accbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < 3 * 32 * 16; i += NTHREADS) {  // w2l[h][j][o] = W2_h[o][32 s + j]
        const int h = i / 512, j = (i >> 4) & 31, o = i & 15;
        const float* w2 = h == 0 ? p.w.pr_w2 : (h == 1 ? p.w.au_w2 : p.w.vi_w2);
        sm.w2l[i] = w2[o * D + s * 32 + j];
    }
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *sm.tmem_base;

    Ring ring;
    uint32_t accph = 0;
    unsigned epoch = 0;
    const uint32_t lboA = BM * 16, sbo = 128;
    const long long blk = (long long)bb * D * BM;

    auto load = [&](const __nv_bfloat16* a_src, const __nv_bfloat16* b_src, uint32_t b_bytes) {
        mbar_wait(&sm.empty[ring.slot], ring.phase ^ 1);
        mbar_expect_tx(&sm.full[ring.slot], A_BYTES + b_bytes);
        unsigned char* st = sm.ring + (size_t)ring.slot * BSTAGE_BYTES;
        bulk_g2s(st, a_src, A_BYTES, &sm.full[ring.slot]);
        bulk_g2s(st + A_BYTES, b_src, b_bytes, &sm.full[ring.slot]);
        ring.advance(BSTAGES);
    };
    auto mma_chunk = [&](uint32_t tcol, int N, bool first) {
        mbar_wait(&sm.full[ring.slot], ring.phase);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sm.ring + (size_t)ring.slot * BSTAGE_BYTES), b0 = a0 + A_BYTES;
        const uint32_t lboB = N * 16;
        const uint32_t idesc = idesc_bf16(N, 0, 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
            umma(tmem + tcol, smem_desc(a0 + kk * 2 * lboA, lboA, sbo), smem_desc(b0 + kk * 2 * lboB, lboB, sbo), idesc,
                 (first && kk == 0) ? 0u : 1u);
        umma_commit(&sm.empty[ring.slot]);
        ring.advance(BSTAGES);
    };

    const int row = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int grow = bb * BM + row;
    const bool rvalid = grow < p.B;
    float* carry = p.carry + (long long)grow * D + s * 32;  // d h carried to the previous step, this thread's 32 units

    // planes of step t
    auto rec = [&](int t, int plane) { return p.rec + (long long)t * p.t_stride + (long long)plane * p.plane_stride; };
    auto drec = [&](int t, int plane) { return p.drec + (long long)t * p.dt_stride + (long long)plane * p.plane_stride; };

    // P1 runs for t = T-1 .. 0 and once more as "t = -1" to finish d z0 / d a_0 of step 0
    for (int t = T - 1; t >= -1; --t) {
        // =================================== P1 ===================================
        const bool have_next = t < T - 1;  // dhid1_{t+1} exists
        if (warp == 4) {
            if (lane == 0 && have_next) {
                const __nv_bfloat16* a_src = drec(t + 1, DP_H1) + blk;
                for (int c = 0; c < KC; ++c) load(a_src + (long long)c * (BM * 64), p.pW1x + (long long)c * (32 * 64), 32 * 64 * 2);
            }
        } else if (warp == 5) {
            if (lane == 0 && have_next) {
                for (int c = 0; c < KC; ++c) mma_chunk(TB_DX, 32, c == 0);
                umma_commit(sm.accbar);
            }
        } else {
            // ---- d [a_{t+1} ; z_t] of this thread's row -------------------------------------------------------------------
            float dx[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) dx[i] = 0.f;
            if (have_next) {
                mbar_wait(sm.accbar, accph), accph ^= 1;
                tc_fence_after();
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float v[16];
                    tmem_ld16(tlane + TB_DX + q * 16, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) dx[q * 16 + i] = v[i];
                }
                if (s == 0 && rvalid && p.d_actions != nullptr) {
#pragma unroll
                    for (int a = 0; a < 8; ++a)
                        if (a < A) p.d_actions[((long long)grow * T + (t + 1)) * A + a] = dx[a];
                }
            }
            // z part starts at column A (runtime): through shared memory
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i >= A && i < A + 16) sm.dzc[row * 16 + (i - A)] = dx[i];
            __syncwarp();
            if (t < 0) {
                if (s == 0 && rvalid) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) p.d_z0[(long long)grow * 16 + j] = sm.dzc[row * 16 + j];
                }
                if (rvalid) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(p.d_h0 + (long long)grow * D + s * 32 + i) = *reinterpret_cast<const float4*>(carry + i);
                }
            } else {
                // ---- row-wise distribution backward: half-warp per row, lane j = stochastic column ---------------------------
                const int j = lane & 15;
                for (int it = 0; it < 16; ++it) {
                    const int r = (warp & 3) * 32 + it * 2 + (lane >> 4);
                    const int gr = bb * BM + r;
                    const bool v = gr < p.B;
                    const long long bt = (long long)(v ? gr : 0) * T + t;
                    const float q = p.post_probs[bt * 16 + j], pp = p.prior_probs[bt * 16 + j];
                    const float la = p.logits[bt * 32 + j], lv = p.logits[bt * 32 + 16 + j];
                    float dq = p.d_feature[bt * F + D + j] + sm.dzc[r * 16 + j];
                    float dpp = 0.f;
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
                    const float dlsa = dmixed * (ea + ef) * inv, dlsv = dmixed * (ev + ef) * inv;
                    float dla = dlsa - MB::exp(lsa) * hsum16(dlsa);
                    float dlv = dlsv - MB::exp(lsv) * hsum16(dlsv);
                    float dlpv = dlp;
                    if (!v) dla = 0.f, dlv = 0.f, dlpv = 0.f;
                    sm.dlog[r * 48 + j] = dlpv, sm.dlog[r * 48 + 16 + j] = dla, sm.dlog[r * 48 + 32 + j] = dlv;
                    if (s == 0) {  // operands of the small weight-gradient contractions: d logits and [a_t ; z_{t-1}]
                        __nv_bfloat16* dl = drec(t, 0) + p.dlg_off;
                        dl[pk_off(bb, r, j, 48)] = __float2bfloat16_rn(dlpv);
                        dl[pk_off(bb, r, 16 + j, 48)] = __float2bfloat16_rn(dla);
                        dl[pk_off(bb, r, 32 + j, 48)] = __float2bfloat16_rn(dlv);
                        __nv_bfloat16* xi = drec(t, 0) + p.xin_off;
                        const float zprev = !v ? 0.f : (t == 0 ? p.z0[(long long)gr * 16 + j] : p.feature[(bt - 1) * F + D + j]);
                        xi[pk_off(bb, r, A + j, 32)] = __float2bfloat16_rn(zprev);
                        const float av = (v && j < A) ? p.actions[bt * A + j] : 0.f;
                        if (j < A) xi[pk_off(bb, r, j, 32)] = __float2bfloat16_rn(av);
                        if (A + 16 + j < 32) xi[pk_off(bb, r, A + 16 + j, 32)] = __float2bfloat16_rn(0.f);
                    }
                }
                __syncwarp();
                // ---- d head hidden of this row x slice: d logits . W2, * ELU' ------------------------------------------------
                float dl[48];
#pragma unroll
                for (int o = 0; o < 48; o += 4) {
                    const float4 x = *reinterpret_cast<const float4*>(sm.dlog + row * 48 + o);
                    dl[o] = x.x, dl[o + 1] = x.y, dl[o + 2] = x.z, dl[o + 3] = x.w;
                }
#pragma unroll 1
                for (int h = 0; h < 3; ++h) {
#pragma unroll 1
                    for (int qd = 0; qd < 4; ++qd) {
                        const long long o = pk_off(bb, row, s * 32 + qd * 8, D);
                        float y[8], g[8];
                        unpack8(*reinterpret_cast<const uint4*>(rec(t, P_PH + h) + o), y);
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
                        *reinterpret_cast<uint4*>(drec(t, DP_PH + h) + o) = pack8(g);
                    }
                }
            }
        }
        if (t < 0) break;
        grid_sync(p.bar, epoch, p.status);
        // =================================== P2 ===================================
        if (warp == 4) {
            if (lane == 0)
                for (int c = 0; c < 3 * KC; ++c)
                    load(drec(t, DP_PH + c / KC) + blk + (long long)(c % KC) * (BM * 64), p.pWhdT + ((long long)s * 3 * KC + c) * (32 * 64), 32 * 64 * 2);
        } else if (warp == 5) {
            if (lane == 0) {
                for (int c = 0; c < 3 * KC; ++c) mma_chunk(TB_DH, 32, c == 0);
                umma_commit(sm.accbar);
            }
        } else {
            mbar_wait(sm.accbar, accph), accph ^= 1;
            tc_fence_after();
            const float* hprev = (t == 0) ? p.h0 + (long long)grow * D : p.feature + ((long long)grow * T + (t - 1)) * F;
            const float* dfe = p.d_feature + ((long long)grow * T + t) * F;
#pragma unroll 1
            for (int qd = 0; qd < 4; ++qd) {
                float dh[8], r[8], z[8], n[8], hn[8], hp[8], g0[8], g1[8], g2[8], g3[8];
                tmem_ld8(tlane + TB_DH + qd * 8, dh);
                const long long o = pk_off(bb, row, s * 32 + qd * 8, D);
                unpack8(*reinterpret_cast<const uint4*>(rec(t, P_R) + o), r);
                unpack8(*reinterpret_cast<const uint4*>(rec(t, P_Z) + o), z);
                unpack8(*reinterpret_cast<const uint4*>(rec(t, P_N) + o), n);
                unpack8(*reinterpret_cast<const uint4*>(rec(t, P_HN) + o), hn);
                if (rvalid) {
                    const float4 a = *reinterpret_cast<const float4*>(hprev + s * 32 + qd * 8), b = *reinterpret_cast<const float4*>(hprev + s * 32 + qd * 8 + 4);
                    hp[0] = a.x, hp[1] = a.y, hp[2] = a.z, hp[3] = a.w, hp[4] = b.x, hp[5] = b.y, hp[6] = b.z, hp[7] = b.w;
                    const float4 c = *reinterpret_cast<const float4*>(dfe + s * 32 + qd * 8), d = *reinterpret_cast<const float4*>(dfe + s * 32 + qd * 8 + 4);
                    dh[0] += c.x, dh[1] += c.y, dh[2] += c.z, dh[3] += c.w, dh[4] += d.x, dh[5] += d.y, dh[6] += d.z, dh[7] += d.w;
                    if (t < T - 1) {
                        const float4 e = *reinterpret_cast<const float4*>(carry + qd * 8), f = *reinterpret_cast<const float4*>(carry + qd * 8 + 4);
                        dh[0] += e.x, dh[1] += e.y, dh[2] += e.z, dh[3] += e.w, dh[4] += f.x, dh[5] += f.y, dh[6] += f.z, dh[7] += f.w;
                    }
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
                *reinterpret_cast<uint4*>(drec(t, DP_GR) + o) = pack8(g0);
                *reinterpret_cast<uint4*>(drec(t, DP_GZ) + o) = pack8(g1);
                *reinterpret_cast<uint4*>(drec(t, DP_GIN) + o) = pack8(g2);
                *reinterpret_cast<uint4*>(drec(t, DP_GHN) + o) = pack8(g3);
                *reinterpret_cast<float4*>(carry + qd * 8) = make_float4(cr[0], cr[1], cr[2], cr[3]);
                *reinterpret_cast<float4*>(carry + qd * 8 + 4) = make_float4(cr[4], cr[5], cr[6], cr[7]);
            }
        }
        grid_sync(p.bar, epoch, p.status);
        // =================================== P3 ===================================
        if (warp == 4) {
            if (lane == 0) {
                for (int c = 0; c < 2 * KC; ++c)
                    load(drec(t, DP_GR + c / KC) + blk + (long long)(c % KC) * (BM * 64), p.pWgT + ((long long)s * 2 * KC + c) * (64 * 64), 64 * 64 * 2);
                for (int c = 0; c < KC; ++c) load(drec(t, DP_GIN) + blk + (long long)c * (BM * 64), p.pWihTn + ((long long)s * KC + c) * (32 * 64), 32 * 64 * 2);
                for (int c = 0; c < KC; ++c) load(drec(t, DP_GHN) + blk + (long long)c * (BM * 64), p.pWhhTn + ((long long)s * KC + c) * (32 * 64), 32 * 64 * 2);
            }
        } else if (warp == 5) {
            if (lane == 0) {
                for (int c = 0; c < 2 * KC; ++c) mma_chunk(TB_P3, 64, c == 0);
                for (int c = 0; c < KC; ++c) mma_chunk(TB_P3, 32, false);
                for (int c = 0; c < KC; ++c) mma_chunk(TB_P3 + 32, 32, false);
                umma_commit(sm.accbar);
            }
        } else {
            mbar_wait(sm.accbar, accph), accph ^= 1;
            tc_fence_after();
#pragma unroll 1
            for (int qd = 0; qd < 4; ++qd) {
                float v[8], c[8];
                tmem_ld8(tlane + TB_P3 + qd * 8, v);
                *reinterpret_cast<uint4*>(drec(t, DP_X2) + pk_off(bb, row, s * 32 + qd * 8, D)) = pack8(v);
                tmem_ld8(tlane + TB_P3 + 32 + qd * 8, c);
                if (rvalid) {
                    float4 e = *reinterpret_cast<const float4*>(carry + qd * 8), f = *reinterpret_cast<const float4*>(carry + qd * 8 + 4);
                    e.x += c[0], e.y += c[1], e.z += c[2], e.w += c[3], f.x += c[4], f.y += c[5], f.z += c[6], f.w += c[7];
                    *reinterpret_cast<float4*>(carry + qd * 8) = e, *reinterpret_cast<float4*>(carry + qd * 8 + 4) = f;
                }
            }
        }
        grid_sync(p.bar, epoch, p.status);
        // =================================== P4 ===================================
        if (warp == 4) {
            if (lane == 0)
                for (int c = 0; c < KC; ++c) load(drec(t, DP_X2) + blk + (long long)c * (BM * 64), p.pW2T + ((long long)s * KC + c) * (32 * 64), 32 * 64 * 2);
        } else if (warp == 5) {
            if (lane == 0) {
                for (int c = 0; c < KC; ++c) mma_chunk(TB_H1, 32, c == 0);
                umma_commit(sm.accbar);
            }
        } else {
            mbar_wait(sm.accbar, accph), accph ^= 1;
            tc_fence_after();
#pragma unroll 1
            for (int qd = 0; qd < 4; ++qd) {
                float v[8], y[8];
                tmem_ld8(tlane + TB_H1 + qd * 8, v);
                const long long o = pk_off(bb, row, s * 32 + qd * 8, D);
                unpack8(*reinterpret_cast<const uint4*>(rec(t, P_HID1) + o), y);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] *= elu_grad_from_out(y[i]);
                *reinterpret_cast<uint4*>(drec(t, DP_H1) + o) = pack8(v);
            }
        }
        grid_sync(p.bar, epoch, p.status);
    }

    tc_fence_before();
    __syncthreads();
    __syncwarp();
    if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// ---- transposed weight images for the backward contractions -----------------------------------------------------------------
// B operand row n, column k of job `y` (per slice s unless stated):
//   0 pW1x   [32 x D]   (n, k) = asp_w1[k][n]            n < A+16, else 0                 (one image, no slices)
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
                case 0: x = n < A + 16 ? p.w.asp_w1[(long long)k * (A + 16) + n] : 0.f; break;
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

cudaError_t launch_wide_pack_bwd_weights(const MrssmWideBwdArgs& a, cudaStream_t s) {
    wide::wide_pack_bwd_weights_kernel<<<dim3(64, 6), 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_mrssm_wide_bwd(const MrssmWideBwdArgs& a, cudaStream_t s) {
    const size_t smem = wide::bwd_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(wide::mrssm_wide_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    MrssmWideBwdArgs args = a;
    void* params[] = {&args};
    return cudaLaunchCooperativeKernel((const void*)wide::mrssm_wide_bwd_kernel, dim3(a.NBB * a.NSL), dim3(wide::NTHREADS), params, smem, s);
}

}  // namespace rssm
