// Wide MoPoE-MRSSM rollout, forward (and imagination): one persistent cooperative kernel for all T steps.
// Replaces the T loop of MoPoE_MRSSM.rollout_representation (mrssm/mopoe_mrssm/core.py:221-256) / BaseRSSM.rollout_transition
// (core.py:170-185) for deterministic_size = hidden_size = D in {64, 128, ..., 512}.  Design notes: wide_common.cuh.
//
// CTA (bb, s) = batch block bb (128 sequences) x slice s (hidden units / head features 32 s .. 32 s + 31).  One step = four
// phases separated by barriers among the slice CTAs of a batch block (each phase consumes what ALL slices of the block produced in the previous one):
//   A  x2[:, slice]  = W2[slice] . hid1_t + b2                 (action_state_projector.2, networks.py:169)      N = 32, K = D
//      gh[:, slice]  = W_hh[r|z|n of slice] . h_{t-1}           (GRUCell, networks.py:170)                       N = 96, K = D
//   B  gi[:, slice]  = W_ih[r|z|n of slice] . x2 ; gates ; h_t  (GRUCell)                                        N = 96, K = D
//   C  hid[:, slice] = ELU([W_prior | W_audio | W_vision][slice] . h_t (+ W_e . embed_t) + b1)                   N = 96, K = D (+64)
//      partial logits of the slice (CUDA cores; 32 features x 16 outputs per head)  -> part[row][s][48]
//   D  the block's rows are dealt out evenly over its slice CTAs: sum the partial logits, MoPoE fusion (mopoe_mrssm/core.py:241-251,135-154),
//      per-group softmax, inverse-CDF draw, KL, per-step outputs; hid1_{t+1} = ELU(W1 . [a_{t+1} ; z_t] + b1) (networks.py:164-169)
// Accumulators live in TMEM; gh is issued in phase A and consumed in phase B.
#include "kernels.h"
#include "wide_common.cuh"

namespace rssm {
namespace wide {

constexpr int STAGES = 5;
constexpr int B_MAX_BYTES = 96 * 64 * 2;
constexpr int STAGE_BYTES = A_BYTES + B_MAX_BYTES;
// TMEM columns: gh lives from phase A to phase B; x2 (phase A), gi (B) and the heads (C) take turns in one region.  With two
// issuers every accumulator has a second copy FWD_ACC_OFF columns further on.
constexpr int TM_GH = 0, TM_R = 96 * N_ISSUERS, TM_X2 = TM_R, TM_GI = TM_R, TM_HD = TM_R, FWD_ACC_OFF = 96;
constexpr int MAX_RPC = 64;  // rows per CTA in phase D (NSL >= 2)
constexpr int HROW = 36;     // padded fp32 row of the CTA's own h slice (conflict-free float4 access)

struct FwdSmem {
    unsigned char* ring;
    float *w1t, *b2s, *bih, *bhh, *bhd, *w2l, *b2l, *zval, *hs, *acts;
    int* zidx;
    uint64_t *full, *empty, *accbar, *firstbar;
    uint32_t* tmem_base;
};

__host__ __device__ inline size_t fwd_smem_bytes(int D, int A) {
    size_t n = 128;  // alignment slack
    n += (size_t)STAGES * STAGE_BYTES;
    n += (size_t)(A + 17) * D * 4;
    n += (32 + 96 * 3) * 4 + 3 * 32 * 16 * 4 + 48 * 4 + MAX_RPC * 16 * 4;
    n += BM * HROW * 4 + MAX_RPC * 8 * 4 + MAX_RPC * 8 * 4;
    n += (2 * STAGES + 2) * 8 + 16;
    return n;
}

__device__ __forceinline__ FwdSmem carve(unsigned char* dyn, int D, int A) {
    FwdSmem s;
    unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn) + 127) & ~(uintptr_t)127);
    s.ring = p, p += (size_t)STAGES * STAGE_BYTES;
    s.w1t = reinterpret_cast<float*>(p), p += (size_t)(A + 17) * D * 4;
    s.b2s = reinterpret_cast<float*>(p), p += 32 * 4;
    s.bih = reinterpret_cast<float*>(p), p += 96 * 4;
    s.bhh = reinterpret_cast<float*>(p), p += 96 * 4;
    s.bhd = reinterpret_cast<float*>(p), p += 96 * 4;
    s.w2l = reinterpret_cast<float*>(p), p += 3 * 32 * 16 * 4;
    s.b2l = reinterpret_cast<float*>(p), p += 48 * 4;
    s.zval = reinterpret_cast<float*>(p), p += MAX_RPC * 16 * 4;
    s.hs = reinterpret_cast<float*>(p), p += BM * HROW * 4;
    s.acts = reinterpret_cast<float*>(p), p += MAX_RPC * 8 * 4;
    s.zidx = reinterpret_cast<int*>(p), p += MAX_RPC * 8 * 4;
    s.full = reinterpret_cast<uint64_t*>(p), p += STAGES * 8;
    s.empty = reinterpret_cast<uint64_t*>(p), p += STAGES * 8;
    s.accbar = reinterpret_cast<uint64_t*>(p), p += 8;
    s.firstbar = reinterpret_cast<uint64_t*>(p), p += 8;
    s.tmem_base = reinterpret_cast<uint32_t*>(p);
    return s;
}

using M = Math<true>;

__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
    for (int m = 8; m >= 1; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
    return v;
}
__device__ __forceinline__ float half_max(float v) {
#pragma unroll
    for (int m = 8; m >= 1; m >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, m));
    return v;
}
// softmax over the K lanes of this lane's group (K = 2, 4, 8, 16 consecutive lanes)
__device__ __forceinline__ float group_softmax(float x, int K) {
    float mx = x;
    for (int m = K >> 1; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, m));
    const float e = M::exp(x - mx);
    float sum = e;
    for (int m = K >> 1; m >= 1; m >>= 1) sum += __shfl_xor_sync(FULL, sum, m);
    return M::div(e, sum);
}
// inverse-CDF draw (A2): idx = min(K-1, #{k : cdf_k <= u}), cdf accumulated in class order; returns 1 if this lane's class is drawn
__device__ __forceinline__ float draw_onehot(float prob, float u, int K, int lane) {
    const int k = lane & (K - 1), g0 = lane & ~(K - 1);
    float cdf = 0.f;
    for (int i = 0; i < K; ++i) {
        const float pi = __shfl_sync(FULL, prob, g0 + i);
        if (i <= k) cdf += pi;
    }
    const unsigned hits = __ballot_sync(FULL, (k < K - 1) && (cdf <= u));
    const unsigned gmask = (K == 32 ? 0xffffffffu : ((1u << K) - 1u)) << g0;
    const int idx = __popc(hits & gmask);
    return k == idx ? 1.f : 0.f;
}

__global__ void __launch_bounds__(NTHREADS, 1) mrssm_wide_fwd_kernel(const MrssmWideFwdArgs p) {
    extern __shared__ unsigned char smem_dyn[];
    const int D = p.D, KC = D >> 6, NSL = p.NSL, A = p.A, T = p.T, K = p.K, C = 16 / p.K, F = D + 16;
    const FwdSmem sm = carve(smem_dyn, D, A);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bb = blockIdx.x / NSL, s = blockIdx.x - bb * NSL;
    const bool imagine = p.imagine != 0;

    // ---- one-time setup ---------------------------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) mbar_init(&sm.full[i], 1), mbar_init(&sm.empty[i], p.cs);  // every CTA of the cluster releases a slot
        mbar_init(sm.accbar, N_ISSUERS), mbar_init(sm.firstbar, 1);  // every issuer commits the accumulator barrier
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // first projector layer, transposed: w1t[a][j] = asp_w1[j][a]; row A+16 = bias
    for (int i = tid; i < (A + 16) * D; i += NTHREADS) {
        const int j = i / (A + 16), a = i - j * (A + 16);
        sm.w1t[a * D + j] = p.w.asp_w1[i];
    }
    for (int j = tid; j < D; j += NTHREADS) sm.w1t[(A + 16) * D + j] = p.w.asp_b1[j];
    if (tid < 32) sm.b2s[tid] = p.w.asp_b2[s * 32 + tid];
    if (tid < 96) {
        const int part = tid >> 5, u = tid & 31;
        sm.bih[tid] = p.w.b_ih[part * D + s * 32 + u];
        sm.bhh[tid] = p.w.b_hh[part * D + s * 32 + u];
        const float* b1 = part == 0 ? p.w.pr_b1 : (part == 1 ? p.w.au_b1 : p.w.vi_b1);
        sm.bhd[tid] = (b1 != nullptr) ? b1[s * 32 + u] : 0.f;
    }
    for (int i = tid; i < 3 * 32 * 16; i += NTHREADS) {  // w2l[h][j][o] = W2_h[o][32 s + j]
        const int h = i / 512, j = (i >> 4) & 31, o = i & 15;
        const float* w2 = h == 0 ? p.w.pr_w2 : (h == 1 ? p.w.au_w2 : p.w.vi_w2);
        sm.w2l[i] = (w2 != nullptr) ? w2[o * D + s * 32 + j] : 0.f;
    }
    if (tid < 48) {
        const int h = tid >> 4, o = tid & 15;
        const float* b2 = h == 0 ? p.w.pr_b2 : (h == 1 ? p.w.au_b2 : p.w.vi_b2);
        sm.b2l[tid] = (b2 != nullptr) ? b2[o] : 0.f;
    }
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *sm.tmem_base;
    const uint32_t crank = p.cs > 1 ? cluster_rank() : 0;
    const uint16_t cmask = (uint16_t)((1u << p.cs) - 1u);
    if (p.cs > 1) cluster_sync_all();  // every CTA's barriers are initialised before any remote arrive / multicast copy

    // phase-D rows of this CTA: the rows of its own batch block are dealt out over the block's slice CTAs
    const int rpc = (BM + NSL - 1) / NSL;
    const int row0 = bb * BM + s * rpc;
    const int nrows = max(0, min(min(rpc, BM - s * rpc), p.B - row0));
    const int FG = D >> 3;

    // hid1 of step tn from the previous step's stochastic state and the action of step tn (staged in sm.acts).
    // onehot: the state is a drawn sample (one index per group in sm.zidx) -> C gathered columns; else generic values in sm.zval.
    auto compute_hid1 = [&](int tn, bool onehot) {
        __nv_bfloat16* dst = p.rec + (long long)tn * p.t_stride + (long long)P_HID1 * p.plane_stride;
        for (int item = tid; item < nrows * FG; item += NTHREADS) {
            const int fg = item / nrows, rl = item - fg * nrows, row = row0 + rl;
            float acc[8];
            const float* wb = sm.w1t + fg * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = wb[(A + 16) * D + i];
            for (int a = 0; a < A; ++a) {
                const float x = sm.acts[rl * 8 + a];
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(x, wb[a * D + i], acc[i]);
            }
            if (onehot) {
                for (int g = 0; g < C; ++g) {
                    const float* wz = wb + (A + g * K + sm.zidx[rl * 8 + g]) * D;
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] += wz[i];
                }
            } else {
#pragma unroll 4
                for (int c = 0; c < 16; ++c) {
                    const float z = sm.zval[rl * 16 + c];
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = fmaf(z, wb[(A + c) * D + i], acc[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = M::elu(acc[i]);
            *reinterpret_cast<uint4*>(dst + pk_off(row >> 7, row & 127, fg * 8, D)) = pack8(acc);
        }
    };
    auto stage_actions = [&](int tn) {
        for (int i = tid; i < nrows * A; i += NTHREADS) {
            const int rl = i / A, a = i - rl * A;
            sm.acts[rl * 8 + a] = p.actions[((long long)(row0 + rl) * T + tn) * A + a];
        }
    };

    // ---- prologue: z0 -> zval, hid1_0, packed h0 ---------------------------------------------------------------------------
    for (int i = tid; i < nrows * 16; i += NTHREADS) sm.zval[i] = p.z0[(long long)(row0 + (i >> 4)) * 16 + (i & 15)];
    for (int item = tid; item < nrows * FG; item += NTHREADS) {
        const int fg = item / nrows, rl = item - fg * nrows, row = row0 + rl;
        float v[8];
        const float4 a = *reinterpret_cast<const float4*>(p.h0 + (long long)row * D + fg * 8);
        const float4 b = *reinterpret_cast<const float4*>(p.h0 + (long long)row * D + fg * 8 + 4);
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
        *reinterpret_cast<uint4*>(p.h0p + pk_off(row >> 7, row & 127, fg * 8, D)) = pack8(v);
    }
    stage_actions(0);
    {   // this CTA's own slice of h_{t-1} stays in shared memory for the whole rollout
        const int r = tid;
        if (r < BM) {
            const int gr = bb * BM + r;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gr < p.B) x = *reinterpret_cast<const float4*>(p.h0 + (long long)gr * D + s * 32 + i);
                *reinterpret_cast<float4*>(sm.hs + r * HROW + i) = x;
            }
        }
    }
    __syncthreads();
    compute_hid1(0, false);
    unsigned epoch = 0;
    grid_sync(p.bar + bb * 64, epoch, p.status, NSL);

    Ring ring;            // producer and MMA issuer walk the same chunk sequence
    uint32_t accph = 0;   // epilogue warps: parity of the accumulator barrier
    const uint32_t lboA = BM * 16, sbo = 128;
    const long long blk = (long long)bb * D * BM;  // element offset of block bb inside a plane

    auto load = [&](const __nv_bfloat16* a_src, const __nv_bfloat16* b_src, uint32_t b_bytes) {
        mbar_wait(&sm.empty[ring.slot], ring.phase ^ 1);
        unsigned char* st = sm.ring + (size_t)ring.slot * STAGE_BYTES;
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
        ring.advance(STAGES);
    };
    const int issuer = warp == MMA_WARP ? 0 : 1;  // (meaningful in the issuer warps only)
    // chunk c of an accumulation group: issuer c % N_ISSUERS accumulates it into its own copy of the accumulator
    auto mma_chunk = [&](uint32_t tcol, int N, int c) {
        if (N_ISSUERS > 1 && (c & 1) != issuer) {
            ring.advance(STAGES);
            return;
        }
        mbar_wait(&sm.full[ring.slot], ring.phase);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sm.ring + (size_t)ring.slot * STAGE_BYTES), b0 = a0 + A_BYTES;
        const uint32_t lboB = N * 16;
        const uint32_t idesc = idesc_bf16(N, 0, 0);
        const uint32_t dcol = tmem + tcol + (N_ISSUERS > 1 ? issuer * FWD_ACC_OFF : 0);
        const bool first = c < N_ISSUERS;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = p.desc_swap ? smem_desc(a0 + kk * 2 * lboA, sbo, lboA) : smem_desc(a0 + kk * 2 * lboA, lboA, sbo);
            const uint64_t db = p.desc_swap ? smem_desc(b0 + kk * 2 * lboB, sbo, lboB) : smem_desc(b0 + kk * 2 * lboB, lboB, sbo);
            if (p.exp != 2) umma(dcol, da, db, idesc, (first && kk == 0) ? 0u : 1u);
        }
        if (p.cs == 1) umma_commit(&sm.empty[ring.slot]);
        else umma_commit_mc(&sm.empty[ring.slot], cmask);
        ring.advance(STAGES);
    };
    const int row = (warp & 3) * 32 + lane;               // epilogue: batch row inside the block (TMEM lane quadrant = warp & 3)
    const int half = (warp >> 2) & 1;                      // epilogue: warps 0-3 take column groups 0-1 of the slice, warps 4-7 groups 2-3
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int grow = bb * BM + row;                        // row inside the launch group
    const bool rvalid = grow < p.B;

    int tidx = 0;
    auto stamp = [&]() {
        if (p.timing != nullptr && blockIdx.x == 0 && tid == 0 && tidx < 500) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            p.timing[tidx++] = now;
        }
    };
    stamp();
    for (int t = 0; t < T; ++t) {
        __nv_bfloat16* rec_t = p.rec + (long long)t * p.t_stride;
        const __nv_bfloat16* hb_prev = (t == 0) ? p.h0p + blk : (p.rec + (long long)(t - 1) * p.t_stride + (long long)P_HB * p.plane_stride + blk);
        // =================================== phase A ===================================
        if (warp == PRODUCER_WARP) {
            if (lane == 0) {
                const __nv_bfloat16* a_src = rec_t + (long long)P_HID1 * p.plane_stride + blk;
                for (int c = 0; c < KC; ++c) load(a_src + (long long)c * (BM * 64), p.pW2 + ((long long)s * KC + c) * (32 * 64), 32 * 64 * 2);
                for (int c = 0; c < KC; ++c) load(hb_prev + (long long)c * (BM * 64), p.pWhh + ((long long)s * KC + c) * (96 * 64), 96 * 64 * 2);
            }
        } else if (warp == MMA_WARP || (N_ISSUERS > 1 && warp == MMA_WARP2)) {
            if (lane == 0) {
                for (int c = 0; c < KC; ++c) mma_chunk(TM_X2, 32, c);
                umma_commit(sm.accbar);
                for (int c = 0; c < KC; ++c) mma_chunk(TM_GH, 96, c);
            }
        } else {
            mbar_wait(sm.accbar, accph), accph ^= 1;
            tc_fence_after();
            __nv_bfloat16* x2 = rec_t + (long long)P_X2 * p.plane_stride;
#pragma unroll 1
            for (int q = 2 * half; q < 2 * half + 2; ++q) {
                float v[8];
                acc_ld8(tlane + TM_X2 + q * 8, FWD_ACC_OFF, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] += sm.b2s[q * 8 + i];
                *reinterpret_cast<uint4*>(x2 + pk_off(bb, row, s * 32 + q * 8, D)) = pack8(v);
            }
        }
        stamp();
        grid_sync(p.bar + bb * 64, epoch, p.status, NSL);
        stamp();
        // =================================== phase B ===================================
        if (warp == PRODUCER_WARP) {
            if (lane == 0) {
                const __nv_bfloat16* a_src = rec_t + (long long)P_X2 * p.plane_stride + blk;
                for (int c = 0; c < KC; ++c) load(a_src + (long long)c * (BM * 64), p.pWih + ((long long)s * KC + c) * (96 * 64), 96 * 64 * 2);
            }
        } else if (warp == MMA_WARP || (N_ISSUERS > 1 && warp == MMA_WARP2)) {
            if (lane == 0) {
                for (int c = 0; c < KC; ++c) mma_chunk(TM_GI, 96, c);
                umma_commit(sm.accbar);  // covers the gh MMAs of phase A as well
            }
        } else {
            mbar_wait(sm.accbar, accph), accph ^= 1;
            tc_fence_after();
            float* hsr = sm.hs + row * HROW;
            float* hout = p.feature + ((long long)grow * T + t) * F;
#pragma unroll 1
            for (int q = 2 * half; q < 2 * half + 2; ++q) {
                float gi[8], gh[8], r[8], z[8], n[8], hn[8], h[8];
                acc_ld8(tlane + TM_GI + q * 8, FWD_ACC_OFF, gi);
                acc_ld8(tlane + TM_GH + q * 8, FWD_ACC_OFF, gh);
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = M::sigmoid(gi[i] + gh[i] + sm.bih[q * 8 + i] + sm.bhh[q * 8 + i]);
                acc_ld8(tlane + TM_GI + 32 + q * 8, FWD_ACC_OFF, gi);
                acc_ld8(tlane + TM_GH + 32 + q * 8, FWD_ACC_OFF, gh);
#pragma unroll
                for (int i = 0; i < 8; ++i) z[i] = M::sigmoid(gi[i] + gh[i] + sm.bih[32 + q * 8 + i] + sm.bhh[32 + q * 8 + i]);
                acc_ld8(tlane + TM_GI + 64 + q * 8, FWD_ACC_OFF, gi);
                acc_ld8(tlane + TM_GH + 64 + q * 8, FWD_ACC_OFF, gh);
                float hp[8];
                {
                    const float4 a = *reinterpret_cast<const float4*>(hsr + q * 8), b = *reinterpret_cast<const float4*>(hsr + q * 8 + 4);
                    hp[0] = a.x, hp[1] = a.y, hp[2] = a.z, hp[3] = a.w, hp[4] = b.x, hp[5] = b.y, hp[6] = b.z, hp[7] = b.w;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    hn[i] = gh[i] + sm.bhh[64 + q * 8 + i];
                    n[i] = M::tanh(gi[i] + sm.bih[64 + q * 8 + i] + r[i] * hn[i]);
                    h[i] = (1.f - z[i]) * n[i] + z[i] * hp[i];
                }
                *reinterpret_cast<float4*>(hsr + q * 8) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(hsr + q * 8 + 4) = make_float4(h[4], h[5], h[6], h[7]);
                if (rvalid) {
                    *reinterpret_cast<float4*>(hout + s * 32 + q * 8) = make_float4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4*>(hout + s * 32 + q * 8 + 4) = make_float4(h[4], h[5], h[6], h[7]);
                }
                const long long o = pk_off(bb, row, s * 32 + q * 8, D);
                *reinterpret_cast<uint4*>(rec_t + (long long)P_HB * p.plane_stride + o) = pack8(h);
                if (p.t_stride != 0) {  // gate record for the backward
                    *reinterpret_cast<uint4*>(rec_t + (long long)P_R * p.plane_stride + o) = pack8(r);
                    *reinterpret_cast<uint4*>(rec_t + (long long)P_Z * p.plane_stride + o) = pack8(z);
                    *reinterpret_cast<uint4*>(rec_t + (long long)P_N * p.plane_stride + o) = pack8(n);
                    *reinterpret_cast<uint4*>(rec_t + (long long)P_HN * p.plane_stride + o) = pack8(hn);
                }
            }
        }
        stamp();
        grid_sync(p.bar + bb * 64, epoch, p.status, NSL);
        stamp();
        // =================================== phase C ===================================
        if (warp == PRODUCER_WARP) {
            if (lane == 0) {
                const __nv_bfloat16* a_src = rec_t + (long long)P_HB * p.plane_stride + blk;
                for (int c = 0; c < KC; ++c) load(a_src + (long long)c * (BM * 64), p.pWhd + ((long long)s * KC + c) * (96 * 64), 96 * 64 * 2);
                if (!imagine) {
                    load(p.emb_a + (long long)t * p.emb_t_stride + (long long)bb * (BM * 64), p.pWae + (long long)s * (32 * 64), 32 * 64 * 2);
                    load(p.emb_v + (long long)t * p.emb_t_stride + (long long)bb * (BM * 64), p.pWve + (long long)s * (32 * 64), 32 * 64 * 2);
                }
            }
        } else if (warp == MMA_WARP || (N_ISSUERS > 1 && warp == MMA_WARP2)) {
            if (lane == 0) {
                for (int c = 0; c < KC; ++c) mma_chunk(TM_HD, 96, c);
                if (!imagine) {
                    mma_chunk(TM_HD + 32, 32, KC);
                    mma_chunk(TM_HD + 64, 32, KC + 1);
                }
                umma_commit(sm.accbar);
            }
        } else {
            mbar_wait(sm.accbar, accph), accph ^= 1;
            tc_fence_after();
            const int nheads = imagine ? 1 : 3;
            float accs[3][16];
#pragma unroll
            for (int h = 0; h < 3; ++h) {
#pragma unroll
                for (int o = 0; o < 16; ++o) accs[h][o] = 0.f;
                if (h < nheads) {
#pragma unroll 1
                    for (int q = 2 * half; q < 2 * half + 2; ++q) {
                        float v[8];
                        acc_ld8(tlane + TM_HD + h * 32 + q * 8, FWD_ACC_OFF, v);
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = M::elu(v[i] + sm.bhd[h * 32 + q * 8 + i]);
                        if (p.t_stride != 0)
                            *reinterpret_cast<uint4*>(rec_t + (long long)(P_PH + h) * p.plane_stride + pk_off(bb, row, s * 32 + q * 8, D)) = pack8(v);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4* w = reinterpret_cast<const float4*>(sm.w2l + (h * 32 + q * 8 + i) * 16);
#pragma unroll
                            for (int o4 = 0; o4 < 4; ++o4) {
                                const float4 ww = w[o4];
                                accs[h][4 * o4] = fmaf(v[i], ww.x, accs[h][4 * o4]), accs[h][4 * o4 + 1] = fmaf(v[i], ww.y, accs[h][4 * o4 + 1]);
                                accs[h][4 * o4 + 2] = fmaf(v[i], ww.z, accs[h][4 * o4 + 2]), accs[h][4 * o4 + 3] = fmaf(v[i], ww.w, accs[h][4 * o4 + 3]);
                            }
                        }
                    }
                }
            }
            // the two column halves of a row meet in shared memory (the operand ring is idle here: every MMA of the phase is done)
            float* ex = reinterpret_cast<float*>(sm.ring) + row * 52;
            if (half == 1) {
#pragma unroll
                for (int h = 0; h < 3; ++h)
#pragma unroll
                    for (int o4 = 0; o4 < 4; ++o4)
                        *reinterpret_cast<float4*>(ex + h * 16 + o4 * 4) = make_float4(accs[h][4 * o4], accs[h][4 * o4 + 1], accs[h][4 * o4 + 2], accs[h][4 * o4 + 3]);
            }
            epi_sync();
            if (half == 0) {
                float* part = p.part + ((long long)grow * NSL + s) * 48;
#pragma unroll
                for (int h = 0; h < 3; ++h)
#pragma unroll
                    for (int o4 = 0; o4 < 4; ++o4) {
                        const float4 x = *reinterpret_cast<const float4*>(ex + h * 16 + o4 * 4);
                        *reinterpret_cast<float4*>(part + h * 16 + o4 * 4) =
                            make_float4(accs[h][4 * o4] + x.x, accs[h][4 * o4 + 1] + x.y, accs[h][4 * o4 + 2] + x.z, accs[h][4 * o4 + 3] + x.w);
                    }
            }
        }
        stamp();
        grid_sync(p.bar + bb * 64, epoch, p.status, NSL);
        stamp();
        // =================================== phase D ===================================
        {
            if (t + 1 < T) stage_actions(t + 1);
            const int j = lane & 15, g = j / K;
            for (int base = warp * 2; base < nrows; base += (NTHREADS / 32) * 2) {
                const int rl = base + (lane >> 4);
                const bool valid = rl < nrows;
                const int rw = row0 + (valid ? rl : nrows - 1);
                const float* part = p.part + (long long)rw * NSL * 48;
                // all partial-logit loads in flight at once (NSL <= 16; clamped index, masked value); L2 loads: other CTAs wrote them
                float pv[16][3];
#pragma unroll
                for (int sl = 0; sl < 16; ++sl) {
                    const float* ps = part + (sl < NSL ? sl : NSL - 1) * 48 + j;
                    pv[sl][0] = __ldcg(ps);
                    pv[sl][1] = imagine ? 0.f : __ldcg(ps + 16);
                    pv[sl][2] = imagine ? 0.f : __ldcg(ps + 32);
                }
                float lp = sm.b2l[j], la = sm.b2l[16 + j], lv = sm.b2l[32 + j];
#pragma unroll
                for (int sl = 0; sl < 16; ++sl) {
                    const float m = sl < NSL ? 1.f : 0.f;
                    lp = fmaf(m, pv[sl][0], lp), la = fmaf(m, pv[sl][1], la), lv = fmaf(m, pv[sl][2], lv);
                }
                const long long bt = (long long)rw * T + t;
                const float pp = group_softmax(lp, K);
                float zs;  // the sample fed back
                if (!imagine) {
                    // flat log-softmaxes, PoE = sum, MoE = logsumexp of the three experts (1/3 each)
                    const float ma = half_max(la), mv = half_max(lv);
                    const float lsa = la - ma - M::log(half_sum(M::exp(la - ma)));
                    const float lsv = lv - mv - M::log(half_sum(M::exp(lv - mv)));
                    const float f = lsa + lsv, mx = fmaxf(lsa, fmaxf(lsv, f));
                    const float mixed = -1.0986122886681098f + mx + M::log(M::exp(lsa - mx) + M::exp(lsv - mx) + M::exp(f - mx));
                    const float q = group_softmax(mixed, K);
                    const float u = p.u_post[bt * C + g];
                    zs = draw_onehot(q, u, K, lane);
                    const float klt = half_sum(q * (clamp_log<true>(q) - clamp_log<true>(pp)));
                    if (valid) {
                        if (p.logits != nullptr) p.logits[bt * 32 + j] = la, p.logits[bt * 32 + 16 + j] = lv;
                        p.post_probs[bt * 16 + j] = q;
                        if (j == 0) p.kl[bt] = klt;
                    }
                    if (p.u_prior != nullptr) {
                        const float zp = draw_onehot(pp, p.u_prior[bt * C + g], K, lane);
                        if (valid && p.prior_stoch != nullptr) p.prior_stoch[bt * 16 + j] = zp;
                    }
                } else {
                    zs = draw_onehot(pp, p.u_prior[bt * C + g], K, lane);
                }
                if (valid) {
                    p.prior_probs[bt * 16 + j] = pp;
                    p.feature[bt * F + D + j] = zs;
                    if (zs != 0.f) sm.zidx[rl * 8 + g] = j - g * K;
                }
            }
            stamp();
            __syncthreads();
            stamp();
            if (t + 1 < T) compute_hid1(t + 1, true);
        }
        stamp();
        grid_sync(p.bar + bb * 64, epoch, p.status, NSL);
        stamp();
    }

    if (p.cs > 1) {
        // no CTA may leave while a peer can still multicast into its shared memory or arrive on its barriers: take every ring
        // slot once more (= all peers' last commits have arrived here), then meet the cluster
        if (warp == PRODUCER_WARP && lane == 0)
            for (int i = 0; i < STAGES; ++i) mbar_wait(&sm.empty[ring.slot], ring.phase ^ 1), ring.advance(STAGES);
        __syncthreads();
        cluster_sync_all();
    }
    tc_fence_before();
    __syncthreads();
    __syncwarp();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- packing kernels ---------------------------------------------------------------------------------------------------
__global__ void wide_pack_weights_kernel(const WidePackJobs jobs) {
    const WidePackJob& j = jobs.job[blockIdx.y];
    const int N = 32 * j.nparts, KC = j.K >> 6;
    const long long total = (long long)jobs.NSL * KC * 8 * N;  // 16-byte groups
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i % N);
        const int kg = (int)((i / N) % 8);
        const int c = (int)((i / (8LL * N)) % KC);
        const int s = (int)(i / (8LL * N * KC));
        const int part = n >> 5, u = n & 31;
        const float* src = j.src[part] + (long long)(s * 32 + u) * j.ld[part] + j.coloff + c * 64 + kg * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = src[e];
        *reinterpret_cast<uint4*>(j.dst + i * 8) = pack8(v);
    }
}

__global__ void wide_pack_rows_kernel(const float* __restrict__ src, int B, int T, int Fc, int ld, int coloff, __nv_bfloat16* __restrict__ dst,
                                      int blocks) {
    const int FG = Fc >> 3;
    const long long total = (long long)B * T * FG;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int fg = (int)(i % FG);
        const int t = (int)((i / FG) % T);
        const int b = (int)(i / ((long long)FG * T));
        const float* s = src + ((long long)b * T + t) * ld + coloff + fg * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = s[e];
        *reinterpret_cast<uint4*>(dst + ((((long long)t * blocks + (b >> 7)) * FG + fg) * BM + (b & 127)) * 8) = pack8(v);
    }
}

}  // namespace wide

size_t mrssm_wide_fwd_smem(int D) { return wide::fwd_smem_bytes(D, 8); }

// cooperative launch (co-residency for the barriers) of clusters of a.cs CTAs along the slice index
cudaError_t launch_wide_persistent(const void* kernel, void* args, int grid, int cs, size_t smem, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(wide::NTHREADS), cfg.dynamicSmemBytes = smem, cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative, attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cs, attr[1].val.clusterDim.y = 1, attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = cs > 1 ? 2 : 1;
    void* params[] = {args};
    return cudaLaunchKernelExC(&cfg, kernel, params);
}

cudaError_t launch_mrssm_wide_fwd(const MrssmWideFwdArgs& a, cudaStream_t s) {
    MrssmWideFwdArgs args = a;
    return launch_wide_persistent((const void*)wide::mrssm_wide_fwd_kernel, &args, a.NBB * a.NSL, a.cs, wide::fwd_smem_bytes(a.D, a.A), s);
}

cudaError_t launch_wide_pack_weights(const WidePackJobs& jobs, cudaStream_t s) {
    wide::wide_pack_weights_kernel<<<dim3(64, jobs.njobs), 256, 0, s>>>(jobs);
    return cudaGetLastError();
}

cudaError_t launch_wide_pack_rows(const float* src, int B, int T, int F, int ld, int coloff, __nv_bfloat16* dst, int blocks, cudaStream_t s) {
    const long long total = (long long)B * T * (F >> 3);
    const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    wide::wide_pack_rows_kernel<<<grid > 0 ? grid : 1, 256, 0, s>>>(src, B, T, F, ld, coloff, dst, blocks);
    return cudaGetLastError();
}

}  // namespace rssm
