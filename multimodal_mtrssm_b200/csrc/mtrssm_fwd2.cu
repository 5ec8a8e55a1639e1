// MoPoE-MMTRSSM forward rollout, bf16 tensor-core policy: TWO warps per 16-sequence tile.
//
// Why.  clock64 timelines of the one-warp-per-tile forward (profiles/r2_b_timing.txt) show a warp-step of ~7.1 k cycles alone on
// an SM and ~11-16 k cycles with 8 tiles per SM -- ~900 instructions at ~0.1 instructions per cycle per warp: a chain of dependent
// LDS -> HMMA -> MUFU -> shuffle latencies, with only two warps per scheduler to hide it (198 registers per thread).  A step has two
// halves that only meet at the cells:
//     "state" warp (role 0): the two leaky-integrator cells, the HIGHER prior and posterior heads, the draw of z_h, KL_h
//     "obs"   warp (role 1): audio + vision heads on d_l, MoPoE fusion, the draw of z_l, then (off the recurrence) the LOWER
//                            prior head, its draw and KL_l
// They exchange d_l (state -> obs, 1 KB of ready-made A fragments) and z_l (obs -> state, 512 B) through shared memory under two
// named barriers per tile (producer bar.arrive, consumer bar.sync; 8 tiles use all 16 hardware barriers -- id 0 is free after the
// prologue's __syncthreads; ONE shared id would be wrong: the producer's non-blocking arrive lets it reach its own consumer-side
// sync first and complete the barrier alone).  Each warp needs ~half the registers, 16 warps fit an SM, and the per-step
// critical path drops from (cells + all heads) to (cells + modality heads + fusion).
//
// Same arithmetic and operand roundings as mtrssm_fwd_kernel<1, KL, KH, false>; one fp32 summation order differs (the embedding
// half of the modality heads' first layer is accumulated BEFORE the d_l half, so that it runs ahead of the hand-over), i.e. the two
// kernels agree to fp32 rounding of bf16-operand dot products, not bit for bit (tests/test_rollout_gpu.py::
// test_mtrssm_fwd2_matches_one_warp_kernel; RSSM_FWD_ONE_WARP=1 selects the one-warp kernel).
//
// Reference semantics: MoPoE_MMTRSSM.rollout_representation, mmtrssm/mopoe_mmtrssm/core.py:364-494 (see mtrssm_kernels.cu).
#include <stdlib.h>

#include "frag.cuh"
#include "kernels.h"
#include "mtrssm_common.cuh"

namespace rssm {

namespace f2 {
constexpr int XD = 0;                       // state -> obs: d_l as A fragments  [2 k-tiles][4 regs][32 lanes] u32
constexpr int XZ = XD + 2 * 4 * 32 * 4;     // obs -> state: z_l as an A fragment [4 regs][32 lanes] u32
constexpr int XBYTES = XZ + 4 * 32 * 4;     // 1536
constexpr int TILE_BYTES = 2 * stg::FLOATS * 4 + XBYTES;  // double-buffered input stage + exchange: 20,992
}  // namespace f2

__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void pair_arrive(int id) {
    __threadfence_block();
    asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory");
}

// uniforms of ONE draw site for this lane's two rows, fetched one step ahead into registers:
// K == 2: u[2t], u[2t+1];  K >= 4: u[t / (K/4)] in [0]
template <int K>
__device__ __forceinline__ void fetch_uniforms(const float* __restrict__ u, size_t iA, size_t iB, int lane, float (&uA)[2], float (&uB)[2]) {
    constexpr int C = 16 / K;
    const int t = lane & 3;
    if constexpr (K == 2) {
        const float2 a = *reinterpret_cast<const float2*>(u + iA * C + 2 * t), b = *reinterpret_cast<const float2*>(u + iB * C + 2 * t);
        uA[0] = a.x, uA[1] = a.y, uB[0] = b.x, uB[1] = b.y;
    } else {
        constexpr int LANES = K / 4;
        uA[0] = u[iA * C + t / LANES], uB[0] = u[iB * C + t / LANES];
        uA[1] = uB[1] = 0.f;
    }
}

// sample_onehot (frag.cuh) with the uniforms already in registers (fetch_uniforms)
template <int K>
__device__ __forceinline__ void sample_onehot_regs(const float (&p)[2][4], const float (&uA)[2], const float (&uB)[2], float (&z)[2][4],
                                                   int lane) {
    const int t = lane & 3, qbase = lane & ~3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float mine[4] = {p[0][2 * h], p[0][2 * h + 1], p[1][2 * h], p[1][2 * h + 1]};
        int hit[4];
        if constexpr (K == 2) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const float uu = h == 0 ? uA[half] : uB[half];
                const int idx = mine[2 * half] <= uu ? 1 : 0;
                hit[2 * half] = idx == 0, hit[2 * half + 1] = idx == 1;
            }
        } else {
            constexpr int LANES = K / 4;
            const int first = t & ~(LANES - 1), pos = 4 * (t & (LANES - 1));
            const float uu = h == 0 ? uA[0] : uB[0];
            float cdf = 0.f;
            int idx = 0;
#pragma unroll
            for (int s = 0; s < LANES; ++s) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float pk = mine[i];
                    if constexpr (LANES > 1) pk = __shfl_sync(FULL, pk, qbase + first + s);
                    cdf += pk;
                    if (4 * s + i < K - 1) idx += (cdf <= uu) ? 1 : 0;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) hit[i] = idx == pos + i;
        }
        z[0][2 * h] = hit[0] ? 1.f : 0.f, z[0][2 * h + 1] = hit[1] ? 1.f : 0.f;
        z[1][2 * h] = hit[2] ? 1.f : 0.f, z[1][2 * h + 1] = hit[3] ? 1.f : 0.f;
    }
}

// obs_projected: the two modality partials P = e . W1[:, 32:]^T, [B,T,32] fp32, staged like the embeddings (regions EA / EV of the
// stage, rows of 128 B, 16-byte chunks XOR-swizzled by (row & 7) so that the C-tile read pattern below is conflict-free) + u0
__device__ __forceinline__ void stage_projected(float* stage, const float* __restrict__ pa, const float* __restrict__ pv,
                                                const float* __restrict__ u0, int n0, int row0, int B, int T, int t, int lane) {
    const int chunk = lane & 7;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rl = (lane >> 3) + 4 * i;
        const size_t src = ((size_t)min(row0 + rl, B - 1) * T + t) * 32 + chunk * 4;
        const int dst = rl * 32 + ((chunk ^ (rl & 7)) << 2);
        cp_async16(stage + stg::EA + dst, pa + src);
        cp_async16(stage + stg::EV + dst, pv + src);
    }
    stage_inputs(stage, nullptr, nullptr, nullptr, 0, u0, n0, nullptr, 0, row0, B, T, t, lane);  // the uniforms (+ commit)
}
// acc (C tiles, 32 columns) += the staged partial rows of this lane
__device__ __forceinline__ void add_staged_projected(float (&acc)[4][4], const float* tile, int g, int t) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int c = ((4 * j + t) ^ (g & 7)) << 2;
        const float4 a = *reinterpret_cast<const float4*>(tile + g * 32 + c), b = *reinterpret_cast<const float4*>(tile + (g + 8) * 32 + c);
        acc[2 * j][0] += a.x, acc[2 * j][1] += a.y, acc[2 * j + 1][0] += a.z, acc[2 * j + 1][1] += a.w;
        acc[2 * j][2] += b.x, acc[2 * j][3] += b.y, acc[2 * j + 1][2] += b.z, acc[2 * j + 1][3] += b.w;
    }
}

// Where a step's saved record goes.  Row layout: [B][T][saved_ld] (A / B = this lane's two rows).  Tile-blocked layout
// (RSSM_PRECISION_BF16_FUSED): `tile` = the tile-step's contiguous block [26 chunks][16 rows][8 bf16] -- exactly the tcgen05 operand
// image the fused backward wants, so it stages a head's hiddens with ONE bulk copy; a warp's record stores of one step land in
// one 6.5 KB block instead of 16 rows x 5 segments scattered at a T x 416-byte stride.  All 16 rows are written (the pad rows of
// the last tile hold copies of the last real row: they meet zero dY rows in the weight-gradient MMAs and must stay finite).
struct RecDst {
    __nv_bfloat16 *A, *B, *tile;
    __device__ __forceinline__ explicit operator bool() const { return A != nullptr || tile != nullptr; }
};
template <int NT, bool TILED>
__device__ __forceinline__ void store_rec_dst(const float (&c)[NT][4], const RecDst& d, int col0, const Rows& r) {
    if constexpr (TILED) {
#pragma unroll
        for (int j = 0; j < NT / 2; ++j) {
            __nv_bfloat16* q = d.tile + ((col0 + 16 * j) / 8 + (r.t >> 1)) * 128 + r.g * 8 + (r.t & 1) * 4;
            *reinterpret_cast<uint2*>(q) = make_uint2(pack_bf16(c[2 * j][0], c[2 * j][1]), pack_bf16(c[2 * j + 1][0], c[2 * j + 1][1]));
            *reinterpret_cast<uint2*>(q + 64) = make_uint2(pack_bf16(c[2 * j][2], c[2 * j][3]), pack_bf16(c[2 * j + 1][2], c[2 * j + 1][3]));
        }
    } else {
        store_rec<NT>(c, d.A + col0, d.B + col0, r);
    }
}

// hidden -> ELU -> (saved) -> logits, as head_l2 of mtrssm_kernels.cu (NS = 1)
template <bool TILED, bool PAIRED>
__device__ __forceinline__ void head2_l2(float (&acc)[4][4], float (&logits)[2][4], const float* bias2, const uint2* w2, const RecDst& sv,
                                         int sv_off, const Rows& r, int lane) {
    map_c<4>(acc, EluOp<true>{});
    if (sv) store_rec_dst<4, TILED>(acc, sv, sv_off, r);
    AFrag<1, 2> f1;
    to_afrag<1, 2>(f1, acc);
    init_bias<2>(logits, bias2, r.t);
    gemm<1, 2, 2, PAIRED>(logits, f1, w2, lane);
}

// blockDim.x = 64 * (tiles per CTA), 1 .. 8 tiles: small batches run few tiles per CTA to reach more SMs
// LAYOUT: 0 = row-layout record, dense outputs (bf16 two-kernel policy); 1 = tile-blocked record, dense outputs; 2 = tile-blocked
// record, GROUPED outputs (one 256-float row per (b,t), RssmMtrssmOutputs.ld_*).  The pitches are compile-time constants: runtime
// pitches cost 3-4% on the latency-bound one-tile rollouts (64-bit IMADs in front of every store; profiles/r2_p_grouped_rows_ab.txt)
// PAIRED: B fragments fetched in n-tile pairs (frag.cuh, bfrag_slot): faster when several tiles share an SM, slower on a lone tile's chain
template <int KL, int KH, int LAYOUT, bool PAIRED>
__global__ void __launch_bounds__(512, 1) mtrssm_fwd2_kernel(const MtrssmFwdArgs p) {
    constexpr bool TILED = LAYOUT >= 1, GROUPED = LAYOUT == 2;
    constexpr int NS = 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2* W = reinterpret_cast<uint2*>(smem_raw);
    float* bias = reinterpret_cast<float*>(W + (size_t)mt::FWD_TILES * 32);
    unsigned char* tiles = reinterpret_cast<unsigned char*>(bias + mt::FWD_BIAS);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int A = p.A;
    __shared__ PackTable tb;
    if (tid == 0) {
        using namespace mt;
        const int ldin = A + 32;
        tb.nblocks = tb.ntiles = 0;
        pack_add(tb, false, wblk<NS>(W, L_D2H), p.w.l_d2h_w, 32, 0, 0, 32, 32, 2, 4, PAIRED);
        pack_add(tb, false, wblk<NS>(W, L_IN_ZL), p.w.l_in_w, ldin, 0, A, 16, 32, 1, 4, PAIRED);
        pack_add(tb, false, wblk<NS>(W, L_IN_ZH), p.w.l_in_w, ldin, 0, A + 16, 16, 32, 1, 4, PAIRED);
        pack_add(tb, false, wblk<NS>(W, L_IN_A), p.w.l_in_w, ldin, 0, 0, A, 32, 1, 4, PAIRED);
        pack_add(tb, false, wblk<NS>(W, H_D2H), p.w.h_d2h_w, 32, 0, 0, 32, 32, 2, 4, PAIRED);
        pack_add(tb, false, wblk<NS>(W, H_IN), p.w.h_in_w, 16, 0, 0, 16, 32, 1, 4, PAIRED);
        pack_add(tb, false, wblk<NS>(W, LP1), p.w.lp_w1, 32, 0, 0, 32, 32, 2, 4, PAIRED);
        pack_add(tb, false, wblk<NS>(W, LP2), p.w.lp_w2, 32, 0, 0, 32, 16, 2, 2, PAIRED);
        pack_add(tb, false, wblk<NS>(W, HP1), p.w.hp_w1, 32, 0, 0, 32, 32, 2, 4, PAIRED);
        pack_add(tb, false, wblk<NS>(W, HP2), p.w.hp_w2, 32, 0, 0, 32, 16, 2, 2, PAIRED);
        pack_add(tb, false, wblk<NS>(W, HQ1L), p.w.hq_w1, 64, 0, 0, 32, 32, 2, 4, PAIRED);
        pack_add(tb, false, wblk<NS>(W, HQ1H), p.w.hq_w1, 64, 0, 32, 32, 32, 2, 4, PAIRED);
        pack_add(tb, false, wblk<NS>(W, HQ2), p.w.hq_w2, 32, 0, 0, 32, 16, 2, 2, PAIRED);
        pack_add(tb, false, wblk<NS>(W, A1H), p.w.au_w1, 96, 0, 0, 32, 32, 2, 4, PAIRED);
        pack_add(tb, false, wblk<NS>(W, A1E), p.w.au_w1, 96, 0, 32, 64, 32, 4, 4, PAIRED);
        pack_add(tb, false, wblk<NS>(W, A2), p.w.au_w2, 32, 0, 0, 32, 16, 2, 2, PAIRED);
        pack_add(tb, false, wblk<NS>(W, V1H), p.w.vi_w1, 96, 0, 0, 32, 32, 2, 4, PAIRED);
        pack_add(tb, false, wblk<NS>(W, V1E), p.w.vi_w1, 96, 0, 32, 64, 32, 4, 4, PAIRED);
        pack_add(tb, false, wblk<NS>(W, V2), p.w.vi_w2, 32, 0, 0, 32, 16, 2, 2, PAIRED);
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

    const int lane = tid & 31, warp = tid >> 5, pair = warp >> 1, role = warp & 1;
    const int T = p.T, ntiles = (p.B + 15) / 16;
    constexpr int CL = 16 / KL, CH = 16 / KH, F = 96;
    unsigned char* my = tiles + (size_t)pair * f2::TILE_BYTES;
    float* stage_base = reinterpret_cast<float*>(my);
    uint32_t* xd = reinterpret_cast<uint32_t*>(my + 2 * stg::FLOATS * 4 + f2::XD);
    uint32_t* xz = reinterpret_cast<uint32_t*>(my + 2 * stg::FLOATS * 4 + f2::XZ);
    const int TPC = nthr >> 6;  // <= 8: two named barriers per tile, 16 hardware barriers
    const int bar_d = 2 * pair, bar_z = 2 * pair + 1;  // d_l: state -> obs;  z_l: obs -> state
    __nv_bfloat16* saved = reinterpret_cast<__nv_bfloat16*>(p.saved);
    const bool prior_draws = p.u_prior_l != nullptr;
    constexpr int ldF = GROUPED ? MTRSSM_ROW_PITCH : F, ldH = GROUPED ? MTRSSM_ROW_PITCH : 32, ldP = GROUPED ? MTRSSM_ROW_PITCH : 16,
                  ldS = GROUPED ? MTRSSM_ROW_PITCH : 16, ldK = GROUPED ? 2 : 1;

    // persistent over tiles: a warp pair walks tiles blockIdx.x * TPC + pair, + gridDim.x * TPC, ... (tiles are independent)
    for (int tile = blockIdx.x * TPC + pair; tile < ntiles; tile += gridDim.x * TPC) {
        const int row0 = tile * 16;
        const Rows r = make_rows(row0, p.B, lane);
        // the action pad columns of the stage stay zero; each warp owns disjoint regions of the two stage buffers
        if (role == 0) {
            for (int i = lane; i < 2 * 128; i += 32) stage_base[(i >> 7) * stg::FLOATS + stg::ACT + (i & 127)] = 0.f;
            __syncwarp();
        }
        if (role == 0) {
            // =========================== state warp: cells, higher prior / posterior, z_h ===========================
            const float keep_l = 1.f - p.inv_tau_l, keep_h = 1.f - p.inv_tau_h;
            stage_inputs(stage_base, nullptr, nullptr, p.actions, A, nullptr, 0, p.u_post_h, CH, row0, p.B, T, 0, lane);
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
            float upA[2] = {0.f, 0.f}, upB[2] = {0.f, 0.f};  // uniforms of the prior's own z_h draw, one step ahead
            if (prior_draws) fetch_uniforms<KH>(p.u_prior_h, (size_t)r.rA * T, (size_t)r.rB * T, lane, upA, upB);
            for (int t = 0; t < T; ++t) {
                const size_t iA = (size_t)r.rA * T + t, iB = (size_t)r.rB * T + t;
                RecDst sv{nullptr, nullptr, nullptr};
                if (saved != nullptr) {
                    if constexpr (TILED) sv.tile = saved + ((size_t)tile * T + t) * (MTRSSM_SAVED_BF16 * 16);
                    else sv.A = saved + iA * p.saved_ld, sv.B = saved + iB * p.saved_ld;
                }
                const float* stage = stage_base + (t & 1) * stg::FLOATS;
                cp_async_wait_all();
                __syncwarp();
                if (t + 1 < T)
                    stage_inputs(stage_base + ((t + 1) & 1) * stg::FLOATS, nullptr, nullptr, p.actions, A, nullptr, 0, p.u_post_h, CH, row0, p.B,
                                 T, t + 1, lane);
                float pl[4][4], ph[4][4];
                AFrag<NS, 1> fa;
                load_a_staged_act<NS>(fa, stage + stg::ACT, r.g, r.t);
                // everything of the two cells that does not need z_l(t-1) ...
                init_bias<4>(pl, bias + mt::B_L, r.t);
                gemm<NS, 2, 4, PAIRED>(pl, dlf, wblk<NS>(W, mt::L_D2H), lane);
                init_bias<4>(ph, bias + mt::B_H, r.t);
                gemm<NS, 2, 4, PAIRED>(ph, dhf, wblk<NS>(W, mt::H_D2H), lane);
                if (t > 0) {  // ... then z_l of the previous step from the obs warp
                    pair_sync(bar_z);
#pragma unroll
                    for (int i = 0; i < 4; ++i) zlf.r[0][0][i] = xz[i * 32 + lane];
                }
                gemm<NS, 1, 4, PAIRED>(pl, zlf, wblk<NS>(W, mt::L_IN_ZL), lane);
                gemm<NS, 1, 4, PAIRED>(pl, zhf, wblk<NS>(W, mt::L_IN_ZH), lane);
                gemm<NS, 1, 4, PAIRED>(pl, fa, wblk<NS>(W, mt::L_IN_A), lane);
                gemm<NS, 1, 4, PAIRED>(ph, zhf, wblk<NS>(W, mt::H_IN), lane);
                float dl[4][4], dh[4][4];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        ul[nt][j] = keep_l * ul[nt][j] + pl[nt][j] * p.inv_tau_l;
                        dl[nt][j] = Math<true>::tanh(ul[nt][j]);
                    }
                to_afrag<NS, 2>(dlf, dl);
#pragma unroll
                for (int kt = 0; kt < 2; ++kt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) xd[(kt * 4 + i) * 32 + lane] = dlf.r[0][kt][i];
                pair_arrive(bar_d);  // d_l(t) is in XD: the obs warp starts its heads
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uh[nt][j] = keep_h * uh[nt][j] + ph[nt][j] * p.inv_tau_h;
                        dh[nt][j] = Math<true>::tanh(uh[nt][j]);
                    }
                to_afrag<NS, 2>(dhf, dh);
                store_c<4>(dh, p.feature + iA * ldF, p.feature + iB * ldF, r);
                store_c<4>(dl, p.feature + iA * ldF + 48, p.feature + iB * ldF + 48, r);
                store_c<4>(uh, p.hidden_h + iA * ldH, p.hidden_h + iB * ldH, r);
                store_c<4>(ul, p.hidden_l + iA * ldH, p.hidden_l + iB * ldH, r);
                // ---- higher prior (:311-312) ------------------------------------------------------------------------------
                float pph[2][4];
                {
                    float acc[4][4], lg[2][4];
                    init_bias<4>(acc, bias + mt::B_HP1, r.t);
                    gemm<NS, 2, 4, PAIRED>(acc, dhf, wblk<NS>(W, mt::HP1), lane);
                    head2_l2<TILED, PAIRED>(acc, lg, bias + mt::B_HP2, wblk<NS>(W, mt::HP2), sv, mts::HP_HID, r, lane);
                    softmax_groups<KH, true>(lg, pph);
                    store_c<2>(pph, p.prior_probs_h + iA * ldP, p.prior_probs_h + iB * ldP, r);
                }
                if (prior_draws) {  // the prior MTState's own draw (state.py:48)
                    float zh[2][4];
                    sample_onehot_regs<KH>(pph, upA, upB, zh, lane);
                    if (p.prior_stoch_h != nullptr) store_c<2>(zh, p.prior_stoch_h + iA * ldS, p.prior_stoch_h + iB * ldS, r);
                    if (t + 1 < T) fetch_uniforms<KH>(p.u_prior_h, iA + 1, iB + 1, lane, upA, upB);
                }
                // ---- higher posterior on [d_l ; d_h] (:315-317), sample (:464) -------------------------------------------
                {
                    float acc[4][4], lg[2][4], q[2][4], zs[2][4];
                    init_bias<4>(acc, bias + mt::B_HQ1, r.t);
                    gemm<NS, 2, 4, PAIRED>(acc, dlf, wblk<NS>(W, mt::HQ1L), lane);
                    gemm<NS, 2, 4, PAIRED>(acc, dhf, wblk<NS>(W, mt::HQ1H), lane);
                    head2_l2<TILED, PAIRED>(acc, lg, bias + mt::B_HQ2, wblk<NS>(W, mt::HQ2), sv, mts::HQ_HID, r, lane);
                    softmax_groups<KH, true>(lg, q);
                    store_c<2>(q, p.post_probs_h + iA * ldP, p.post_probs_h + iB * ldP, r);
                    sample_onehot<KH>(q, stage + stg::U1 + r.g * 8, stage + stg::U1 + (r.g + 8) * 8, zs, lane);
                    store_c<2>(zs, p.feature + iA * ldF + 32, p.feature + iB * ldF + 32, r);
                    to_afrag<NS, 1>(zhf, zs);
                    float kl[2];
                    kl_rows<true>(q, pph, kl);
                    if (r.t == 0) {
                        if (r.vA) p.kl_h[iA * ldK] = kl[0];
                        if (r.vB) p.kl_h[iB * ldK] = kl[1];
                    }
                }
            }
            pair_sync(bar_z);  // consume the obs warp's last hand-over (keeps the barrier's phases paired across tiles)
        } else {
            // =========================== obs warp: modality heads, MoPoE, z_l, lower prior ===========================
            if (p.obs_projected) stage_projected(stage_base, p.embed_a, p.embed_v, p.u_post_l, CL, row0, p.B, T, 0, lane);
            else stage_inputs(stage_base, p.embed_a, p.embed_v, nullptr, A, p.u_post_l, CL, nullptr, 0, row0, p.B, T, 0, lane);
            float upA[2] = {0.f, 0.f}, upB[2] = {0.f, 0.f};  // uniforms of the prior's own z_l draw, one step ahead
            if (prior_draws) fetch_uniforms<KL>(p.u_prior_l, (size_t)r.rA * T, (size_t)r.rB * T, lane, upA, upB);
            for (int t = 0; t < T; ++t) {
                const size_t iA = (size_t)r.rA * T + t, iB = (size_t)r.rB * T + t;
                RecDst sv{nullptr, nullptr, nullptr};
                if (saved != nullptr) {
                    if constexpr (TILED) sv.tile = saved + ((size_t)tile * T + t) * (MTRSSM_SAVED_BF16 * 16);
                    else sv.A = saved + iA * p.saved_ld, sv.B = saved + iB * p.saved_ld;
                }
                const float* stage = stage_base + (t & 1) * stg::FLOATS;
                cp_async_wait_all();
                __syncwarp();
                if (t + 1 < T) {
                    float* nxt = stage_base + ((t + 1) & 1) * stg::FLOATS;
                    if (p.obs_projected) stage_projected(nxt, p.embed_a, p.embed_v, p.u_post_l, CL, row0, p.B, T, t + 1, lane);
                    else stage_inputs(nxt, p.embed_a, p.embed_v, nullptr, A, p.u_post_l, CL, nullptr, 0, row0, p.B, T, t + 1, lane);
                }
                // the embedding halves of the two first layers do not need d_l: they run before the hand-over
                float acca[4][4], accv[4][4];
                init_bias<4>(acca, bias + mt::B_A1, r.t);
                init_bias<4>(accv, bias + mt::B_V1, r.t);
                if (p.obs_projected) {  // the caller multiplied the embeddings by W1[:, 32:] in one GEMM before the loop
                    add_staged_projected(acca, stage + stg::EA, r.g, r.t);
                    add_staged_projected(accv, stage + stg::EV, r.g, r.t);
                } else {
                    AFrag<NS, 4> fe;
                    load_a_staged64<NS>(fe, stage + stg::EA, r.g, r.t);
                    gemm<NS, 4, 4, PAIRED>(acca, fe, wblk<NS>(W, mt::A1E), lane);
                    load_a_staged64<NS>(fe, stage + stg::EV, r.g, r.t);
                    gemm<NS, 4, 4, PAIRED>(accv, fe, wblk<NS>(W, mt::V1E), lane);
                }
                AFrag<NS, 2> dlf;
                pair_sync(bar_d);  // d_l(t) is in XD
#pragma unroll
                for (int kt = 0; kt < 2; ++kt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) dlf.r[0][kt][i] = xd[(kt * 4 + i) * 32 + lane];
                // ---- lower posterior: modality heads on d_l (:422-433), MoPoE fusion (:436-455), sample (:456) ----------
                float la[2][4], lv[2][4];
                gemm<NS, 2, 4, PAIRED>(acca, dlf, wblk<NS>(W, mt::A1H), lane);
                head2_l2<TILED, PAIRED>(acca, la, bias + mt::B_A2, wblk<NS>(W, mt::A2), sv, mts::A_HID, r, lane);
                gemm<NS, 2, 4, PAIRED>(accv, dlf, wblk<NS>(W, mt::V1H), lane);
                head2_l2<TILED, PAIRED>(accv, lv, bias + mt::B_V2, wblk<NS>(W, mt::V2), sv, mts::V_HID, r, lane);
                float q[2][4];
                {
                    float zs[2][4];
                    mopoe_posterior_fast<KL>(la, lv, q);  // probability-domain MoPoE (frag.cuh): q = s / sum_group(s)
                    sample_onehot<KL>(q, stage + stg::U0 + r.g * 8, stage + stg::U0 + (r.g + 8) * 8, zs, lane);
                    AFrag<NS, 1> zlf;
                    to_afrag<NS, 1>(zlf, zs);
#pragma unroll
                    for (int i = 0; i < 4; ++i) xz[i * 32 + lane] = zlf.r[0][0][i];
                    pair_arrive(bar_z);  // z_l(t) is in XZ: the state warp's next cells may run
                    store_c<2>(zs, p.feature + iA * ldF + 80, p.feature + iB * ldF + 80, r);
                }
                store_c<2>(q, p.post_probs_l + iA * ldP, p.post_probs_l + iB * ldP, r);
                if (sv) {
                    store_rec_dst<2, TILED>(la, sv, mts::LA, r);
                    store_rec_dst<2, TILED>(lv, sv, mts::LV, r);
                }
                // ---- lower prior (:285-286): off the recurrence, in the shadow of the state warp's cells -------------------
                float ppl[2][4];
                {
                    float acc[4][4], lg[2][4];
                    init_bias<4>(acc, bias + mt::B_LP1, r.t);
                    gemm<NS, 2, 4, PAIRED>(acc, dlf, wblk<NS>(W, mt::LP1), lane);
                    head2_l2<TILED, PAIRED>(acc, lg, bias + mt::B_LP2, wblk<NS>(W, mt::LP2), sv, mts::LP_HID, r, lane);
                    softmax_groups<KL, true>(lg, ppl);
                    store_c<2>(ppl, p.prior_probs_l + iA * ldP, p.prior_probs_l + iB * ldP, r);
                }
                if (prior_draws) {  // the prior MTState's own draw (state.py:49)
                    float zl[2][4];
                    sample_onehot_regs<KL>(ppl, upA, upB, zl, lane);
                    if (p.prior_stoch_l != nullptr) store_c<2>(zl, p.prior_stoch_l + iA * ldS, p.prior_stoch_l + iB * ldS, r);
                    if (t + 1 < T) fetch_uniforms<KL>(p.u_prior_l, iA + 1, iB + 1, lane, upA, upB);
                }
                float kl[2];
                kl_rows<true>(q, ppl, kl);
                if (r.t == 0) {
                    if (r.vA) p.kl_l[iA * ldK] = kl[0];
                    if (r.vB) p.kl_l[iB * ldK] = kl[1];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------------
template <int KL, int KH>
static cudaError_t launch_fwd2_k(const MtrssmFwdArgs& a, cudaStream_t s) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // tiles per CTA: as few as still cover the batch with one CTA per SM (B = 256 -> 16 CTAs of one tile, not 2 CTAs of 8)
    const int ntiles = (a.B + 15) / 16;
    int tpc = 1;
    while (tpc < 8 && (ntiles + tpc - 1) / tpc > sms) tpc *= 2;
    const int groups = (ntiles + tpc - 1) / tpc;
    const size_t smem = (size_t)mt::FWD_TILES * 32 * sizeof(uint2) + mt::FWD_BIAS * sizeof(float) + (size_t)tpc * f2::TILE_BYTES;
    // a lone tile per CTA (small batches, long horizons) is a serial dependent chain: unpaired B fragments there (grouped layout only)
    auto kernel = !a.rec_tiled ? mtrssm_fwd2_kernel<KL, KH, 0, true>
                  : a.ld_feature == 0 ? mtrssm_fwd2_kernel<KL, KH, 1, true>
                  : tpc == 1 ? mtrssm_fwd2_kernel<KL, KH, 2, false> : mtrssm_fwd2_kernel<KL, KH, 2, true>;
    if (!a.rec_tiled && a.ld_feature != 0) return cudaErrorInvalidValue;
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)((size_t)mt::FWD_TILES * 32 * sizeof(uint2) + mt::FWD_BIAS * sizeof(float) + 8 * f2::TILE_BYTES));
    if (err != cudaSuccess) return err;
    kernel<<<groups < sms ? groups : sms, 64 * tpc, smem, s>>>(a);
    return cudaGetLastError();
}

// bf16 policy, posterior rollout only (imagination and the fp32-parity policy keep mtrssm_fwd_kernel)
cudaError_t launch_mtrssm_fwd2(const MtrssmFwdArgs& a, cudaStream_t s) {
    if (a.KL == 4 && a.KH == 2) return launch_fwd2_k<4, 2>(a, s);
#ifndef RSSM_EXP_ONLY_DEFAULT
    if (a.KL == 4 && a.KH == 4) return launch_fwd2_k<4, 4>(a, s);
    if (a.KL == 2 && a.KH == 2) return launch_fwd2_k<2, 2>(a, s);
    if (a.KL == 8 && a.KH == 8) return launch_fwd2_k<8, 8>(a, s);
    if (a.KL == 16 && a.KH == 16) return launch_fwd2_k<16, 16>(a, s);
#endif
    return cudaErrorInvalidValue;
}

}  // namespace rssm
