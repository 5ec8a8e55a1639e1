// Shared between the MoPoE-MMTRSSM kernels (mtrssm_kernels.cu, mtrssm_fused_bwd.cu): weight-block / record layouts and the
// backward kernels' staging helpers.
#pragma once
#include "frag.cuh"
#include "kernels.h"

namespace rssm {

namespace mt {
// ---- forward weight blocks (tile offsets) ----------------------------------------------------------
constexpr int L_D2H = 0;              // d_l_prev (32) -> l pre (32)       KT2 NT4
constexpr int L_IN_ZL = L_D2H + 8;    // z_l_prev (16) -> l pre            KT1 NT4
constexpr int L_IN_ZH = L_IN_ZL + 4;  // z_h_prev (16) -> l pre            KT1 NT4
constexpr int L_IN_A = L_IN_ZH + 4;   // action        -> l pre            KT1 NT4
constexpr int H_D2H = L_IN_A + 4;     // d_h_prev      -> h pre            KT2 NT4
constexpr int H_IN = H_D2H + 8;       // z_h_prev      -> h pre            KT1 NT4
constexpr int LP1 = H_IN + 4, LP2 = LP1 + 8;
constexpr int HP1 = LP2 + 4, HP2 = HP1 + 8;
constexpr int HQ1L = HP2 + 4, HQ1H = HQ1L + 8, HQ2 = HQ1H + 8;
constexpr int A1H = HQ2 + 4, A1E = A1H + 8, A2 = A1E + 16;
constexpr int V1H = A2 + 4, V1E = V1H + 8, V2 = V1E + 16;
constexpr int FWD_TILES = V2 + 4;  // 132
constexpr int B_L = 0, B_H = 32, B_LP1 = 64, B_LP2 = 96, B_HP1 = 112, B_HP2 = 144, B_HQ1 = 160, B_HQ2 = 192, B_A1 = 208,
              B_A2 = 240, B_V1 = 256, B_V2 = 288, FWD_BIAS = 304;
// ---- backward (transposed) weight blocks -----------------------------------------------------------
constexpr int T_A2 = 0, T_V2 = 4, T_LP2 = 8, T_HP2 = 12, T_HQ2 = 16;                       // KT1 NT4
constexpr int T_A1H = 20, T_V1H = 28, T_LP1 = 36, T_HP1 = 44, T_HQ1L = 52, T_HQ1H = 60;   // KT2 NT4
constexpr int T_A1E = 68, T_V1E = 84;                                                       // KT2 NT8
constexpr int T_L_D2H = 100, T_H_D2H = 108;                                                 // KT2 NT4
constexpr int T_L_IN_ZL = 116, T_L_IN_ZH = 120, T_H_IN = 124;                               // KT2 NT2
constexpr int T_L_IN_A = 128;                                                               // KT2 NT2 (one 16-block)
constexpr int BWD_TILES = 132;
}  // namespace mt

namespace mts {  // saved record (192 used of MTRSSM_SAVED_FLOATS)
constexpr int LP_HID = 0, HP_HID = 32, HQ_HID = 64, A_HID = 96, V_HID = 128, LA = 160, LV = 176;
}

namespace mtd {  // dpre record (304 used of MTRSSM_DPRE_FLOATS)
constexpr int L = 0, H = 32, LP1 = 64, LPL = 96, HP1 = 112, HPL = 144, HQ1 = 160, HQL = 192, A1 = 208, LA = 240, V1 = 256, LV = 288;
}

template <int NS>
__device__ __forceinline__ uint2* wblk(uint2* W, int tile_off) {
    return W + (size_t)NS * tile_off * 32;
}

// ---- per-warp staging of the backward kernel's per-step inputs (bf16 path) ---------------------------------------
// Single-buffered "consume early, refill at once": each buffer is read by step t and, right after its last read,
// re-filled for step t-1, which then has most of a step to land.
//   * the three wide rows -- d_feature (384 B), the saved record (384 B), feature[0:80] (320 B: deter_h, stoch_h,
//     deter_l) -- travel as BULK copies (cp.async.bulk, one per row, completion on a per-warp mbarrier): the bulk path
//     sustains ~6.5 TB/s with 4-8 warps per SM where 16-byte cp.async saturates the LSU at ~3.7 TB/s
//     (profiles/r1_umma_probe.txt).  Rows land linearly; the row pitch is padded so that the fragment-pattern reads
//     are bank-conflict free (pitch = 64 mod 128 bytes for 16-byte reads, 32 mod 128 for the 8-byte bf16 reads);
//   * the four 64-byte probability rows stay on cp.async (swizzled), bulk copies being op-rate bound for small rows.
// Layout (32-bit words per warp):  DF [16][112] | SV [16][104] | FT [16][80] | PR [16][64]
namespace bst {
constexpr int DF_LD = 112, SV_LD = 104, FT_LD = 80;
constexpr int DF = 0, SV = DF + 16 * DF_LD, FT = SV + 16 * SV_LD, PR = FT + 16 * FT_LD, WORDS = PR + 16 * 64;  // 23040 bytes per warp
constexpr int DF_BYTES = 384, SV_BYTES = 384, FT_BYTES = 320;
enum { BAR_DF, BAR_SV, BAR_FT, NBAR };
}  // namespace bst

__device__ __forceinline__ int sw32(int chunk, int row) { return chunk ^ (4 * (row & 1)); }   // fp32 rows, float4 reads

// Issue one bulk copy per row (lanes 0..15) of `bytes` from src + idx(row) * ld_bytes into dst + row * pitch_words.
// Every lane has finished reading the buffer (the caller's __syncwarp); the proxy fence orders those generic-proxy
// reads before the async-proxy writes.
// (One lane issuing all 16 copies from warp-uniform addresses needs 100 instead of 170 instructions per call -- the bulk-copy
// instruction takes uniform operands, so the 16 lanes below are serialised by a vote loop -- but is 5 % SLOWER at the bench batch:
// measured, profiles/r2_ab_variants.txt.)
__device__ __forceinline__ void bulk_rows(float* dst, int pitch_words, const char* src, size_t ld_bytes, uint32_t bytes, int row0, int B,
                                          int T, int t, uint64_t* bar, int lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane == 0) mbar_expect_tx(bar, 16 * bytes);
    __syncwarp();
    if (lane < 16) {
        const size_t idx = (size_t)min(row0 + lane, B - 1) * T + t;
        bulk_g2s(dst + lane * pitch_words, src + idx * ld_bytes, bytes, bar);
    }
}

// the four probability tensors of step t (one cp.async group): lane -> chunk column c8 = lane & 7 of the rows rq + 4j
// LDP: row pitch (floats) of the probability tensors -- 16 (dense) or MTRSSM_ROW_PITCH (grouped rows)
template <int LDP = 16>
__device__ __forceinline__ void bstage_pr(float* st, const MtrssmBwdArgs& p, int row0, int t, int lane) {
    if (t >= 0) {
        const int c8 = lane & 7, rq = lane >> 3;
        // 16 rows x (4 tensors x 4 chunks): chunk 8*cgrp + c8  ->  tensor 2*cgrp + (c8 >> 2), its chunk c8 & 3
        float* d2 = st + bst::PR + rq * 64 + 4 * (c8 ^ (4 * (rq & 1)));
        const bool hi = (c8 >> 2) != 0;
        const float* src0 = hi ? p.post_probs_l : p.post_probs_h;
        const float* src1 = hi ? p.prior_probs_l : p.prior_probs_h;
        constexpr int ldP = LDP;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const size_t idx = (size_t)min(row0 + rq + 4 * j, p.B - 1) * p.T + t;
            cp_async16(d2 + j * 256, src0 + idx * ldP + 4 * (c8 & 3));
            cp_async16(d2 + j * 256 + 32, src1 + idx * ldP + 4 * (c8 & 3));
        }
    }
    cp_async_commit();
}

// NT tiles starting at logical column col0 of a staged fp32 buffer with `stride` words per row;
// SWZ: 16-byte-chunk XOR swizzle (cp.async-staged buffers), else linear padded rows (bulk-staged buffers)
template <int NT, bool SWZ>
__device__ __forceinline__ void load_staged(float (&c)[NT][4], const float* buf, int stride, int col0, int g, int t) {
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) {
        const int ch = SWZ ? sw32(col0 / 4 + 4 * j + t, g) : col0 / 4 + 4 * j + t;
        const float4 a = *reinterpret_cast<const float4*>(buf + g * stride + 4 * ch);
        const float4 b = *reinterpret_cast<const float4*>(buf + (g + 8) * stride + 4 * ch);
        c[2 * j][0] = a.x, c[2 * j][1] = a.y, c[2 * j + 1][0] = a.z, c[2 * j + 1][1] = a.w;
        c[2 * j][2] = b.x, c[2 * j][3] = b.y, c[2 * j + 1][2] = b.z, c[2 * j + 1][3] = b.w;
    }
}

// NT tiles starting at record element `off` of the staged bf16 saved rows (linear, pitch bst::SV_LD words)
template <int NT>
__device__ __forceinline__ void load_staged_rec(float (&c)[NT][4], const float* buf, int off, int g, int t) {
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) {
        const int w = (off + 16 * j + 4 * t) >> 1;  // 32-bit word of the row
        const uint2 a = *reinterpret_cast<const uint2*>(buf + g * bst::SV_LD + w), b = *reinterpret_cast<const uint2*>(buf + (g + 8) * bst::SV_LD + w);
        const __nv_bfloat162 a0 = *reinterpret_cast<const __nv_bfloat162*>(&a.x), a1 = *reinterpret_cast<const __nv_bfloat162*>(&a.y);
        const __nv_bfloat162 b0 = *reinterpret_cast<const __nv_bfloat162*>(&b.x), b1 = *reinterpret_cast<const __nv_bfloat162*>(&b.y);
        c[2 * j][0] = __low2float(a0), c[2 * j][1] = __high2float(a0), c[2 * j + 1][0] = __low2float(a1), c[2 * j + 1][1] = __high2float(a1);
        c[2 * j][2] = __low2float(b0), c[2 * j][3] = __high2float(b0), c[2 * j + 1][2] = __low2float(b1), c[2 * j + 1][3] = __high2float(b1);
    }
}

// d logits (16) -> through W2^T -> * ELU'(hidden) -> dpre1 (stored) ; returns dpre1 as A operand.
// `hid` = the head's saved post-ELU hidden.
template <int NS>
__device__ __forceinline__ void head_bwd(const float (&dlogit)[2][4], const uint2* w2t, const float (&hid)[4][4], typename Rec<NS>::T* dpA,
                                         typename Rec<NS>::T* dpB, int dp_logit_off, int dp1_off, AFrag<NS, 2>& f1, const Rows& r,
                                         int lane) {
    store_rec<2>(dlogit, dpA + dp_logit_off, dpB + dp_logit_off, r);
    AFrag<NS, 1> fl;
    to_afrag<NS, 1>(fl, dlogit);
    float dhid[4][4];
    zero_c<4>(dhid);
    gemm<NS, 1, 4>(dhid, fl, w2t, lane);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) dhid[nt][j] *= elu_grad_from_out(hid[nt][j]);
    store_rec<4>(dhid, dpA + dp1_off, dpB + dp1_off, r);
    to_afrag<NS, 2>(f1, dhid);
}

template <int NT>
__device__ __forceinline__ void add_global(float (&acc)[NT][4], const float* base, size_t offA, size_t offB, int t) {
    if (base == nullptr) return;
    float g[NT][4];
    load_c<NT>(g, base + offA, base + offB, t);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[nt][j] += g[nt][j];
}

}  // namespace rssm
