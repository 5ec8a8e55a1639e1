// MoPoE-MMTRSSM: BPTT backward FUSED with the weight-gradient contractions (bf16 tensor-core path).
//
// Same recurrence and per-warp register scheme as mtrssm_bwd_kernel (mtrssm_kernels.cu): one warp owns 16 sequences and
// walks t = T-1 .. 0 carrying only data gradients.  What changes is where the per-(b,t) layer pre-activation gradients
// ("dY", 304 columns) go: instead of a [B,T,304] record in HBM that a second kernel re-reads together with every layer
// input (the `dpre` round trip: 608 B written + ~1.9 KB read per (b,t)), each warp drops its 16 x 304 bf16 dY tile into
// shared memory in the tcgen05 operand layout, next to bf16 copies of the layer inputs that the forward kernel left in
// the saved record, and one elected lane issues
//     dW[128 x N] (TMEM, fp32) += X^T[128 x 16 rows] . dY[16 rows x N]         tcgen05.mma, cta_group::1, kind::f16
// for every layer (8 + 3 MMAs per warp-step, K = the warp's 16 rows).  All warps of the CTA accumulate into the same
// TMEM tiles (448 of the 512 columns); at the end the CTA reads them back (tcgen05.ld) and adds them to global memory
// with one atomicAdd per weight element per CTA.  The tensor work is ~1 % of the tcgen05 pipe; the point is that the
// accumulators cost no registers and the operands no HBM round trip.
//
// Operand layout (validated by profiles/src/umma_probe.cu, profiles/r1_umma_probe.txt): MN-major, no swizzle,
//     element (column c, row k) of a 16-row tile  ->  (c / 8) * 256 + (k / 8) * 128 + (k % 8) * 16 + (c % 8) * 2   bytes,
// i.e. "chunks" of 8 columns (256 B); descriptor LBO = 128 (k-group stride), SBO = 256 (chunk stride).  An MMA's M (or
// N) window is any run of consecutive chunks, so layers address sub-ranges of one staged row image directly.
//
// Reference semantics: autograd of MoPoE_MMTRSSM.rollout_representation (mmtrssm/mopoe_mmtrssm/core.py:364-494).
#include <stdlib.h>

#include "frag.cuh"
#include "kernels.h"
#include "mtrssm_common.cuh"

// timing experiments only (wrong weight gradients): -DFZ_NO_FENCE drops the proxy fences before the MMAs, -DFZ_NO_MMA the MMAs
#ifdef FZ_NO_FENCE
#define FZ_FENCE()
#else
#define FZ_FENCE() asm volatile("fence.proxy.async.shared::cta;" ::: "memory")
#endif
#ifdef FZ_NO_MMA
#define FZ_MMA false
#define FZ_WAIT(bar, ph) (void)0
#else
#define FZ_MMA true
#define FZ_WAIT(bar, ph) mbar_wait(bar, ph)
#endif

// -DFZ_TIMING (profiles/src/build_variant.sh): tile 0 of CTA 0 stamps clock64() at fixed points of every step of its first group;
// the launcher synchronises and prints the mean interval between consecutive stamps (cycles) for both warps of the tile.
#ifdef FZ_TIMING
#include <stdio.h>
__device__ long long fz_dbg[3][1024][12];
__device__ long long fz_grp[64];  // CTA 0, thread 0: kernel entry, prologue done, per group {start, end}, epilogue start / end
__device__ long long fz_cta[256][3];  // per CTA: globaltimer at entry / exit (ns), SM id
#define FZ_GS(i)                                                   \
    do {                                                           \
        if (threadIdx.x == 0 && blockIdx.x == 0) fz_grp[i] = clock64(); \
        if (threadIdx.x == 0 && ((i) == 0 || (i) == 41) && blockIdx.x < 256) {                      \
            long long gt_;                                                                          \
            unsigned sm_;                                                                           \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                                 \
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_));                                        \
            fz_cta[blockIdx.x][(i) == 0 ? 0 : 1] = gt_, fz_cta[blockIdx.x][2] = sm_;                \
        }                                                                                           \
    } while (0)
#define FZ_TS(i)                                                                                                  \
    do {                                                                                                          \
        if (lane == 0 && blockIdx.x == 0 && tile == 0 && grp == (int)blockIdx.x && t < 1024) fz_dbg[role][t][i] = clock64(); \
    } while (0)
#else
#define FZ_TS(i) do {} while (0)
#define FZ_GS(i) do {} while (0)
#endif

namespace rssm {

namespace fz {
constexpr int CH = 256;  // bytes of one operand chunk (8 columns x 16 rows, bf16)
// ---- dY column order (elements); every MMA's N window is one contiguous run ----------------------------------------
constexpr int Y_LPL = 0, Y_HPL = 16, Y_HQL = 32, Y_LA = 48, Y_LV = 64;            // second-layer (logit) gradients
constexpr int Y_HQ1 = 80, Y_LP1 = 112, Y_A1 = 144, Y_V1 = 176, Y_HP1 = 208;       // first-layer pre-activation gradients
constexpr int Y_L = 240, Y_H = 272;                                                // the two MTRNN cells
}  // namespace fz

// one (lane range, column range) block of a TMEM accumulator tile -> global weight gradient
struct FusedFlush {
    float* dst;      // element (lane L, column c) is added to dst[(c - col0) * ld + (L - lane0) + koff]   (bias rows: ld = 0)
    int tcol;        // first TMEM column of the block
    int ncols;       // columns of the block
    int lane0, nlanes;
    int ld, koff;
};
constexpr int MAX_FUSED_FLUSH = 40;
struct FusedFlushTable {
    int n;
    FusedFlush e[MAX_FUSED_FLUSH];
};

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {  // MN-major, no swizzle: LBO = 128 B, SBO = 256 B, version 1
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t umma_idesc(int N) {  // bf16 x bf16 -> fp32, A and B MN-major, M = 128
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// The descriptors' high words are one constant (SBO, version) and the low word is (address >> 4) | LBO << 16: the callers keep the
// low words of their images (desc_lo, once per group) and add compile-time chunk offsets -- ~5 instructions per MMA in the
// single-lane issue block instead of ~20 (shift / mask / or / 64-bit assembly of both descriptors).
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3fffu) | ((uint32_t)(128 >> 4) << 16); }
constexpr uint32_t DESC_HI = (uint32_t)(256 >> 4) | (1u << 14);
__host__ __device__ constexpr uint32_t desc_off(int bytes) { return (uint32_t)bytes >> 4; }  // shared memory is < 256 KB: the 14-bit field cannot carry
__device__ __forceinline__ void umma_acc(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, int N) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(umma_idesc(N)), "r"(1u), "r"(DESC_HI)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// C tiles (NT/2 blocks of 16 columns, starting at operand column col0) -> bf16 operand chunks.  Rows of sequences beyond the batch
// are never written: a tile's images are zeroed once per group (and a row's validity does not change within a group), so they
// stay zero and contribute nothing to the weight gradients -- predicated stores instead of two selects per store.
template <int NT>
__device__ __forceinline__ void store_op(const float (&c)[NT][4], unsigned char* op, int col0, const Rows& r) {
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) {
        unsigned char* q = op + ((col0 + 16 * j) / 8 + (r.t >> 1)) * fz::CH + r.g * 16 + (r.t & 1) * 8;
        const uint2 a = make_uint2(pack_bf16(c[2 * j][0], c[2 * j][1]), pack_bf16(c[2 * j + 1][0], c[2 * j + 1][1]));
        const uint2 b = make_uint2(pack_bf16(c[2 * j][2], c[2 * j][3]), pack_bf16(c[2 * j + 1][2], c[2 * j + 1][3]));
        if (r.vA) *reinterpret_cast<uint2*>(q) = a;        // row g     (k-group 0)
        if (r.vB) *reinterpret_cast<uint2*>(q + 128) = b;  // row g + 8 (k-group 1)
    }
}
template <int NT>
__device__ __forceinline__ void load_op(float (&c)[NT][4], const unsigned char* op, int col0, int g, int t) {
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) {
        const unsigned char* q = op + ((col0 + 16 * j) / 8 + (t >> 1)) * fz::CH + g * 16 + (t & 1) * 8;
        const uint2 a = *reinterpret_cast<const uint2*>(q), b = *reinterpret_cast<const uint2*>(q + 128);
        const __nv_bfloat162 a0 = *reinterpret_cast<const __nv_bfloat162*>(&a.x), a1 = *reinterpret_cast<const __nv_bfloat162*>(&a.y);
        const __nv_bfloat162 b0 = *reinterpret_cast<const __nv_bfloat162*>(&b.x), b1 = *reinterpret_cast<const __nv_bfloat162*>(&b.y);
        c[2 * j][0] = __low2float(a0), c[2 * j][1] = __high2float(a0), c[2 * j + 1][0] = __low2float(a1), c[2 * j + 1][1] = __high2float(a1);
        c[2 * j][2] = __low2float(b0), c[2 * j][3] = __high2float(b0), c[2 * j + 1][2] = __low2float(b1), c[2 * j + 1][3] = __high2float(b1);
    }
}

// (Round 1 / early round 2 staged a ROW-layout record [B][T][208] here: a gather of 16-byte pieces at a T x 416-byte stride, one
// shared-memory wavefront per lane.  The fused policy's record is tile-blocked now; the row layout lives on in the two-kernel path.)
// Tile-blocked record (MtrssmBwdArgs.rec_tiled): chunks c0 .. c0+n-1 of a tile-step are n x 256 CONTIGUOUS bytes in global memory
// AND in the operand image: the warp copies them as a linear run of 16-byte pieces (piece = lane + 32 i), whole 128-byte lines on
// both sides -- 4 shared-memory wavefronts per instruction where a row-layout gather needs 32 (one per lane).
// (A cp.async.bulk per run was measured as well: no faster at the bench size, slower for one tile -- fence + single-lane issue.)
// `tile_rec` = the tile's record of step 0, this lane's piece (saved + tile * T * 6656 B + lane * 16: computed once per group)
__device__ __forceinline__ void stage_chunks_tiled(unsigned char* dst0, const unsigned char* tile_rec, int c0, int n, int t, int lane) {
    const unsigned char* src = tile_rec + (size_t)(uint32_t)t * (MTRSSM_SAVED_BF16 * 16 * 2) + c0 * 256;
    unsigned char* dst = dst0 + lane * 16;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        if (lane + 32 * i < n * 16) cp_async16(dst + 512 * i, src + 512 * i);
    }
}

// d logits (16) -> through W2^T -> * ELU'(hidden) -> dY1; both land in the dY operand image; returns dY1 as A operand
__device__ __forceinline__ void head_bwd_op(const float (&dlogit)[2][4], const uint2* w2t, const float (&hid)[4][4], unsigned char* dy,
                                            int y_logit, int y1, AFrag<1, 2>& f1, const Rows& r, int lane) {
    store_op<2>(dlogit, dy, y_logit, r);
    AFrag<1, 1> fl;
    to_afrag<1, 1>(fl, dlogit);
    float dhid[4][4];
    zero_c<4>(dhid);
    gemm<1, 1, 4>(dhid, fl, w2t, lane);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) dhid[nt][j] *= elu_grad_from_out(hid[nt][j]);
    store_op<4>(dhid, dy, y1, r);
    to_afrag<1, 2>(f1, dhid);
}

// the same with ELU'(hidden) precomputed by the caller (mod warp: off the recurrence's critical section)
__device__ __forceinline__ void head_bwd_elu(const float (&dlogit)[2][4], const uint2* w2t, const float (&elu_grad)[4][4], unsigned char* dy,
                                             int y_logit, int y1, AFrag<1, 2>& f1, const Rows& r, int lane, float* dpA = nullptr,
                                             float* dpB = nullptr) {
    store_op<2>(dlogit, dy, y_logit, r);
    AFrag<1, 1> fl;
    to_afrag<1, 1>(fl, dlogit);
    float dhid[4][4];
    zero_c<4>(dhid);
    gemm<1, 1, 4>(dhid, fl, w2t, lane);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) dhid[nt][j] *= elu_grad[nt][j];
    store_op<4>(dhid, dy, y1, r);
    to_afrag<1, 2>(f1, dhid);
    if (dpA != nullptr) store_c<4>(dhid, dpA, dpB, r);  // obs_projected: dY1 IS the gradient of the pre-multiplied partial (fp32)
}

// [action | ones] X-operand columns of one step: lanes with t < 2 own action columns 4t .. 4t+3, column 8 is the ones column
__device__ __forceinline__ void load_action_columns(float (&actc)[2][4], const float* actions, size_t iA, size_t iB, int A, bool valid,
                                                    const Rows& r) {
    if (valid && r.t < 2) {
        const float* aA = actions + iA * A + 4 * r.t;
        const float* aB = actions + iB * A + 4 * r.t;
        if (4 * r.t < A) {
            const float2 x = *reinterpret_cast<const float2*>(aA), y = *reinterpret_cast<const float2*>(aB);
            actc[0][0] = x.x, actc[0][1] = x.y, actc[0][2] = y.x, actc[0][3] = y.y;
        }
        if (4 * r.t + 2 < A) {
            const float2 x = *reinterpret_cast<const float2*>(aA + 2), y = *reinterpret_cast<const float2*>(aB + 2);
            actc[1][0] = x.x, actc[1][1] = x.y, actc[1][2] = y.x, actc[1][3] = y.y;
        }
    }
    if (r.t == 2) actc[0][0] = actc[0][2] = 1.f;
}

// =====================================================================================================================
// Version 2: TWO warps per 16-sequence tile.
// The single-warp kernel above needs ~46 KB of shared memory per tile, i.e. four tiles = four warps per SM = one warp per
// scheduler, and every dependent-instruction latency is exposed.  A backward step has two branches that are independent
// given the carried gradients and join only at the two cells:
//     "core" warp: higher posterior + prior heads, the two leaky-integrator cells (owns ddl ddh dul duh dzh)
//     "mod"  warp: MoPoE-fusion backward, audio and vision heads, lower prior head (receives dzl, returns its contribution to ddl)
// They share the tile's staged inputs and operand images and exchange 3 KB per step through shared memory, ordered by two
// named barriers (X: mod -> core "my heads are done", Y: core -> mod "dzl of the next step is ready").  Same footprint per
// tile, twice the warps per SM, and a per-step critical path of (MoPoE + two modality heads) + cells.
// MMAs are issued by the warp that owns the operands: mod issues {[a|v hid] x [LA|LV], [lp hid] x [LPL], [embed_a|embed_v] x [A1|V1]},
// core issues {[hp|hq hid] x [HPL|HQL]} after its heads and {deter x all first layers, cells, biases} at the cells.
// =====================================================================================================================
namespace fz2 {
constexpr int CH = fz::CH;
// ---- per-TILE shared-memory map (bytes) ---------------------------------------------------------------------------------
constexpr int DOP = 0;                      // [d_l 4][d_h 4] chunks: bf16 deter of this step (core, from the staged feature row)
constexpr int ZOP = DOP + 8 * CH;           // [z_l 2][z_h 2][act 1][ones 1] chunks: this step's stoch, the NEXT step's action, ones column
constexpr int SVOP = ZOP + 6 * CH;          // record chunks 0..23 (core: lp, hp, hq hid; mod: a, v hid, LA, LV) + embedding images 24..39 (mod)
constexpr int DYOP = SVOP + 40 * CH;        // dY: 30 chunks (columns 0..239, fz::Y_* order) + 2 x 8 chunks [L | H], double buffered
constexpr int DF = DYOP + 46 * CH;          // d_feature [16][112 words] (bulk, core)
constexpr int FT = DF + 16 * bst::DF_LD * 4;   // feature row [16][112 words] (bulk, core)
constexpr int PR = FT + 16 * bst::DF_LD * 4;   // 4 probability tensors [16][64 words] (cp.async, core)
constexpr int PRL = PR + 16 * 64 * 4;          // post_l | prior_l [16][32 words] (cp.async, mod's own copy)
constexpr int XDDL = PRL + 16 * 32 * 4;        // mod -> core: contribution to d deter_l   [4][32 lanes][4] fp32
constexpr int XDZL = XDDL + 2048;              // core -> mod: d stoch_l of the next step  [2][32 lanes][4] fp32
constexpr int BYTES = XDZL + 1024;             // 49,152
// ---- TMEM accumulator columns ---------------------------------------------------------------------------------------
constexpr int T_E = 0;      // [lp|hp|hq|a hid]  x [LPL|HPL|HQL]                    48 columns
constexpr int T_M = 48;     // [a|v hid|..]      x [LA|LV]                          32
constexpr int T_EMB = 80;   // [embed_a|embed_v] x [A1|V1]                          64
constexpr int T_D1 = 144;   // [d_l|d_h|..]      x [HQ1|LP1|A1|V1|HP1]             160
constexpr int T_C = 304;    // [d_l|d_h|z_l|z_h|act|ones|..] of step t x [L|H] of step t+1   64 (the ones row carries the cells' biases)
constexpr int T_B = 368;    // 2 x (dY window x [ones|..])                      2 x 16
constexpr int T_COLS = 400;
enum { BAR_DF, BAR_FT, BAR_E, BAR_END, BAR_M, NBAR };  // per tile
}  // namespace fz2

__device__ __forceinline__ void nbar_sync(int id, int count = 64) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int count = 64) {
    __threadfence_block();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// C tiles <-> the lane-linear fp32 exchange buffers ([tile][32 lanes][4 floats]: conflict-free 16-byte accesses)
template <int NT>
__device__ __forceinline__ void xch_store(const float (&c)[NT][4], float* buf, int lane) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<float4*>(buf + (nt * 32 + lane) * 4) = make_float4(c[nt][0], c[nt][1], c[nt][2], c[nt][3]);
}
template <int NT>
__device__ __forceinline__ void xch_load(float (&c)[NT][4], const float* buf, int lane) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const float4 v = *reinterpret_cast<const float4*>(buf + (nt * 32 + lane) * 4);
        c[nt][0] = v.x, c[nt][1] = v.y, c[nt][2] = v.z, c[nt][3] = v.w;
    }
}

// WPT = warps per tile.  2: core + mod (above).  3: core + mod + AUX -- clock64 timelines (profiles/r2_b_timing.txt) show both
// warps of the two-warp kernel busy ~8.7 k cycles per step with NO waiting on each other: each is a serial instruction stream at
// ~0.1 instructions per cycle, and a third of it is "service" work with no place in the recurrence.  The aux warp takes that
// work: it converts the feature row and the fp32 embeddings into the tcgen05 operand images, issues the end-of-step MMA group
// (and the embedding MMA), computes the embedding / action gradients from the dY images (d_embed = dY1 . W1e, d_action = dY_l .
// W_in) and stores them, and refills the feature-row stage.  12 warps per SM instead of 8, <= 168 registers each.
// GROUPED: the forward's outputs share one MTRSSM_ROW_PITCH-float row per (b,t) (RssmMtrssmOutputs.ld_*); compile-time pitches.
// PROJ: pre-multiplied observation partials (dims.obs_projected; three-warp kernel only).  The saved record is always tile-blocked.
// (Layout / mode switches are template parameters, not runtime flags: the dead branches of an 8 k-instruction kernel cost 2.4 % at
// the bench batch -- 15 % of this kernel's stall samples are instruction fetch.)
template <int KL, int KH, int WPT, bool GROUPED, bool PROJ>
__global__ void __launch_bounds__(128 * WPT, 1) mtrssm_bwd_fused2_kernel(const MtrssmBwdArgs p, const FusedFlushTable ft) {
    constexpr int NS = 1;
    static_assert(WPT == 2 || WPT == 3, "two (core, mod) or three (core, mod, aux) warps per tile");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars_all[4][fz2::NBAR];
    __shared__ uint32_t tmem_base_s;
    uint2* W = reinterpret_cast<uint2*>(smem_raw);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int A = p.A;
    FZ_GS(0);
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
    // the warp index as a BROADCAST value: the compiler then knows tile / role / row0 (and every address derived from them) are
    // warp-uniform -- uniform registers and uniform branches instead of per-lane integer arithmetic
    const int lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), tile = warp / WPT, role = warp % WPT;  // role 0 = core, 1 = mod, 2 = aux
    // one tile per CTA (batches of at most one tile per SM, e.g. cfg4's 16 tiles): 128 threads = the tile's three warps + one
    // HELPER warp that only takes part in the CTA-wide steps (weight packing, TMEM zeroing and read-back need four warps)
    const bool active = tile < nthr / (32 * WPT);
    unsigned char* my = smem_raw + (size_t)mt::BWD_TILES * 32 * sizeof(uint2) + (size_t)(active ? tile : 0) * fz2::BYTES;
    // zero the tile's images and exchange buffers once (the warps of the tile split the range)
    if (active)
        for (int i = lane + 32 * role; i < fz2::BYTES / 16; i += 32 * WPT) reinterpret_cast<uint4*>(my)[i] = make_uint4(0u, 0u, 0u, 0u);
    uint64_t* bars = bars_all[active ? tile : 0];
    if (active && role == 0 && lane == 0) {
#pragma unroll
        for (int i = 0; i < fz2::NBAR; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (warp < 4) {
        const uint32_t z = 0u;
        for (int c = 0; c < fz2::T_COLS; c += 16)
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(
                    tmem + ((uint32_t)(32 * warp) << 16) + c),
                "r"(z)
                : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // Persistent over tile groups: a CTA (2 or 4 tiles = 4 or 8 warps) walks groups blockIdx.x, blockIdx.x + gridDim.x, ... and keeps
    // accumulating the weight gradients in the SAME TMEM columns, so the weight packing above, the TMEM allocation and the
    // read-back + atomics below happen once per SM instead of once per group (4 groups per SM at the bench size).
    const int tpc = nthr / (32 * WPT), ngroups = ((p.B + 15) / 16 + tpc - 1) / tpc;
    FZ_GS(1);
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    FZ_GS(2 + 2 * (grp / (int)gridDim.x));
    const int row0 = (grp * tpc + tile) * 16;
    if (active && row0 < p.B) {
        const Rows r = make_rows(row0, p.B, lane);
        const int T = p.T;
        // this lane's 16-byte piece of the tile's saved record of step 0 (tile-blocked: 6656 contiguous bytes per tile-step)
        const unsigned char* tile_rec =
            reinterpret_cast<const unsigned char*>(p.saved) + (size_t)(row0 >> 4) * T * (MTRSSM_SAVED_BF16 * 16 * 2) + lane * 16;
        unsigned char* dop = my + fz2::DOP;
        unsigned char* svop = my + fz2::SVOP;
        unsigned char* zop = my + fz2::ZOP;
        unsigned char* dy = my + fz2::DYOP;
        float* xddl = reinterpret_cast<float*>(my + fz2::XDDL);
        float* xdzl = reinterpret_cast<float*>(my + fz2::XDZL);
        const uint32_t s_dop = desc_lo(smem_u32(dop)), s_sv = desc_lo(smem_u32(svop)), s_dy = desc_lo(smem_u32(dy)),
                       s_ones = desc_lo(smem_u32(zop) + 5 * fz2::CH);  // descriptor low words of the tile's images
        // named barriers of the tile.  X: mod -> core (+ aux) "my dY columns and XDDL are complete"; Y: core -> mod "d stoch_l is in
        // XDZL"; Z (three warps): core -> aux "my dY columns are complete and I am done with the feature-row stage"
        const int bar_x = 1 + WPT * tile, bar_y = 2 + WPT * tile, bar_z = 3 + WPT * tile;
        // row pitches of the grouped forward outputs (include/rssm_rollout.h: feature | hidden | probabilities | prior draws share one
        // 1 KB row per (b,t)); natural widths when the caller keeps them in separate tensors
        constexpr size_t ft_pitch = (size_t)(GROUPED ? MTRSSM_ROW_PITCH : 96) * sizeof(float);
        constexpr int ldP = GROUPED ? MTRSSM_ROW_PITCH : 16;
        constexpr int X_COUNT = 32 * WPT;
        (void)bar_z;

        // step 0's cell gradients pair with the INITIAL state (run by the warp that issues the end-of-step MMAs, after the loop)
        auto pair_initial_state = [&](uint32_t& ph_end) {
            FZ_WAIT(&bars[fz2::BAR_END], ph_end), ph_end ^= 1;
            float c4[4][4], c2[2][4], actc[2][4];
            load_c<4>(c4, p.deter_l0 + (size_t)r.rA * 32, p.deter_l0 + (size_t)r.rB * 32, r.t);
            store_op<4>(c4, dop, 0, r);
            load_c<4>(c4, p.deter_h0 + (size_t)r.rA * 32, p.deter_h0 + (size_t)r.rB * 32, r.t);
            store_op<4>(c4, dop, 32, r);
            load_c<2>(c2, p.stoch_l0 + (size_t)r.rA * 16, p.stoch_l0 + (size_t)r.rB * 16, r.t);
            store_op<2>(c2, zop, 0, r);
            load_c<2>(c2, p.stoch_h0 + (size_t)r.rA * 16, p.stoch_h0 + (size_t)r.rB * 16, r.t);
            store_op<2>(c2, zop, 16, r);
            zero_c<2>(actc);
            load_action_columns(actc, p.actions, (size_t)r.rA * T, (size_t)r.rB * T, A, true, r);
            store_op<2>(actc, zop, 32, r);
            FZ_FENCE();
            __syncwarp();
            if (FZ_MMA && lane == 0) {
                umma_acc(tmem + fz2::T_C, s_dop, s_dy + desc_off((fz::Y_L / 8) * fz2::CH), 64);  // buffer 0 holds step 0's [L | H]
                umma_commit(&bars[fz2::BAR_END]);
            }
            FZ_WAIT(&bars[fz2::BAR_END], ph_end);
        };

        if (role == 1) {
            // ============================== mod warp: MoPoE backward + audio / vision heads ===============================
            float* stPRL = reinterpret_cast<float*>(my + fz2::PRL);  // [16][32]: post_l | prior_l, swizzled like PR
            // (row index x T and the lane's source / destination are fixed for the group: per step one add and one wide multiply-add
            // per copy instead of a clamp and two 64-bit multiplies)
            const int prl_c8 = lane & 7, prl_rq = lane >> 3;
            float* prl_dst = stPRL + prl_rq * 32 + 4 * (prl_c8 ^ (4 * (prl_rq & 1)));
            const float* prl_src = ((prl_c8 >> 2) ? p.prior_probs_l : p.post_probs_l) + 4 * (prl_c8 & 3);
            uint32_t prl_row[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) prl_row[j] = (uint32_t)min(row0 + prl_rq + 4 * j, p.B - 1) * (uint32_t)T;
            auto stage_prl = [&](int t) {
                if (t >= 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) cp_async16(prl_dst + j * 128, prl_src + (size_t)(prl_row[j] + (uint32_t)t) * ldP);
                }
            };
                        auto stage_logits = [&](int t) {  // LA, LV: read by registers only, refilled right after the MoPoE math
                if (t >= 0) {
                    stage_chunks_tiled(svop + 20 * fz2::CH, tile_rec, 20, 4, t, lane);
                }
                stage_prl(t);
                cp_async_commit();
            };
            auto stage_rest = [&](int t) {  // a / v / lp hiddens: free once this warp's MMAs have completed
                if (t >= 0) {
                    stage_chunks_tiled(svop + 12 * fz2::CH, tile_rec, 12, 8, t, lane);
                    stage_chunks_tiled(svop, tile_rec, 0, 4, t, lane);  // lp hidden
                }
                cp_async_commit();
            };
            // the embedding operand images (chunks 24..39) are converted from the fp32 inputs by this warp itself, in the window
            // where it would otherwise wait for the core warp's cells (the rows were pulled into L2 one step earlier)
            auto embed_images = [&](int t) {
                const size_t jA = ((size_t)r.rA * T + t) * 64, jB = ((size_t)r.rB * T + t) * 64;
                float e[8][4];
                load_c<8>(e, p.embed_a + jA, p.embed_a + jB, r.t);
                store_op<8>(e, svop, 24 * 8, r);
                load_c<8>(e, p.embed_v + jA, p.embed_v + jB, r.t);
                store_op<8>(e, svop, 32 * 8, r);
            };
            // lanes 0..15: one 256-byte row of each embedding into L2.  (Per-lane `prefetch.global.L2` of the same lines is 1.7 % SLOWER
            // at the bench batch although it saves ~260 issue slots per step: profiles/r2_ab_variants.txt)
            auto embed_prefetch = [&](int t) {
                if (t >= 0 && lane < 16) {
                    const size_t j = ((size_t)min(row0 + lane, p.B - 1) * T + t) * 64;
                    prefetch_bulk_l2(p.embed_a + j, 256);
                    prefetch_bulk_l2(p.embed_v + j, 256);
                }
            };
            if constexpr (WPT == 2) {
                embed_prefetch(T - 2);
                embed_images(T - 1);
            }
            stage_logits(T - 1);
            stage_rest(T - 1);
            const float* dkl_src = p.d_kl_l;  // lanes t = 0 / 1 of a quad fetch rows A / B one step ahead
            const size_t dkl_row = (size_t)((r.t & 1) ? r.rB : r.rA) * T;
            float dkl_next = dkl_src != nullptr ? dkl_src[dkl_row + T - 1] : 0.f;
            uint32_t ph_m = 0, ph_end = 0;
            for (int t = T - 1; t >= 0; --t) {
                const size_t iA = (size_t)r.rA * T + t, iB = (size_t)r.rB * T + t;
                const float dkl_cur = dkl_next;
                if (t > 0 && dkl_src != nullptr) dkl_next = dkl_src[dkl_row + t - 1];
                // ---- everything that does NOT depend on the carried d stoch_l runs BEFORE the hand-over barrier, in the window
                // where this warp used to idle: the two flat log-softmaxes, the mixture responsibilities, the KL / upstream part of
                // d post_probs_l, exp(ls) for the log-softmax backward and ELU' of the two hiddens ----------------------------
                FZ_TS(0);
                cp_async_wait<1>();  // LA, LV, post_l, prior_l of step t have landed (the rest may still be in flight)
                __syncwarp();
                FZ_TS(1);
                float q[2][4], dzl_pre[2][4], lsa[2][4], lsv[2][4], ra[2][4], rv[2][4];
                float dlg_lp[2][4];  // d logits of the lower prior head (this warp owns the whole "l" side of the KL term)
                {
                    float pp[2][4], dpp[2][4], la[2][4], lv[2][4];
                    load_staged<2, true>(q, stPRL, 32, 0, r.g, r.t);
                    load_staged<2, true>(pp, stPRL, 32, 16, r.g, r.t);
                    zero_c<2>(dzl_pre), zero_c<2>(dpp);
                    add_global<2>(dzl_pre, p.d_post_probs_l, iA * 16, iB * 16, r.t);
                    add_global<2>(dpp, p.d_prior_probs_l, iA * 16, iB * 16, r.t);
                    add_global<2>(dpp, p.d_prior_stoch_l, iA * 16, iB * 16, r.t);
                    {   // (the shuffles stay outside the branch: inside it they compile to a guarded collective, ~20 instructions each)
                        const float dkl[2] = {__shfl_sync(FULL, dkl_cur, (lane & ~3) + 0), __shfl_sync(FULL, dkl_cur, (lane & ~3) + 1)};
                        if (p.d_kl_l != nullptr) kl_rows_bwd<true>(q, pp, dkl, p.kl_wq, p.kl_wp, dzl_pre, dpp);
                    }
                    softmax_groups_bwd<KL>(pp, dpp, dlg_lp);
                    load_op<2>(la, svop, mts::LA, r.g, r.t);
                    load_op<2>(lv, svop, mts::LV, r.g, r.t);
                    // lsa / lsv = softmax(la) / softmax(lv) (all the flat log-softmax backward needs), ra / rv = the mixture
                    // responsibilities, evaluated in the probability domain (frag.cuh)
                    mopoe_responsibilities_fast(la, lv, lsa, lsv, ra, rv);
                }
                __syncwarp();
                FZ_TS(2);
                stage_logits(t - 1);  // LA, LV and the probability rows are in registers: refill them
                cp_async_wait<1>();   // hiddens of step t have landed (the embedding images were written by this warp itself)
                __syncwarp();
                float eluA[4][4], eluV[4][4];  // ELU'(hidden) of the two modality heads
                load_op<4>(eluA, svop, mts::A_HID, r.g, r.t);
                load_op<4>(eluV, svop, mts::V_HID, r.g, r.t);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        eluA[nt][j] = elu_grad_from_out(eluA[nt][j]);
                        eluV[nt][j] = elu_grad_from_out(eluV[nt][j]);
                    }
                // ---- lower prior head (moved here from the core warp, which is the longest instruction stream of a tile: this warp
                // waited a third of its time at the hand-over barrier).  It depends on nothing that is carried, so it runs before the
                // hand-over; its d deter_l joins this warp's contribution (XDDL).  This warp's dY columns (LPL, LP1, LA, A1, LV, V1)
                // were last read by the end-of-step MMAs of step t+1 (and by its own MMAs, waited for at the end of that step) ------
                if (t < T - 1) FZ_WAIT(&bars[fz2::BAR_END], ph_end), ph_end ^= 1;
                float ddl[4][4];
                zero_c<4>(ddl);
                {
                    float hid[4][4];
                    AFrag<NS, 2> f1;
                    load_op<4>(hid, svop, mts::LP_HID, r.g, r.t);
                    head_bwd_op(dlg_lp, wblk<NS>(W, mt::T_LP2), hid, dy, fz::Y_LPL, fz::Y_LP1, f1, r, lane);
                    gemm<NS, 2, 4>(ddl, f1, wblk<NS>(W, mt::T_LP1), lane);
                }
                // ---- the recurrence's critical section: d stoch_l(t) -> ... -> contribution to d deter_l(t) --------------------
                FZ_TS(3);
                nbar_sync(bar_y);  // d stoch_l of step t is in XDZL
                FZ_TS(4);
                float dla[2][4], dlv[2][4];
                {
                    float dzl[2][4], dm[2][4];
                    xch_load<2>(dzl, xdzl, lane);
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int j = 0; j < 4; ++j) dzl[nt][j] += dzl_pre[nt][j];
                    softmax_groups_bwd<KL>(q, dzl, dm);
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            ra[nt][j] *= dm[nt][j];
                            rv[nt][j] *= dm[nt][j];
                        }
                    // backward of the flat log-softmax with softmax(x) at hand: dx = dls - softmax(x) * sum(dls)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float sa[2] = {ra[0][2 * h] + ra[0][2 * h + 1], ra[1][2 * h] + ra[1][2 * h + 1]};
                        float sv[2] = {rv[0][2 * h] + rv[0][2 * h + 1], rv[1][2 * h] + rv[1][2 * h + 1]};
                        group_reduce<16, false>(sa);
                        group_reduce<16, false>(sv);
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                dla[nt][2 * h + e] = ra[nt][2 * h + e] - lsa[nt][2 * h + e] * sa[0];
                                dlv[nt][2 * h + e] = rv[nt][2 * h + e] - lsv[nt][2 * h + e] * sv[0];
                            }
                    }
                }
                FZ_TS(5);
                FZ_TS(6);
                AFrag<NS, 2> f1a, f1v;
                if constexpr (PROJ) {
                    head_bwd_elu(dla, wblk<NS>(W, mt::T_A2), eluA, dy, fz::Y_LA, fz::Y_A1, f1a, r, lane, p.d_embed_a + iA * 32, p.d_embed_a + iB * 32);
                    head_bwd_elu(dlv, wblk<NS>(W, mt::T_V2), eluV, dy, fz::Y_LV, fz::Y_V1, f1v, r, lane, p.d_embed_v + iA * 32, p.d_embed_v + iB * 32);
                } else {
                    head_bwd_elu(dla, wblk<NS>(W, mt::T_A2), eluA, dy, fz::Y_LA, fz::Y_A1, f1a, r, lane);
                    head_bwd_elu(dlv, wblk<NS>(W, mt::T_V2), eluV, dy, fz::Y_LV, fz::Y_V1, f1v, r, lane);
                }
                gemm<NS, 2, 4>(ddl, f1a, wblk<NS>(W, mt::T_A1H), lane);
                gemm<NS, 2, 4>(ddl, f1v, wblk<NS>(W, mt::T_V1H), lane);
                xch_store<4>(ddl, xddl, lane);
                FZ_FENCE();
                nbar_arrive(bar_x, X_COUNT);  // XDDL and this warp's dY columns are complete (and visible to the async proxy)
                FZ_TS(7);
                // ---- off the critical path: this warp's weight-gradient MMAs, the embedding gradients, next step's staging -------
                __syncwarp();
                if (FZ_MMA && lane == 0) {
                    umma_acc(tmem + fz2::T_M, s_sv + desc_off(12 * fz2::CH), s_dy + desc_off((fz::Y_LA / 8) * fz2::CH), 32);
                    // [lp hid | ..] x [LPL]: the first 16 columns of the T_E tile (rows 32..127 of the window -- hp / hq / a hiddens,
                    // possibly mid-refill by the core warp -- land in accumulator rows that are never read back)
                    umma_acc(tmem + fz2::T_E, s_sv, s_dy + desc_off((fz::Y_LPL / 8) * fz2::CH), 16);
                    if constexpr (WPT == 2) umma_acc(tmem + fz2::T_EMB, s_sv + desc_off(24 * fz2::CH), s_dy + desc_off((fz::Y_A1 / 8) * fz2::CH), 64);
                    umma_commit(&bars[fz2::BAR_M]);
                }
                if constexpr (WPT == 2) {  // three warps per tile: the aux warp computes the embedding gradients from the dY image
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        float de[8][4];
                        zero_c<8>(de);
                        gemm<NS, 2, 8>(de, m == 0 ? f1a : f1v, wblk<NS>(W, m == 0 ? mt::T_A1E : mt::T_V1E), lane);
                        float* dE = m == 0 ? p.d_embed_a : p.d_embed_v;
                        store_c<8>(de, dE + iA * 64, dE + iB * 64, r);
                    }
                }
                FZ_TS(8);
                FZ_WAIT(&bars[fz2::BAR_M], ph_m), ph_m ^= 1;  // own MMAs done: the hiddens may be rewritten
                FZ_TS(9);
                stage_rest(t - 1);
                if constexpr (WPT == 2) {
                    embed_prefetch(t - 2);
                    // (requesting the fp32 embeddings before the embedding-gradient GEMMs and converting them here was measured: the
                    // stall moves into the GEMMs, 245 registers, 0.78 -> 0.815 ms at the bench size -- not kept)
                    if (t > 0) embed_images(t - 1);
                }
                FZ_TS(10);
            }
            cp_async_wait_all();
        } else if (role == 0) {
            // ============================== core warp: prior / higher heads + the two cells ===============================
            const float keep_l = 1.f - p.inv_tau_l, keep_h = 1.f - p.inv_tau_h;
            float* stDF = reinterpret_cast<float*>(my + fz2::DF);
            float* stFT = reinterpret_cast<float*>(my + fz2::FT);
            float* stPR = reinterpret_cast<float*>(my + fz2::PR) - bst::PR;  // bstage_pr adds bst::PR itself
            uint32_t ph_df = 0, ph_ft = 0, ph_e = 0, ph_end = 0;
                        auto stage_hid = [&](int t) {  // hp, hq hiddens: free once the E-group MMA has completed
                if (t >= 0) {
                    stage_chunks_tiled(svop + 4 * fz2::CH, tile_rec, 4, 8, t, lane);
                }
                cp_async_commit();
            };
            bulk_rows(stDF, bst::DF_LD, reinterpret_cast<const char*>(p.d_feature), 384, bst::DF_BYTES, row0, p.B, T, T - 1, &bars[fz2::BAR_DF], lane);
            if constexpr (WPT == 2)
                bulk_rows(stFT, bst::DF_LD, reinterpret_cast<const char*>(p.feature), ft_pitch, 384, row0, p.B, T, T - 1, &bars[fz2::BAR_FT], lane);
            // the four probability tensors of a step (one cp.async group): lane -> chunk column c8 of the rows rq + 4j; as bstage_pr
            // (mtrssm_common.cuh) with the row offsets, sources and destination hoisted out of the time loop
            bstage_pr<ldP>(stPR, p, row0, T - 1, lane);  // cp.async groups per step, in issue order: PR(t-1) | HID(t-1)
            stage_hid(T - 1);
            const float* dkl_src = (r.t < 2) ? p.d_kl_h : p.d_kl_l;  // quad lanes: 0 kl_h row A, 1 kl_h row B, 2 kl_l row A, 3 kl_l row B
            const size_t dkl_row = (size_t)((r.t & 1) ? r.rB : r.rA) * T;
            float dkl_next = dkl_src != nullptr ? dkl_src[dkl_row + T - 1] : 0.f;
            float ddl[4][4], ddh[4][4], dul[4][4], duh[4][4], dzh[2][4];
            zero_c<4>(dul), zero_c<4>(duh);
            // upstream d_feature of step T-1: d stoch_l goes to the mod warp, the rest seeds the carried gradients
            {
                float dzl0[2][4];
                mbar_wait(&bars[fz2::BAR_DF], ph_df), ph_df ^= 1;
                load_staged<4, false>(ddh, stDF, bst::DF_LD, 0, r.g, r.t);
                load_staged<2, false>(dzh, stDF, bst::DF_LD, 32, r.g, r.t);
                load_staged<4, false>(ddl, stDF, bst::DF_LD, 48, r.g, r.t);
                load_staged<2, false>(dzl0, stDF, bst::DF_LD, 80, r.g, r.t);
                xch_store<2>(dzl0, xdzl, lane);
                nbar_arrive(bar_y);
                __syncwarp();
                if (T > 1)
                    bulk_rows(stDF, bst::DF_LD, reinterpret_cast<const char*>(p.d_feature), 384, bst::DF_BYTES, row0, p.B, T, T - 2,
                              &bars[fz2::BAR_DF], lane);
            }
            FZ_GS(20 + 2 * (grp / (int)gridDim.x));
            for (int t = T - 1; t >= 0; --t) {
                const size_t iA = (size_t)r.rA * T + t, iB = (size_t)r.rB * T + t;
                const float dkl_cur = dkl_next;
                if (t > 0 && dkl_src != nullptr) dkl_next = dkl_src[dkl_row + t - 1];
                float hid[4][4];
                FZ_TS(0);
                cp_async_wait_all();  // PR(t) and the hiddens of step t have landed
                __syncwarp();
                // the end-of-step MMAs of step t+1 are done with dY, Dop and Zop
                if (t < T - 1) FZ_WAIT(&bars[fz2::BAR_END], ph_end), ph_end ^= 1;
                FZ_TS(1);
                // the action of step t+1 (X operand of the l cell's weight gradient, paired below with dY of step t+1): lanes with
                // t < 2 own its columns 4t .. 4t+3; column 8 of the 16-column block is the ones column
                float actc[2][4];
                zero_c<2>(actc);
                if constexpr (WPT == 2) load_action_columns(actc, p.actions, iA + 1, iB + 1, A, t + 1 < T, r);
                // (the lower prior head runs in the mod warp)
                FZ_TS(2);
                // ---- higher layer: posterior + prior heads ----------------------------------------------------------------
                {
                    float q[2][4], pp[2][4], dpp[2][4];
                    load_staged<2, true>(q, stPR + bst::PR, 64, 0, r.g, r.t);
                    load_staged<2, true>(pp, stPR + bst::PR, 64, 32, r.g, r.t);
                    __syncwarp();  // every lane is done with PR: refill it for the next (earlier) step
                    bstage_pr<ldP>(stPR, p, row0, t - 1, lane);
                    add_global<2>(dzh, p.d_post_probs_h, iA * 16, iB * 16, r.t);
                    zero_c<2>(dpp);
                    add_global<2>(dpp, p.d_prior_probs_h, iA * 16, iB * 16, r.t);
                    add_global<2>(dpp, p.d_prior_stoch_h, iA * 16, iB * 16, r.t);
                    {
                        const float dkl[2] = {__shfl_sync(FULL, dkl_cur, (lane & ~3) + 0), __shfl_sync(FULL, dkl_cur, (lane & ~3) + 1)};
                        if (p.d_kl_h != nullptr) kl_rows_bwd<true>(q, pp, dkl, p.kl_wq, p.kl_wp, dzh, dpp);
                    }
                    float dlg[2][4];
                    AFrag<NS, 2> f1;
                    softmax_groups_bwd<KH>(q, dzh, dlg);
                    load_op<4>(hid, svop, mts::HQ_HID, r.g, r.t);
                    head_bwd_op(dlg, wblk<NS>(W, mt::T_HQ2), hid, dy, fz::Y_HQL, fz::Y_HQ1, f1, r, lane);
                    gemm<NS, 2, 4>(ddl, f1, wblk<NS>(W, mt::T_HQ1L), lane);
                    gemm<NS, 2, 4>(ddh, f1, wblk<NS>(W, mt::T_HQ1H), lane);
                    softmax_groups_bwd<KH>(pp, dpp, dlg);
                    load_op<4>(hid, svop, mts::HP_HID, r.g, r.t);
                    head_bwd_op(dlg, wblk<NS>(W, mt::T_HP2), hid, dy, fz::Y_HPL, fz::Y_HP1, f1, r, lane);
                    gemm<NS, 2, 4>(ddh, f1, wblk<NS>(W, mt::T_HP1), lane);
                }
                FZ_TS(3);
                // second-layer weight gradients of this warp's two heads
                FZ_FENCE();
                __syncwarp();
                if (FZ_MMA && lane == 0) {
                    umma_acc(tmem + fz2::T_E + 16, s_sv, s_dy + desc_off((fz::Y_HPL / 8) * fz2::CH), 32);  // [.. | hp | hq hid | ..] x [HPL | HQL]
                    umma_commit(&bars[fz2::BAR_E]);
                }
                FZ_TS(4);
                // ---- the two leaky integrators --------------------------------------------------------------------------
                float dh[4][4], dl[4][4], ph[4][4], pl[4][4];
                mbar_wait(&bars[fz2::BAR_FT], ph_ft), ph_ft ^= 1;  // the feature row of step t has landed
                load_staged<4, false>(dh, stFT, bst::DF_LD, 0, r.g, r.t);
                load_staged<4, false>(dl, stFT, bst::DF_LD, 48, r.g, r.t);
                if constexpr (WPT == 2) {  // three warps per tile: the aux warp builds the X-operand images and refills the stage
                    float zh[2][4], zl[2][4];
                    load_staged<2, false>(zh, stFT, bst::DF_LD, 32, r.g, r.t);
                    load_staged<2, false>(zl, stFT, bst::DF_LD, 80, r.g, r.t);
                    __syncwarp();
                    if (t > 0)
                        bulk_rows(stFT, bst::DF_LD, reinterpret_cast<const char*>(p.feature), ft_pitch, 384, row0, p.B, T, t - 1, &bars[fz2::BAR_FT], lane);
                    // X operands: bf16 [d_l | d_h | z_l | z_h](t) and [action(t+1) | ones]
                    store_op<4>(dl, dop, 0, r);
                    store_op<4>(dh, dop, 32, r);
                    store_op<2>(zl, zop, 0, r);
                    store_op<2>(zh, zop, 16, r);
                    store_op<2>(actc, zop, 32, r);
                }
                FZ_TS(5);
                FZ_WAIT(&bars[fz2::BAR_E], ph_e), ph_e ^= 1;  // the hiddens may be refilled
                stage_hid(t - 1);
                FZ_TS(6);
                // ---- the higher cell depends on nothing the mod warp produces: all of it runs BEFORE the hand-over barrier ------
                float dzh_n[2][4], ddh_n[4][4];
                AFrag<NS, 2> fh;
                zero_c<4>(ddh_n), zero_c<2>(dzh_n);
                if (p.d_hidden_h != nullptr) {  // upstream gradients of the hidden outputs (u_t of the two cells) join the carried d u
                    const size_t hA = ((size_t)r.rA * T + t) * 32, hB = ((size_t)r.rB * T + t) * 32;
                    add_global<4>(duh, p.d_hidden_h, hA, hB, r.t);
                    add_global<4>(dul, p.d_hidden_l, hA, hB, r.t);
                }
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float gh = duh[nt][j] + ddh[nt][j] * (1.f - dh[nt][j] * dh[nt][j]);
                        ph[nt][j] = gh * p.inv_tau_h;
                        duh[nt][j] = gh * keep_h;
                        dl[nt][j] = 1.f - dl[nt][j] * dl[nt][j];  // tanh' of the lower cell, ready for the critical section
                    }
                // the cells' dY is double buffered: step t's lands in buffer t & 1 and is multiplied at step t-1 with that step's
                // state (= this step's INPUT state), so no copies of the previous state are needed anywhere
                store_op<4>(ph, dy, fz::Y_H + 64 * (t & 1), r);
                to_afrag<NS, 2>(fh, ph);
                gemm<NS, 2, 4>(ddh_n, fh, wblk<NS>(W, mt::T_H_D2H), lane);
                gemm<NS, 2, 2>(dzh_n, fh, wblk<NS>(W, mt::T_H_IN), lane);
                float g2[2][4];  // upstream d stoch_l of step t-1
                zero_c<2>(g2);
                if (t > 0) {
                    mbar_wait(&bars[fz2::BAR_DF], ph_df), ph_df ^= 1;  // d_feature(t-1) has landed
                    load_staged<2, false>(g2, stDF, bst::DF_LD, 80, r.g, r.t);
                }
                // ---- the recurrence's critical section: mod's d deter_l(t) -> lower cell -> d stoch_l(t-1) back to the mod warp --
                FZ_TS(7);
                nbar_sync(bar_x, X_COUNT);  // the mod warp's dY columns and its contribution to d deter_l are complete
                FZ_TS(8);
                {
                    float c[4][4];
                    xch_load<4>(c, xddl, lane);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float gl = dul[nt][j] + (ddl[nt][j] + c[nt][j]) * dl[nt][j];
                            pl[nt][j] = gl * p.inv_tau_l;
                            dul[nt][j] = gl * keep_l;
                        }
                }
                AFrag<NS, 2> fl;
                to_afrag<NS, 2>(fl, pl);
                float dzl[2][4];
                zero_c<2>(dzl);
                gemm<NS, 2, 2>(dzl, fl, wblk<NS>(W, mt::T_L_IN_ZL), lane);
                if (t > 0) {  // hand d stoch_l of step t-1 (+ its upstream part) to the mod warp
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int j = 0; j < 4; ++j) dzl[nt][j] += g2[nt][j];
                    xch_store<2>(dzl, xdzl, lane);
                    nbar_arrive(bar_y);
                }
                FZ_TS(9);
                // ---- off the critical path: the end-of-step weight-gradient MMAs and the rest of the lower cell -------------------
                store_op<4>(pl, dy, fz::Y_L + 64 * (t & 1), r);
                FZ_FENCE();
                if constexpr (WPT == 3) {
                    nbar_arrive(bar_z);  // every dY column of this warp is complete and it is done with the feature-row stage
                } else {
                    __syncwarp();
                    if (FZ_MMA && lane == 0) {
                        umma_acc(tmem + fz2::T_D1, s_dop, s_dy + desc_off((fz::Y_HQ1 / 8) * fz2::CH), 160);
                        if (t < T - 1) umma_acc(tmem + fz2::T_C, s_dop, s_dy + desc_off(((fz::Y_L + 64 * ((t + 1) & 1)) / 8) * fz2::CH), 64);
                        umma_acc(tmem + fz2::T_B, s_dy, s_ones, 16);                     // dY columns   0..127
                        umma_acc(tmem + fz2::T_B + 16, s_dy + desc_off(16 * fz2::CH), s_ones, 16);  // dY columns 128..239 (+ junk)
                        umma_commit(&bars[fz2::BAR_END]);
                    }
                }
                FZ_TS(10);
                zero_c<4>(ddl);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) ddh[nt][j] = ddh_n[nt][j];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dzh[nt][j] = dzh_n[nt][j];
                gemm<NS, 2, 4>(ddl, fl, wblk<NS>(W, mt::T_L_D2H), lane);
                gemm<NS, 2, 2>(dzh, fl, wblk<NS>(W, mt::T_L_IN_ZH), lane);
                if (WPT == 2 && p.d_actions != nullptr) {  // three warps per tile: the aux warp, from the dY image
                    float da[2][4];
                    zero_c<2>(da);
                    gemm<NS, 2, 2>(da, fl, wblk<NS>(W, mt::T_L_IN_A), lane);
                    store_c_partial(da, p.d_actions + iA * A, p.d_actions + iB * A, r, A);
                }
                if (t > 0) {  // the rest of d_feature(t-1) seeds the carried gradients of the next step
                    float g4[4][4], g2[2][4];
                    load_staged<4, false>(g4, stDF, bst::DF_LD, 0, r.g, r.t);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                        for (int j = 0; j < 4; ++j) ddh[nt][j] += g4[nt][j];
                    load_staged<2, false>(g2, stDF, bst::DF_LD, 32, r.g, r.t);
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int j = 0; j < 4; ++j) dzh[nt][j] += g2[nt][j];
                    load_staged<4, false>(g4, stDF, bst::DF_LD, 48, r.g, r.t);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                        for (int j = 0; j < 4; ++j) ddl[nt][j] += g4[nt][j];
                    __syncwarp();
                    if (t > 1)
                        bulk_rows(stDF, bst::DF_LD, reinterpret_cast<const char*>(p.d_feature), 384, bst::DF_BYTES, row0, p.B, T, t - 2,
                                  &bars[fz2::BAR_DF], lane);
                } else {  // t == 0: the gradients w.r.t. the initial state
                    store_c<4>(ddh, p.d_deter_h0 + (size_t)r.rA * 32, p.d_deter_h0 + (size_t)r.rB * 32, r);
                    store_c<4>(ddl, p.d_deter_l0 + (size_t)r.rA * 32, p.d_deter_l0 + (size_t)r.rB * 32, r);
                    store_c<4>(duh, p.d_hidden_h0 + (size_t)r.rA * 32, p.d_hidden_h0 + (size_t)r.rB * 32, r);
                    store_c<4>(dul, p.d_hidden_l0 + (size_t)r.rA * 32, p.d_hidden_l0 + (size_t)r.rB * 32, r);
                    store_c<2>(dzh, p.d_stoch_h0 + (size_t)r.rA * 16, p.d_stoch_h0 + (size_t)r.rB * 16, r);
                    store_c<2>(dzl, p.d_stoch_l0 + (size_t)r.rA * 16, p.d_stoch_l0 + (size_t)r.rB * 16, r);
                }
                FZ_TS(11);
            }
            FZ_GS(21 + 2 * (grp / (int)gridDim.x));
            if constexpr (WPT == 2) pair_initial_state(ph_end);
            cp_async_wait_all();
        } else {
            // ============================== aux warp (WPT == 3): operand images, MMAs, input gradients ======================
            float* stFT = reinterpret_cast<float*>(my + fz2::FT);
            uint32_t ph_ft = 0, ph_end = 0;
            // lanes 0..15: one 256-byte row of each embedding into L2.  (Per-lane `prefetch.global.L2` of the same lines is 1.7 % SLOWER
            // at the bench batch although it saves ~260 issue slots per step: profiles/r2_ab_variants.txt)
            auto embed_prefetch = [&](int t) {
                if (t >= 0 && lane < 16) {
                    const size_t j = ((size_t)min(row0 + lane, p.B - 1) * T + t) * 64;
                    prefetch_bulk_l2(p.embed_a + j, 256);
                    prefetch_bulk_l2(p.embed_v + j, 256);
                }
            };
            bulk_rows(stFT, bst::DF_LD, reinterpret_cast<const char*>(p.feature), ft_pitch, 384, row0, p.B, T, T - 1, &bars[fz2::BAR_FT], lane);
            constexpr bool emb = !PROJ;  // obs_projected: no embedding operand images, no embedding MMA, no d_embed GEMMs here
            if (emb) embed_prefetch(T - 2);
            // the fp32 embeddings of a step are requested one step ahead (registers) and converted at the top of their step
            float ea[8][4], ev[8][4];
            zero_c<8>(ea), zero_c<8>(ev);
            if (emb) {
                const size_t jA = ((size_t)r.rA * T + T - 1) * 64, jB = ((size_t)r.rB * T + T - 1) * 64;
                load_c<8>(ea, p.embed_a + jA, p.embed_a + jB, r.t);
                load_c<8>(ev, p.embed_v + jA, p.embed_v + jB, r.t);
            }
            // [action(t+1) | ones] columns, fetched one step ahead as well (a load next to its use costs a DRAM round trip per step)
            float actc[2][4];
            zero_c<2>(actc);
            load_action_columns(actc, p.actions, 0, 0, A, false, r);  // step T-1 pairs with no action: the ones column only
            for (int t = T - 1; t >= 0; --t) {
                const size_t iA = (size_t)r.rA * T + t, iB = (size_t)r.rB * T + t;
                FZ_TS(0);
                // the MMAs of step t+1 are done with the operand images
                if (t < T - 1) FZ_WAIT(&bars[fz2::BAR_END], ph_end), ph_end ^= 1;
                FZ_TS(1);
                store_op<2>(actc, zop, 32, r);
                if (t > 0) {  // the action of step t, used by step t-1
                    zero_c<2>(actc);
                    load_action_columns(actc, p.actions, iA, iB, A, true, r);
                }
                if (emb) {
                    store_op<8>(ea, svop, 24 * 8, r);
                    store_op<8>(ev, svop, 32 * 8, r);
                    if (t > 0) {
                        const size_t jA = (iA - 1) * 64, jB = (iB - 1) * 64;
                        load_c<8>(ea, p.embed_a + jA, p.embed_a + jB, r.t);
                        load_c<8>(ev, p.embed_v + jA, p.embed_v + jB, r.t);
                    }
                    embed_prefetch(t - 2);
                }
                FZ_TS(2);
                {   // X operands: bf16 [d_l | d_h | z_l | z_h](t) and [action(t+1) | ones]
                    float c4[4][4], c2[2][4];
                    mbar_wait(&bars[fz2::BAR_FT], ph_ft), ph_ft ^= 1;  // the feature row of step t has landed
                    load_staged<4, false>(c4, stFT, bst::DF_LD, 48, r.g, r.t);
                    store_op<4>(c4, dop, 0, r);
                    load_staged<4, false>(c4, stFT, bst::DF_LD, 0, r.g, r.t);
                    store_op<4>(c4, dop, 32, r);
                    load_staged<2, false>(c2, stFT, bst::DF_LD, 80, r.g, r.t);
                    store_op<2>(c2, zop, 0, r);
                    load_staged<2, false>(c2, stFT, bst::DF_LD, 32, r.g, r.t);
                    store_op<2>(c2, zop, 16, r);
                }
                FZ_TS(3);
                // ---- the mod warp's dY columns: embedding gradients d e = dY1 . W1[:, 32:] (mopoe_mmtrssm/core.py:259-260) ------
                nbar_sync(bar_x, X_COUNT);
                FZ_TS(4);
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    if (!emb) break;
                    float dhid[4][4], de[8][4];
                    AFrag<NS, 2> f1;
                    load_op<4>(dhid, dy, m == 0 ? fz::Y_A1 : fz::Y_V1, r.g, r.t);
                    to_afrag<NS, 2>(f1, dhid);
                    zero_c<8>(de);
                    gemm<NS, 2, 8>(de, f1, wblk<NS>(W, m == 0 ? mt::T_A1E : mt::T_V1E), lane);
                    float* dE = m == 0 ? p.d_embed_a : p.d_embed_v;
                    store_c<8>(de, dE + iA * 64, dE + iB * 64, r);
                }
                FZ_TS(5);
                // ---- the core warp's dY columns are complete and it is done with the feature-row stage ---------------------------
                nbar_sync(bar_z);
                FZ_TS(6);
                float pl[4][4];
                load_op<4>(pl, dy, fz::Y_L + 64 * (t & 1), r.g, r.t);
                FZ_FENCE();
                __syncwarp();
                if (FZ_MMA && lane == 0) {
                    if (emb) umma_acc(tmem + fz2::T_EMB, s_sv + desc_off(24 * fz2::CH), s_dy + desc_off((fz::Y_A1 / 8) * fz2::CH), 64);
                    umma_acc(tmem + fz2::T_D1, s_dop, s_dy + desc_off((fz::Y_HQ1 / 8) * fz2::CH), 160);
                    if (t < T - 1) umma_acc(tmem + fz2::T_C, s_dop, s_dy + desc_off(((fz::Y_L + 64 * ((t + 1) & 1)) / 8) * fz2::CH), 64);
                    umma_acc(tmem + fz2::T_B, s_dy, s_ones, 16);                     // dY columns   0..127
                    umma_acc(tmem + fz2::T_B + 16, s_dy + desc_off(16 * fz2::CH), s_ones, 16);  // dY columns 128..239 (+ junk)
                    umma_commit(&bars[fz2::BAR_END]);
                }
                FZ_TS(7);
                if (t > 0)
                    bulk_rows(stFT, bst::DF_LD, reinterpret_cast<const char*>(p.feature), ft_pitch, 384, row0, p.B, T, t - 1, &bars[fz2::BAR_FT], lane);
                if (p.d_actions != nullptr) {  // d a = dY_l . W_in[:, :A] (l_rnn._input2h, :283-284)
                    AFrag<NS, 2> fl;
                    to_afrag<NS, 2>(fl, pl);
                    float da[2][4];
                    zero_c<2>(da);
                    gemm<NS, 2, 2>(da, fl, wblk<NS>(W, mt::T_L_IN_A), lane);
                    store_c_partial(da, p.d_actions + iA * A, p.d_actions + iB * A, r, A);
                }
                FZ_TS(8);
            }
            pair_initial_state(ph_end);
        }
    }
    if (grp + (int)gridDim.x < ngroups) {
        // another group follows: put the tile's shared memory and mbarriers back into their initial state (every MMA and bulk
        // copy of this group has completed: both warps waited on BAR_END / their copies above)
        const int bar_g = WPT == 2 ? 9 + tile : 2 + WPT * tile;  // three warps: the Y barrier's id is idle between groups
        nbar_sync(bar_g, 32 * WPT);
        for (int i = lane + 32 * role; i < fz2::BYTES / 16; i += 32 * WPT) reinterpret_cast<uint4*>(my)[i] = make_uint4(0u, 0u, 0u, 0u);
        if (role == 0 && lane == 0) {
#pragma unroll
            for (int i = 0; i < fz2::NBAR; ++i) {
                asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars[i])) : "memory");
                mbar_init(&bars[i], 1);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        nbar_sync(bar_g, 32 * WPT);
    }
    FZ_GS(3 + 2 * (grp / (int)gridDim.x));
    }  // tile groups
    FZ_GS(40);
    // ---- epilogue: TMEM accumulators -> global weight gradients (one atomicAdd per element per CTA) -----------------------
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        // a warp may read the TMEM lane quarter (warp % 4); with 8 warps, warps w and w + 4 split a quarter's 16-column chunks
        const int quarter = warp & 3, part = warp >> 2, nparts = nthr >> 7;
        const int L = 32 * quarter + lane;  // TMEM lane read by this thread
        int chunk = 0;
        for (int i = 0; i < ft.n; ++i) {
            const FusedFlush f = ft.e[i];
            if (f.lane0 + f.nlanes <= 32 * quarter || f.lane0 >= 32 * quarter + 32) continue;  // warp-uniform
            for (int c0 = 0; c0 < f.ncols; c0 += 16, ++chunk) {
                if (chunk % nparts != part) continue;  // warp-uniform
                uint32_t v[16];
                const uint32_t taddr = tmem + ((uint32_t)(32 * quarter) << 16) + f.tcol + c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                      "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (L >= f.lane0 && L < f.lane0 + f.nlanes) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < f.ncols) atomicAdd(f.dst + (size_t)(c0 + j) * f.ld + (L - f.lane0) + f.koff, __uint_as_float(v[j]));
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    FZ_GS(41);
}

// ---- host side: which TMEM block goes where ------------------------------------------------------------------------------
static void add_flush(FusedFlushTable& t, float* dst, int tcol, int ncols, int lane0, int nlanes, int ld, int koff) {
    if (dst == nullptr || nlanes <= 0) return;
    FusedFlush& f = t.e[t.n++];
    f.dst = dst, f.tcol = tcol, f.ncols = ncols, f.lane0 = lane0, f.nlanes = nlanes, f.ld = ld, f.koff = koff;
}

cudaError_t launch_mtrssm_bwd_fused(const MtrssmBwdArgs& a, const RssmMtrssmWeightGrads& g, cudaStream_t s) {
    using namespace fz;
    const int A = a.A, LDIN = A + 32;
    // which (lane range, column range) block of a TMEM tile is which weight gradient
    FusedFlushTable u{};
    {
        namespace z = fz2;
        add_flush(u, g.lp_w2, z::T_E + 0, 16, 0, 32, 32, 0);
        add_flush(u, g.hp_w2, z::T_E + 16, 16, 32, 32, 32, 0);
        add_flush(u, g.hq_w2, z::T_E + 32, 16, 64, 32, 32, 0);
        add_flush(u, g.au_w2, z::T_M + 0, 16, 0, 32, 32, 0);
        add_flush(u, g.vi_w2, z::T_M + 16, 16, 32, 32, 32, 0);
        if (!a.obs_projected) {  // projected partials: the columns W1[:, 32:] belong to the caller's GEMM
            add_flush(u, g.au_w1, z::T_EMB + 0, 32, 0, 64, 96, 32);
            add_flush(u, g.vi_w1, z::T_EMB + 32, 32, 64, 64, 96, 32);
        }
        add_flush(u, g.hq_w1, z::T_D1 + 0, 32, 0, 64, 64, 0);
        add_flush(u, g.lp_w1, z::T_D1 + 32, 32, 0, 32, 32, 0);
        add_flush(u, g.au_w1, z::T_D1 + 64, 32, 0, 32, 96, 0);
        add_flush(u, g.vi_w1, z::T_D1 + 96, 32, 0, 32, 96, 0);
        add_flush(u, g.hp_w1, z::T_D1 + 128, 32, 32, 32, 32, 0);
        add_flush(u, g.l_d2h_w, z::T_C + 0, 32, 0, 32, 32, 0);
        add_flush(u, g.h_d2h_w, z::T_C + 32, 32, 32, 32, 32, 0);
        add_flush(u, g.l_in_w, z::T_C + 0, 32, 64, 32, LDIN, A);
        add_flush(u, g.h_in_w, z::T_C + 32, 32, 80, 16, 16, 0);
        add_flush(u, g.l_in_w, z::T_C + 0, 32, 96, A, LDIN, 0);
        add_flush(u, g.lp_b2, z::T_B, 1, Y_LPL, 16, 0, 0);
        add_flush(u, g.hp_b2, z::T_B, 1, Y_HPL, 16, 0, 0);
        add_flush(u, g.hq_b2, z::T_B, 1, Y_HQL, 16, 0, 0);
        add_flush(u, g.au_b2, z::T_B, 1, Y_LA, 16, 0, 0);
        add_flush(u, g.vi_b2, z::T_B, 1, Y_LV, 16, 0, 0);
        add_flush(u, g.hq_b1, z::T_B, 1, Y_HQ1, 32, 0, 0);
        add_flush(u, g.lp_b1, z::T_B, 1, Y_LP1, 16, 0, 0);
        add_flush(u, g.lp_b1, z::T_B + 16, 1, 0, 16, 0, 16);
        add_flush(u, g.au_b1, z::T_B + 16, 1, Y_A1 - 128, 32, 0, 0);
        add_flush(u, g.vi_b1, z::T_B + 16, 1, Y_V1 - 128, 32, 0, 0);
        add_flush(u, g.hp_b1, z::T_B + 16, 1, Y_HP1 - 128, 32, 0, 0);
        add_flush(u, g.l_d2h_b, z::T_C + 0, 32, 104, 1, 1, 0);  // ones row (lane 104) of the cells' tile: column n -> bias n
        add_flush(u, g.l_in_b, z::T_C + 0, 32, 104, 1, 1, 0);
        add_flush(u, g.h_d2h_b, z::T_C + 32, 32, 104, 1, 1, 0);
        add_flush(u, g.h_in_b, z::T_C + 32, 32, 104, 1, 1, 0);
    }
    // warps per tile: 3 (core, mod, aux) by default; RSSM_BWD_TWO_WARP=1 selects the two-warp kernel (A/B measurements)
    const int wpt = getenv("RSSM_BWD_TWO_WARP") != nullptr ? 2 : 3;
    // tiles per CTA: 4 fill an SM's shared memory; small batches use 2, or 1 (three-warp kernel: + a helper warp, the TMEM
    // read-back needs four warps) when every tile can have an SM of its own -- two tiles sharing an SM run 20 % slower per step
    const int tiles = (a.B + 15) / 16, tpc = tiles > 2 * 148 ? 4 : (tiles > 148 || tiles < 2 || wpt != 3 || getenv("RSSM_BWD_TWO_TILES")) ? 2 : 1;  // (a lone tile keeps the six-warp CTA: faster weight packing)
    const size_t smem = (size_t)mt::BWD_TILES * 32 * sizeof(uint2) + tpc * (size_t)fz2::BYTES;
    if (!a.rec_tiled) return cudaErrorInvalidValue;  // the fused backward reads the tile-blocked record only
    if (a.obs_projected && wpt != 3) return cudaErrorNotSupported;  // the pre-multiplied-partials mode lives in the three-warp kernel
    auto launch = [&](auto kernel) -> cudaError_t {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int groups = (tiles + tpc - 1) / tpc;  // one resident CTA per SM (shared memory): persistent over its groups
        kernel<<<groups < sms ? groups : sms, tpc == 1 ? 128 : 32 * wpt * tpc, smem, s>>>(a, u);
#ifdef FZ_TIMING
        if (getenv("RSSM_FZ_TIMING")) {
            static long long h[3][1024][12];
            cudaStreamSynchronize(s);
            cudaMemcpyFromSymbol(h, fz_dbg, sizeof(h));
            const int T = a.T < 1024 ? a.T : 1024, lo = T > 8 ? 2 : 0, hi = T > 8 ? T - 3 : T - 1;
            const char* names[3][12] = {{"top", "cpwait+END", "lp head", "hq+hp heads", "fence+E mma", "FT wait+images", "E wait+stage_hid",
                                         "h cell (pre-X)", "X SYNC wait", "l cell->Y arrive", "END mmas / Z arrive", "rest of step"},
                                        {"top", "cpwait", "pre-Y math", "refill+ELU'", "Y SYNC wait", "post-Y math", "END wait", "heads->X arrive",
                                         "mma+de gemms", "M wait", "refill+embed imgs", ""},
                                        {"top", "END wait", "embed images+loads", "FT wait+images", "X SYNC wait", "de gemms+stores", "Z SYNC wait",
                                         "MMA issue", "FT refill+d_actions", "", "", ""}};
            const char* role_name[3] = {"core", "mod", "aux"};
            for (int role = 0; role < wpt; ++role) {
                fprintf(stderr, "[fz timing B=%d T=%d] %s warp, mean cycles per interval over steps %d..%d:\n", a.B, a.T, role_name[role], lo, hi);
                const int n = role == 0 ? 12 : role == 1 ? 11 : 9;
                double tot = 0;
                for (int i = 1; i < n; ++i) {
                    double sum = 0;
                    for (int t = lo; t <= hi; ++t) sum += (double)(h[role][t][i] - h[role][t][i - 1]);
                    fprintf(stderr, "   %-20s %8.0f\n", names[role][i], sum / (hi - lo + 1));
                    tot += sum / (hi - lo + 1);
                }
                double step = 0;  // top(t-1) - top(t): steps run t = T-1 .. 0
                for (int t = lo + 1; t <= hi; ++t) step += (double)(h[role][t - 1][0] - h[role][t][0]);
                fprintf(stderr, "   %-20s %8.0f   (sum of intervals %.0f)\n", "WHOLE STEP", step / (hi - lo), tot);
            }
        }
          if (getenv("RSSM_FZ_TIMING")) {
            cudaDeviceSynchronize();
            long long gs[64];
            cudaMemcpyFromSymbol(gs, fz_grp, sizeof(gs));
            const int ng = (groups + sms - 1) / sms;
            fprintf(stderr, "[fz groups] thread 0 of CTA 0, cycles: prologue %lld", gs[1] - gs[0]);
            for (int g2 = 0; g2 < ng && g2 < 16; ++g2)
                fprintf(stderr, " | group %d: %lld = before the step loop %lld + %d steps %lld + after %lld", g2, gs[3 + 2 * g2] - gs[2 + 2 * g2],
                        gs[20 + 2 * g2] - gs[2 + 2 * g2], a.T, gs[21 + 2 * g2] - gs[20 + 2 * g2], gs[3 + 2 * g2] - gs[21 + 2 * g2]);
            fprintf(stderr, " | epilogue %lld | total %lld\n", gs[41] - gs[40], gs[41] - gs[0]);
            static long long ct[256][3];
            cudaMemcpyFromSymbol(ct, fz_cta, sizeof(ct));
            const int nc = groups < sms ? groups : sms;
            long long t0 = ct[0][0], t1 = 0;
            for (int c = 0; c < nc; ++c) t0 = ct[c][0] < t0 ? ct[c][0] : t0, t1 = ct[c][1] > t1 ? ct[c][1] : t1;
            fprintf(stderr, "[fz ctas] %d CTAs, kernel span %.1f us; per CTA (sm: start offset us, duration us):", nc, (t1 - t0) * 1e-3);
            for (int c = 0; c < nc; ++c) fprintf(stderr, " %lld:%.1f,%.1f", ct[c][2], (ct[c][0] - t0) * 1e-3, (ct[c][1] - ct[c][0]) * 1e-3);
            fprintf(stderr, "\n");
          }
#endif
        return cudaGetLastError();
    };
    const bool grouped = a.ld_feature != 0;
#define FUSED_DISPATCH(KLv, KHv) \
    if (a.KL == KLv && a.KH == KHv) {                                                                                                  \
        if (wpt == 2) return grouped ? launch(mtrssm_bwd_fused2_kernel<KLv, KHv, 2, true, false>) : launch(mtrssm_bwd_fused2_kernel<KLv, KHv, 2, false, false>); \
        if (a.obs_projected) return grouped ? launch(mtrssm_bwd_fused2_kernel<KLv, KHv, 3, true, true>) : launch(mtrssm_bwd_fused2_kernel<KLv, KHv, 3, false, true>); \
        return grouped ? launch(mtrssm_bwd_fused2_kernel<KLv, KHv, 3, true, false>) : launch(mtrssm_bwd_fused2_kernel<KLv, KHv, 3, false, false>);              \
    }
    FUSED_DISPATCH(4, 2)
#ifndef RSSM_EXP_ONLY_DEFAULT
    FUSED_DISPATCH(4, 4)
    FUSED_DISPATCH(2, 2)
    FUSED_DISPATCH(8, 8)
    FUSED_DISPATCH(16, 16)
#endif
#undef FUSED_DISPATCH
    return cudaErrorInvalidValue;
}

}  // namespace rssm
