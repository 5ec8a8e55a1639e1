// Warp-tile primitives for the RSSM/MTRSSM rollout kernels (sm_100a).
//
// Execution model: ONE WARP OWNS 16 SEQUENCES (batch rows) and walks all T steps without any
// block-level synchronisation.  Every activation of a step lives in registers in the
// mma.m16n8k16 accumulator layout ("C tile": 16 rows x 8 cols, 4 floats per lane):
//     lane = 4*g + t;   c[0],c[1] -> (row g,   cols 2t, 2t+1)
//                       c[2],c[3] -> (row g+8, cols 2t, 2t+1)
// Two adjacent C tiles re-pack, lane-locally (no shuffles, no shared memory), into one
// m16n8k16 A operand k-tile, so a chain  GEMM -> activation -> GEMM  never leaves the register
// file.
//
// LOGICAL COLUMN MAPPING.  The assignment of a layer's neurons to MMA column positions is free (the
// weight packer applies it on both the N and the K side), so it is chosen for memory coalescing: inside every
// block of 16 logical columns (= a pair of C tiles 2j, 2j+1 = one A k-tile j)
//     tile 2j  , position 2t+i  <->  logical column 16j + 4t + i        (i = 0,1)
//     tile 2j+1, position 2t+i  <->  logical column 16j + 4t + 2 + i
// i.e. lane t of a quad owns the FOUR CONSECUTIVE logical columns 16j+4t .. 16j+4t+3 of its row: every global /
// record access is one 16-byte (8-byte for bf16) access per lane per 16-block, and categorical groups of
// class size 2 or 4 are lane-local.  Weights are packed once per CTA into shared memory as ready-made B fragments
// (`pack_weight`), so the inner loop is  LDS.64 + MMA.
//
// Precision policy NS (number of bf16 splits per operand):
//   NS = 1  bf16 operands, fp32 accumulate                      -> "bf16 tensor-core path"
//   NS = 3  x = h + m + l (three bf16 terms, 24 significand bits), the six products
//           hH, hM, mH, mM, hL, lH are accumulated in fp32      -> "fp32-parity path"
//           (dropped terms are <= 2^-23 relative; matches fp32 FFMA to ~1e-6)
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

// per-(b,t) output stores.  Streaming (evict-first, __stcs) stores were measured and change nothing (forward 0.541 vs 0.540 ms).
#ifdef RSSM_EXP_STREAMING_ST
#define RSSM_ST(ptr, val) __stcs((ptr), (val))
#else
#define RSSM_ST(ptr, val) (*(ptr) = (val))
#endif

namespace rssm {

// ------------------------------------------------------------------------------------------
// bf16 packing / splitting
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);  // .x (low 16 bits) = lo
    return *reinterpret_cast<uint32_t*>(&v);
}

// splits (x0,x1) into NS bf16x2 words: out[0]=hi, out[1]=mid, out[2]=lo
template <int NS>
__device__ __forceinline__ void split_pack(float x0, float x1, uint32_t (&out)[NS]) {
    if constexpr (NS == 3) {
#ifdef RSSM_SPLIT_RN  // round-to-nearest split (first version): four float->bf16 conversions on the quarter-rate XU pipe per pair
        __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
        __nv_bfloat162 v(h0, h1);
        out[0] = *reinterpret_cast<uint32_t*>(&v);
        float r0 = x0 - __bfloat162float(h0), r1 = x1 - __bfloat162float(h1);
        __nv_bfloat16 m0 = __float2bfloat16_rn(r0), m1 = __float2bfloat16_rn(r1);
        float s0 = r0 - __bfloat162float(m0), s1 = r1 - __bfloat162float(m1);
        __nv_bfloat162 vm(m0, m1);
        out[1] = *reinterpret_cast<uint32_t*>(&vm);
        out[2] = pack_bf16(s0, s1);
#else
        // EXACT split by truncation: hi = the top 16 bits of x (8 significant bits), mid = the top 16 bits of x - hi, lo = the rest
        // (<= 8 significant bits, so it is a bf16 as it stands): x = hi + mid + lo bit for bit, and the whole split is integer
        // masks, two subtractions and three byte permutes per pair
        const uint32_t u0 = __float_as_uint(x0), u1 = __float_as_uint(x1);
        const float r0 = x0 - __uint_as_float(u0 & 0xffff0000u), r1 = x1 - __uint_as_float(u1 & 0xffff0000u);
        const uint32_t v0 = __float_as_uint(r0), v1 = __float_as_uint(r1);
        const float s0 = r0 - __uint_as_float(v0 & 0xffff0000u), s1 = r1 - __uint_as_float(v1 & 0xffff0000u);
        out[0] = __byte_perm(u0, u1, 0x7632);  // (low half, high half) = (upper 16 bits of the first, of the second value)
        out[1] = __byte_perm(v0, v1, 0x7632);
        out[2] = __byte_perm(__float_as_uint(s0), __float_as_uint(s1), 0x7632);
#endif
    } else {
        out[0] = pack_bf16(x0, x1);
    }
}

// ------------------------------------------------------------------------------------------
// MMA
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint2 b) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

// A operand: KT k-tiles of 16, NS splits.  r[s][kt][0..3] per the PTX m16n8k16 A layout:
//   [0] (row g, k 2t..2t+1) [1] (row g+8, k 2t..) [2] (row g, k 2t+8..) [3] (row g+8, k 2t+8..)
template <int NS, int KT>
struct AFrag {
    uint32_t r[NS][KT][4];
};

// two adjacent C tiles (cols 16kt .. 16kt+15) -> k-tile kt of an A operand
template <int NS, int KT>
__device__ __forceinline__ void set_ktile(AFrag<NS, KT>& a, int kt, const float (&c0)[4], const float (&c1)[4]) {
    uint32_t p[NS];
    split_pack<NS>(c0[0], c0[1], p);
#pragma unroll
    for (int s = 0; s < NS; ++s) a.r[s][kt][0] = p[s];
    split_pack<NS>(c0[2], c0[3], p);
#pragma unroll
    for (int s = 0; s < NS; ++s) a.r[s][kt][1] = p[s];
    split_pack<NS>(c1[0], c1[1], p);
#pragma unroll
    for (int s = 0; s < NS; ++s) a.r[s][kt][2] = p[s];
    split_pack<NS>(c1[2], c1[3], p);
#pragma unroll
    for (int s = 0; s < NS; ++s) a.r[s][kt][3] = p[s];
}

template <int NS, int KT>
__device__ __forceinline__ void to_afrag(AFrag<NS, KT>& a, const float (&c)[2 * KT][4]) {
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) set_ktile<NS, KT>(a, kt, c[2 * kt], c[2 * kt + 1]);
}

// acc[NT tiles] += A[16 x 16KT] * B, B = packed weight block in shared memory laid out
// [split][kt][nt / 2][lane][nt & 1] as uint2 (bfrag_slot): one 16-byte load per n-tile pair
template <int NS, int KT, int NT, bool PAIRED = true>
__device__ __forceinline__ void gemm(float (&acc)[NT][4], const AFrag<NS, KT>& a, const uint2* __restrict__ w, int lane) {
    static_assert(NT % 2 == 0, "B fragments are stored in n-tile pairs");
    if constexpr (!PAIRED) {
        static_assert(NS == 1, "the unpaired layout is used by the bf16 two-warp forward only");
#pragma unroll
        for (int kt = 0; kt < KT; ++kt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma_bf16(acc[nt], a.r[0][kt], w[(kt * NT + nt) * 32 + lane]);
        return;
    }
    constexpr int SPLIT = KT * NT * 16;  // uint4 units per split plane
    const uint4* __restrict__ w4 = reinterpret_cast<const uint4*>(w);
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
            const uint4* p = w4 + (kt * (NT / 2) + np) * 32 + lane;
            const uint4 bh = p[0];
            if constexpr (NS == 1) {
                mma_bf16(acc[2 * np], a.r[0][kt], make_uint2(bh.x, bh.y));
                mma_bf16(acc[2 * np + 1], a.r[0][kt], make_uint2(bh.z, bh.w));
            } else {
                const uint4 bm = p[SPLIT], bl = p[2 * SPLIT];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint2 xh = h == 0 ? make_uint2(bh.x, bh.y) : make_uint2(bh.z, bh.w);
                    const uint2 xm = h == 0 ? make_uint2(bm.x, bm.y) : make_uint2(bm.z, bm.w);
                    const uint2 xl = h == 0 ? make_uint2(bl.x, bl.y) : make_uint2(bl.z, bl.w);
                    mma_bf16(acc[2 * np + h], a.r[0][kt], xl);  // smallest terms first
                    mma_bf16(acc[2 * np + h], a.r[2][kt], xh);
                    mma_bf16(acc[2 * np + h], a.r[1][kt], xm);
                    mma_bf16(acc[2 * np + h], a.r[0][kt], xm);
                    mma_bf16(acc[2 * np + h], a.r[1][kt], xh);
                    mma_bf16(acc[2 * np + h], a.r[0][kt], xh);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Packs B[k][n] for k in [0,16KT), n in [0,8NT):
//   TRANS == false:  B[k][n] = W[(n0+n)*ld + k0 + k]   (y = x W^T, forward; PyTorch [out,in] weight)
//   TRANS == true :  B[k][n] = W[(n0+k)*ld + k0 + n]   (dx = dy W,  data gradient)
// with zero padding outside k < kvalid / n < nvalid.  dst layout [split][kt][nt / 2][lane][nt & 1] uint2 (bfrag_slot).
// logical column of MMA C/B column position q (0..7) of n-tile nt, and of A/B k position kp (0..15) of k-tile kt
__host__ __device__ constexpr int lcol(int nt, int q) { return 16 * (nt >> 1) + 4 * (q >> 1) + 2 * (nt & 1) + (q & 1); }
__host__ __device__ constexpr int lk(int kt, int kp) { return 16 * kt + 4 * ((kp & 7) >> 1) + 2 * (kp >> 3) + (kp & 1); }

// uint2 slot of the B fragment of (k-tile kt, n-tile nt) for `lane` inside a packed block: the fragments of an n-tile PAIR sit side
// by side, so gemm() fetches both with one 16-byte shared-memory load (half the LDS instructions in the MIO queue of kernels whose
// stall profile is short-scoreboard + MIO throttle; same bytes)
// (paired = false: one fragment per 8-byte load, [kt][nt][lane] -- lower latency to the first MMA of a short dependent chain: the
// two-warp forward uses it when a CTA runs a single tile, where pairing measured 4 % slower per step and 3 % faster at the bench batch)
__host__ __device__ constexpr int bfrag_slot(int kt, int nt, int NT, int lane, bool paired = true) {
    return paired ? ((kt * (NT >> 1) + (nt >> 1)) * 32 + lane) * 2 + (nt & 1) : (kt * NT + nt) * 32 + lane;
}

template <int NS, bool TRANS>
__device__ __forceinline__ void pack_weight(uint2* __restrict__ dst, const float* __restrict__ W, int ld, int n0, int k0,
                                            int kvalid, int nvalid, int KT, int NT, int tid, int nthreads) {
    const int total = KT * NT * 32;
    for (int idx = tid; idx < total; idx += nthreads) {
        const int lane = idx & 31, tile = idx >> 5;
        const int nt = tile % NT, kt = tile / NT;
        const int g = lane >> 2, t = lane & 3;
        const int n = lcol(nt, g);  // B fragment: column n-position = g
        float w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = lk(kt, 2 * t + (j & 1) + ((j >> 1) << 3));  // B fragment: k-positions 2t, 2t+1, 2t+8, 2t+9
            float v = 0.f;
            if (k < kvalid && n < nvalid) v = TRANS ? W[(size_t)(n0 + k) * ld + k0 + n] : W[(size_t)(n0 + n) * ld + k0 + k];
            w[j] = v;
        }
        uint32_t lo[NS], hi[NS];
        split_pack<NS>(w[0], w[1], lo);
        split_pack<NS>(w[2], w[3], hi);
#pragma unroll
        for (int s = 0; s < NS; ++s) dst[s * total + bfrag_slot(kt, nt, NT, lane)] = make_uint2(lo[s], hi[s]);
    }
}

// Batched form of pack_weight.  A kernel's ~20 weight blocks used to be packed by ~20 back-to-back pack_weight loops of one
// or two iterations each: every loop exposed a full L2 / DRAM miss latency (~1 us on a cold launch), 20-25 us of prologue per
// CTA (ncu: 3.6 % of the fused backward's samples before its first barrier + 6.7 % of warps waiting at it).  Here ONE thread
// records the blocks in a small shared-memory table (pack_add), then all threads walk the flattened item list and issue the
// global loads of PACK_BATCH items (16 scalar loads) before the first conversion, so the latency is paid once per batch.
constexpr int PACK_MAX_BLOCKS = 24, PACK_MAX_TILES = 176, PACK_BATCH = 4;
struct PackDesc {
    uint2* dst;
    const float* W;
    int ld;
    short n0, k0, kvalid, nvalid;
    unsigned char KT, NT, trans, tile0;  // tile0: first flattened tile of this block
    unsigned char paired;                // B fragments in n-tile pairs (bfrag_slot)
};
struct PackTable {
    int nblocks, ntiles;
    PackDesc d[PACK_MAX_BLOCKS];
    unsigned char tile2blk[PACK_MAX_TILES];
};

// called by ONE thread per block; same arguments as pack_weight
__device__ __forceinline__ void pack_add(PackTable& tb, bool trans, uint2* dst, const float* W, int ld, int n0, int k0, int kvalid,
                                         int nvalid, int KT, int NT, bool paired = true) {
    const int b = tb.nblocks++;
    PackDesc& d = tb.d[b];
    d.dst = dst, d.W = W, d.ld = ld, d.n0 = (short)n0, d.k0 = (short)k0, d.kvalid = (short)kvalid, d.nvalid = (short)nvalid;
    d.KT = (unsigned char)KT, d.NT = (unsigned char)NT, d.trans = trans ? 1 : 0, d.tile0 = (unsigned char)tb.ntiles, d.paired = paired ? 1 : 0;
    for (int t = 0; t < KT * NT; ++t) tb.tile2blk[tb.ntiles + t] = (unsigned char)b;
    tb.ntiles += KT * NT;
}

// all threads, after the table is complete and visible (__syncthreads)
template <int NS>
__device__ __forceinline__ void pack_run(const PackTable& tb, int tid, int nthreads) {
    const int items = tb.ntiles * 32;
    for (int base = tid; base < items; base += PACK_BATCH * nthreads) {
        float w[PACK_BATCH][4];
        uint2* out[PACK_BATCH];
        int total[PACK_BATCH];
#pragma unroll
        for (int u = 0; u < PACK_BATCH; ++u) {
            const int item = base + u * nthreads;
            out[u] = nullptr;
            total[u] = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) w[u][j] = 0.f;
            if (item < items) {
                const PackDesc d = tb.d[tb.tile2blk[item >> 5]];
                const int lane = item & 31, tile = (item >> 5) - d.tile0;
                const int nt = tile % d.NT, kt = tile / d.NT;
                const int g = lane >> 2, t = lane & 3;
                const int n = lcol(nt, g);
                total[u] = d.KT * d.NT * 32;
                out[u] = d.dst + bfrag_slot(kt, nt, d.NT, lane, d.paired != 0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = lk(kt, 2 * t + (j & 1) + ((j >> 1) << 3));
                    if (k < d.kvalid && n < d.nvalid)
                        w[u][j] = d.trans ? d.W[(size_t)(d.n0 + k) * d.ld + d.k0 + n] : d.W[(size_t)(d.n0 + n) * d.ld + d.k0 + k];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PACK_BATCH; ++u) {
            if (out[u] != nullptr) {
                uint32_t lo[NS], hi[NS];
                split_pack<NS>(w[u][0], w[u][1], lo);
                split_pack<NS>(w[u][2], w[u][3], hi);
#pragma unroll
                for (int s = 0; s < NS; ++s) out[u][s * total[u]] = make_uint2(lo[s], hi[s]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// C-tile helpers
// ------------------------------------------------------------------------------------------
// per-lane row context of a 16-row warp tile
struct Rows {
    int rA, rB;   // clamped row indices (always valid for loads)
    bool vA, vB;  // row really exists (stores)
    int g, t;     // lane = 4g + t
};

__device__ __forceinline__ Rows make_rows(int row0, int B, int lane) {
    Rows r;
    r.g = lane >> 2;
    r.t = lane & 3;
    const int a = row0 + r.g, b = a + 8;
    r.vA = a < B;
    r.vB = b < B;
    r.rA = a < B ? a : B - 1;
    r.rB = b < B ? b : B - 1;
    return r;
}

template <int NT>
__device__ __forceinline__ void init_bias(float (&acc)[NT][4], const float* __restrict__ bias, int t) {
    static_assert(NT % 2 == 0, "activations come in 16-column blocks");
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(bias + 16 * j + 4 * t);
        acc[2 * j][0] = b.x, acc[2 * j][1] = b.y, acc[2 * j][2] = b.x, acc[2 * j][3] = b.y;
        acc[2 * j + 1][0] = b.z, acc[2 * j + 1][1] = b.w, acc[2 * j + 1][2] = b.z, acc[2 * j + 1][3] = b.w;
    }
}

template <int NT>
__device__ __forceinline__ void zero_c(float (&acc)[NT][4]) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
}

// pA/pB: pointers to logical column 0 of this group of NT/2 16-column blocks in row A / row B (16-byte aligned)
template <int NT>
__device__ __forceinline__ void load_c(float (&c)[NT][4], const float* __restrict__ pA, const float* __restrict__ pB, int t) {
    static_assert(NT % 2 == 0, "activations come in 16-column blocks");
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) {
        const float4 a = *reinterpret_cast<const float4*>(pA + 16 * j + 4 * t);
        const float4 b = *reinterpret_cast<const float4*>(pB + 16 * j + 4 * t);
        c[2 * j][0] = a.x, c[2 * j][1] = a.y, c[2 * j + 1][0] = a.z, c[2 * j + 1][1] = a.w;
        c[2 * j][2] = b.x, c[2 * j][3] = b.y, c[2 * j + 1][2] = b.z, c[2 * j + 1][3] = b.w;
    }
}

template <int NT>
__device__ __forceinline__ void store_c(const float (&c)[NT][4], float* __restrict__ pA, float* __restrict__ pB, const Rows& r) {
    static_assert(NT % 2 == 0, "activations come in 16-column blocks");
#ifdef RSSM_EXP_NO_STORES  // timing experiment only: keep a data dependency, skip the traffic
    if (c[0][0] == 1.2345e-30f) *pA = c[0][1];
    return;
#endif
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) {
        if (r.vA) RSSM_ST(reinterpret_cast<float4*>(pA + 16 * j + 4 * r.t), make_float4(c[2 * j][0], c[2 * j][1], c[2 * j + 1][0], c[2 * j + 1][1]));
        if (r.vB) RSSM_ST(reinterpret_cast<float4*>(pB + 16 * j + 4 * r.t), make_float4(c[2 * j][2], c[2 * j][3], c[2 * j + 1][2], c[2 * j + 1][3]));
    }
}

// ---- per-(b,t) records exchanged between forward, backward and wgrad: fp32 on the fp32-parity path (NS = 3),
// bf16 on the bf16 path (NS = 1; the values are MMA operands there anyway, so nothing is lost for the contractions)
template <int NS>
struct Rec {
    using T = float;
};
template <>
struct Rec<1> {
    using T = __nv_bfloat16;
};

template <int NT>
__device__ __forceinline__ void store_rec(const float (&c)[NT][4], float* __restrict__ pA, float* __restrict__ pB, const Rows& r) {
    store_c<NT>(c, pA, pB, r);
}
template <int NT>
__device__ __forceinline__ void store_rec(const float (&c)[NT][4], __nv_bfloat16* __restrict__ pA, __nv_bfloat16* __restrict__ pB,
                                          const Rows& r) {
#if defined(RSSM_EXP_NO_STORES) || defined(RSSM_EXP_NO_REC_STORES)
    if (c[0][0] == 1.2345e-30f) *pA = __float2bfloat16(c[0][1]);
    return;
#endif
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) {
        if (r.vA)
            RSSM_ST(reinterpret_cast<uint2*>(pA + 16 * j + 4 * r.t),
                    make_uint2(pack_bf16(c[2 * j][0], c[2 * j][1]), pack_bf16(c[2 * j + 1][0], c[2 * j + 1][1])));
        if (r.vB)
            RSSM_ST(reinterpret_cast<uint2*>(pB + 16 * j + 4 * r.t),
                    make_uint2(pack_bf16(c[2 * j][2], c[2 * j][3]), pack_bf16(c[2 * j + 1][2], c[2 * j + 1][3])));
    }
}
template <int NT>
__device__ __forceinline__ void load_rec(float (&c)[NT][4], const float* __restrict__ pA, const float* __restrict__ pB, int t) {
    load_c<NT>(c, pA, pB, t);
}
template <int NT>
__device__ __forceinline__ void load_rec(float (&c)[NT][4], const __nv_bfloat16* __restrict__ pA, const __nv_bfloat16* __restrict__ pB,
                                         int t) {
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) {
        const uint2 a = *reinterpret_cast<const uint2*>(pA + 16 * j + 4 * t);
        const uint2 b = *reinterpret_cast<const uint2*>(pB + 16 * j + 4 * t);
        const __nv_bfloat162 a0 = *reinterpret_cast<const __nv_bfloat162*>(&a.x), a1 = *reinterpret_cast<const __nv_bfloat162*>(&a.y);
        const __nv_bfloat162 b0 = *reinterpret_cast<const __nv_bfloat162*>(&b.x), b1 = *reinterpret_cast<const __nv_bfloat162*>(&b.y);
        c[2 * j][0] = __low2float(a0), c[2 * j][1] = __high2float(a0), c[2 * j + 1][0] = __low2float(a1), c[2 * j + 1][1] = __high2float(a1);
        c[2 * j][2] = __low2float(b0), c[2 * j][3] = __high2float(b0), c[2 * j + 1][2] = __low2float(b1), c[2 * j + 1][3] = __high2float(b1);
    }
}

// bf16 A operand (KT k-tiles = 16*KT logical columns) -> bf16 record columns: lane (g,t) owns the four consecutive columns
// 16kt + 4t .. + 3 of rows g (registers 0, 2) and g + 8 (registers 1, 3)
template <int KT>
__device__ __forceinline__ void store_afrag(const AFrag<1, KT>& a, __nv_bfloat16* __restrict__ pA, __nv_bfloat16* __restrict__ pB,
                                            const Rows& r) {
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
        if (r.vA) *reinterpret_cast<uint2*>(pA + 16 * kt + 4 * r.t) = make_uint2(a.r[0][kt][0], a.r[0][kt][2]);
        if (r.vB) *reinterpret_cast<uint2*>(pB + 16 * kt + 4 * r.t) = make_uint2(a.r[0][kt][1], a.r[0][kt][3]);
    }
}

// L2 prefetch (next step's rows): converts the DRAM round trip of a step's dependent loads into L2 hits.
__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifndef RSSM_EXP_NO_PREFETCH
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
// whole range [p, p + bytes): p 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void prefetch_bulk_l2(const void* p, uint32_t bytes) {
#ifndef RSSM_EXP_NO_PREFETCH
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#endif
}

// store the logical columns < nvalid (nvalid even, <= 16) of ONE 16-column block, e.g. the action gradient;
// rows need only be 8-byte aligned
__device__ __forceinline__ void store_c_partial(const float (&c)[2][4], float* __restrict__ pA, float* __restrict__ pB, const Rows& r,
                                                int nvalid) {
    const int c0 = 4 * r.t;
    if (c0 + 1 < nvalid) {
        if (r.vA) *reinterpret_cast<float2*>(pA + c0) = make_float2(c[0][0], c[0][1]);
        if (r.vB) *reinterpret_cast<float2*>(pB + c0) = make_float2(c[0][2], c[0][3]);
    }
    if (c0 + 3 < nvalid) {
        if (r.vA) *reinterpret_cast<float2*>(pA + c0 + 2) = make_float2(c[1][0], c[1][1]);
        if (r.vB) *reinterpret_cast<float2*>(pB + c0 + 2) = make_float2(c[1][2], c[1][3]);
    }
}

// A operand straight from global fp32 rows (8-byte aligned); logical columns >= kvalid (even) read as zero
template <int NS, int KT>
__device__ __forceinline__ void load_a_global(AFrag<NS, KT>& a, const float* __restrict__ pA, const float* __restrict__ pB, int t,
                                              int kvalid) {
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
        float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
        const int k0 = kt * 16 + 4 * t;
        if (k0 + 1 < kvalid) {
            const float2 x = *reinterpret_cast<const float2*>(pA + k0), y = *reinterpret_cast<const float2*>(pB + k0);
            c0[0] = x.x, c0[1] = x.y, c0[2] = y.x, c0[3] = y.y;
        }
        if (k0 + 3 < kvalid) {
            const float2 x = *reinterpret_cast<const float2*>(pA + k0 + 2), y = *reinterpret_cast<const float2*>(pB + k0 + 2);
            c1[0] = x.x, c1[1] = x.y, c1[2] = y.x, c1[3] = y.y;
        }
        set_ktile<NS, KT>(a, kt, c0, c1);
    }
}

// ------------------------------------------------------------------------------------------
// per-warp input staging (forward kernels): the next step's inputs of the warp's 16 rows are pulled into shared
// memory with cp.async while the current step computes, so the step never waits on a global load.
// Stage layout (floats): EA[16][64] | EV[16][64] | ACT[16][8] | U0[16][8] | U1[16][8].  The two embedding tiles use a
// 16-byte-chunk XOR swizzle (chunk ^ 4*(row & 1)) so that the mma A-fragment read pattern (lane (g,t) reads the 16 bytes
// at logical column 16kt + 4t of row g) is bank-conflict free without padding.
// ------------------------------------------------------------------------------------------
namespace stg {
constexpr int EA = 0, EV = 1024, ACT = 2048, U0 = 2176, U1 = 2304, FLOATS = 2432;  // 9728 bytes per stage
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N>  // wait until at most N of the most recently committed groups are still pending
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Issue the copies of step `t` for the warp tile starting at row0 into `stage`.
//   ea/ev: [B,T,64] fp32;  act: [B,T,A] fp32 (A even, <= 8);  u0/u1: [B,T,n0] / [B,T,n1] fp32 with n0, n1 in {1,2,4,8}
//   (rows of 4*n bytes; n = 1,2 use 4/8-byte copies).  Any of ea/ev/u0/u1 may be null (skipped).
__device__ __forceinline__ void stage_inputs(float* stage, const float* __restrict__ ea, const float* __restrict__ ev,
                                             const float* __restrict__ act, int A, const float* __restrict__ u0, int n0,
                                             const float* __restrict__ u1, int n1, int row0, int B, int T, int t, int lane) {
    // embeddings: 16 rows x 16 chunks each -> lane owns chunk (lane & 15) of rows (lane >> 4) + 2i
    const int chunk = lane & 15;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rl = (lane >> 4) + 2 * i;
        const int row = min(row0 + rl, B - 1);
        const size_t src = ((size_t)row * T + t) * 64 + chunk * 4;
        const int dst = rl * 64 + ((chunk ^ (4 * (rl & 1))) << 2);
        if (ea != nullptr) cp_async16(stage + stg::EA + dst, ea + src);
        if (ev != nullptr) cp_async16(stage + stg::EV + dst, ev + src);
    }
    // small rows: lane -> (row = lane & 15, half = lane >> 4)
    {
        const int rl = lane & 15, half = lane >> 4;
        const int row = min(row0 + rl, B - 1);
        const size_t idx = (size_t)row * T + t;
        // actions: A/2 pieces of 8 bytes; half 0 copies pieces 0,1 and half 1 pieces 2,3
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int piece = 2 * half + k;
            if (act != nullptr && 2 * piece < A) cp_async8(stage + stg::ACT + rl * 8 + 2 * piece, act + idx * A + 2 * piece);
        }
        const float* us[2] = {u0, u1};
        const int ns[2] = {n0, n1};
        const int offs[2] = {stg::U0, stg::U1};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (us[k] == nullptr) continue;
            const int n = ns[k];
            float* d = stage + offs[k] + rl * 8;
            const float* sp = us[k] + idx * n;
            if (n == 8) {
                cp_async16(d + 4 * half, sp + 4 * half);
            } else if (n == 4) {
                if (half == 0) cp_async16(d, sp);
            } else if (n == 2) {
                if (half == 0) cp_async8(d, sp);
            } else if (half == 0) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(d)), "l"(sp) : "memory");
            }
        }
    }
    cp_async_commit();
}

// ---- mbarrier + bulk-copy (TMA 1-D) helpers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// asynchronous bulk store shared -> global (TMA 1-D); completion tracked per issuing thread in bulk groups
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// until the sources of all of this thread's bulk stores have been read (the shared-memory buffer may be rewritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// bounded wait: a lost transaction traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    // unroll 1: the compiler otherwise unrolls the retry 64 times -- 130 dead instructions at EVERY wait site, in the middle of the
    // hot loops (1.3 k of the fused backward's 7.7 k instructions; instruction fetch was 12-15 % of its stall samples)
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}

// A operand (KT = 4, 64 columns) from a swizzled staged embedding tile
template <int NS>
__device__ __forceinline__ void load_a_staged64(AFrag<NS, 4>& a, const float* tile, int g, int t) {
    const int sw = 4 * (g & 1);  // same for row g and row g + 8
    const float* pA = tile + g * 64;
    const float* pB = tile + (g + 8) * 64;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
        const int c = ((4 * kt + t) ^ sw) << 2;
        const float4 x = *reinterpret_cast<const float4*>(pA + c), y = *reinterpret_cast<const float4*>(pB + c);
        const float q0[4] = {x.x, x.y, y.x, y.y}, q1[4] = {x.z, x.w, y.z, y.w};
        set_ktile<NS, 4>(a, kt, q0, q1);
    }
}

// A operand (KT = 1) from the staged action rows ACT[16][8] (columns >= A are zero)
template <int NS>
__device__ __forceinline__ void load_a_staged_act(AFrag<NS, 1>& a, const float* act, int g, int t) {
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
    if (t < 2) {
        x = *reinterpret_cast<const float4*>(act + g * 8 + 4 * t);
        y = *reinterpret_cast<const float4*>(act + (g + 8) * 8 + 4 * t);
    }
    const float q0[4] = {x.x, x.y, y.x, y.y}, q1[4] = {x.z, x.w, y.z, y.w};
    set_ktile<NS, 1>(a, 0, q0, q1);
}

// ------------------------------------------------------------------------------------------
// elementwise math (fp32; accurate libm variants -- parity first)
// ------------------------------------------------------------------------------------------
// FAST = false: accurate libm (fp32-parity path).  FAST = true: MUFU approximations (ex2/lg2/tanh/rcp.approx,
// ~2^-11 relative) -- used by the bf16 tensor-core path, whose operands are already rounded to 8 bits.
template <bool FAST>
struct Math {
#ifdef RSSM_EXP_NO_MUFU  // timing experiment only: replace transcendental math by one FMA
    static __device__ __forceinline__ float exp(float x) { return 1.f + x * 0.5f; }
    static __device__ __forceinline__ float log(float x) { return x - 1.f; }
    static __device__ __forceinline__ float div(float a, float b) { return a * (2.f - b); }
    static __device__ __forceinline__ float tanh(float x) { return x * 0.5f; }
    static __device__ __forceinline__ float sigmoid(float x) { return 0.5f + 0.25f * x; }
    static __device__ __forceinline__ float elu(float x) { return x > 0.f ? x : 0.5f * x; }
};
template <bool FAST>
struct MathUnused {
#endif
    // FAST: single MUFU ops with flush-to-zero (no denormal guard code around them).
    // !FAST (fp32-parity policy): fp32-accurate to a few ulp, but SHORT.  The libm calls (expf 25, logf 30, tanhf 35, expm1f 30, IEEE
    // division 10 instructions, each inlined ~250 times per step) made the one-warp kernels 11 k instructions = 176 KB of code
    // per step against a 32 KB instruction cache: ncu showed 55 % of the fp32 forward's stall cycles as instruction fetch.  The
    // MUFU units are accurate to 2^-22 relative (ex2, lg2) / 1 ulp (rcp); what libm adds is argument handling, done here by a
    // compensated product (exp), a Newton step (division) and short series near zero (tanh, expm1).  -DRSSM_LIBM restores libm.
    static __device__ __forceinline__ float mufu_ex2(float x) {
        float y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    static __device__ __forceinline__ float exp(float x) {
        if constexpr (FAST) {
            return mufu_ex2(x * 1.4426950408889634f);
        } else {
#ifdef RSSM_LIBM
            return expf(x);
#else
            // x log2(e) in two parts: the rounding error of the leading product is recovered by an FMA, the tail of the constant added
            const float t = x * 1.4426950216293335f;
            const float e = fmaf(x, 1.4426950216293335f, -t) + x * 1.9259629911266175e-8f;
            const float y = mufu_ex2(t);
            return fmaf(y, e * 0.6931471805599453f, y);  // 2^e = 1 + e ln 2 + O(e^2), |e| < 2^-22 |t|
#endif
        }
    }
    static __device__ __forceinline__ float log(float x) {
        if constexpr (FAST) {
            float y;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
            return y * 0.6931471805599453f;
        } else {
#ifdef RSSM_LIBM
            return logf(x);
#else
            float y;
            asm("lg2.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
            return y * 0.6931471805599453f;
#endif
        }
    }
    static __device__ __forceinline__ float div(float a, float b) {
        if constexpr (FAST) {
            float y;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(b));
            return a * y;
        } else {
#ifdef RSSM_LIBM
            return a / b;
#else
            float r;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
            r = fmaf(r, fmaf(-b, r, 1.f), r);  // one Newton step: < 1 ulp (the operands here are sums of exponentials, never denormal)
            return a * r;
#endif
        }
    }
    static __device__ __forceinline__ float tanh(float x) {
        if constexpr (FAST) {
            float y;
            asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
            return y;
        } else {
#ifdef RSSM_LIBM
            return tanhf(x);
#else
            const float ax = fabsf(x), x2 = x * x;
            const float e = exp(-2.f * ax);
            const float big = div(1.f - e, 1.f + e);                                       // cancels for small |x| ...
            const float small = ax * fmaf(x2, fmaf(x2, 0.13333333f, -0.33333334f), 1.f);   // ... where the series is exact to 5e-8 relative
            return copysignf(ax < 0.1f ? small : big, x);
#endif
        }
    }
    static __device__ __forceinline__ float sigmoid(float x) { return div(1.f, 1.f + exp(-x)); }
    static __device__ __forceinline__ float elu(float x) {
        if constexpr (FAST) {
            return x > 0.f ? x : exp(x) - 1.f;
        } else {
#ifdef RSSM_LIBM
            return x > 0.f ? x : expm1f(x);
#else
            // exp(x) - 1 cancels near zero: Taylor series to x^5 there (1.4e-9 absolute at |x| = 0.1)
            const float series = x * fmaf(x, fmaf(x, fmaf(x, fmaf(x, 8.3333333e-3f, 4.1666668e-2f), 0.16666667f), 0.5f), 1.f);
            const float m = x > -0.1f ? series : exp(x) - 1.f;
            return x > 0.f ? x : m;
#endif
        }
    }
};
// d ELU / d pre, from the POST-activation value y: y > 0 -> 1, else exp(x) = y + 1
__device__ __forceinline__ float elu_grad_from_out(float y) { return fminf(y, 0.f) + 1.f; }

template <bool FAST>
struct EluOp {
    __device__ __forceinline__ float operator()(float x) const { return Math<FAST>::elu(x); }
};

template <int NT, typename F>
__device__ __forceinline__ void map_c(float (&c)[NT][4], F f) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[nt][j] = f(c[nt][j]);
}

// ------------------------------------------------------------------------------------------
// categorical helpers on a 16-wide logit vector = 2 C tiles.  Group = K consecutive columns.
// Row A uses c[nt][0..1], row B uses c[nt][2..3]; `h` = 0 (row A) or 1 (row B).
// ------------------------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;

// reduce a per-(tile,row) in-lane value over the K-column group it belongs to; result in v[0], v[1].
// A lane owns logical columns 4t..4t+3 of the 16: tile 0 holds (4t, 4t+1), tile 1 holds (4t+2, 4t+3).
template <int K, bool MAX>
__device__ __forceinline__ void group_reduce(float (&v)[2]) {
    static_assert(K == 2 || K == 4 || K == 8 || K == 16, "class_size must be 2, 4, 8 or 16");
    if constexpr (K >= 4) {
        const float m = MAX ? fmaxf(v[0], v[1]) : v[0] + v[1];
        v[0] = v[1] = m;
    }
    if constexpr (K >= 8) {
        const float o = __shfl_xor_sync(FULL, v[0], 1);
        v[0] = v[1] = MAX ? fmaxf(v[0], o) : v[0] + o;
    }
    if constexpr (K == 16) {
        const float o = __shfl_xor_sync(FULL, v[0], 2);
        v[0] = v[1] = MAX ? fmaxf(v[0], o) : v[0] + o;
    }
}

// per-group softmax (MultiOneHotFactory, A1): p = exp(x - max_g) / sum_g
template <int K, bool FAST>
__device__ __forceinline__ void softmax_groups(const float (&x)[2][4], float (&p)[2][4]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float m[2] = {fmaxf(x[0][2 * h], x[0][2 * h + 1]), fmaxf(x[1][2 * h], x[1][2 * h + 1])};
        group_reduce<K, true>(m);
        float e[2][2], s[2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            e[nt][0] = Math<FAST>::exp(x[nt][2 * h] - m[nt]);
            e[nt][1] = Math<FAST>::exp(x[nt][2 * h + 1] - m[nt]);
            s[nt] = e[nt][0] + e[nt][1];
        }
        group_reduce<K, false>(s);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            p[nt][2 * h] = Math<FAST>::div(e[nt][0], s[nt]);
            p[nt][2 * h + 1] = Math<FAST>::div(e[nt][1], s[nt]);
        }
    }
}

// log_softmax over the FLAT 16 columns (F.log_softmax(dim=-1)); optionally also the softmax
template <bool FAST>
__device__ __forceinline__ void log_softmax_flat(const float (&x)[2][4], float (&ls)[2][4]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float m[2] = {fmaxf(x[0][2 * h], x[0][2 * h + 1]), fmaxf(x[1][2 * h], x[1][2 * h + 1])};
        group_reduce<16, true>(m);
        float s[2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) s[nt] = Math<FAST>::exp(x[nt][2 * h] - m[0]) + Math<FAST>::exp(x[nt][2 * h + 1] - m[0]);
        group_reduce<16, false>(s);
        const float lse = m[0] + Math<FAST>::log(s[0]);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            ls[nt][2 * h] = x[nt][2 * h] - lse;
            ls[nt][2 * h + 1] = x[nt][2 * h + 1] - lse;
        }
    }
}

// MoPoE fusion (mopoe_mrssm/core.py:241-251,135-154): mixed = logsumexp([la, lv, la+lv] + log(1/3)) on
// flat log-softmaxes.  Also returns the per-expert responsibilities needed by the backward:
//   ra = (e^la + e^{la+lv}) / (e^la + e^lv + e^{la+lv}),  rv likewise.
template <bool FAST>
__device__ __forceinline__ void mopoe_mix(const float (&ls_a)[2][4], const float (&ls_v)[2][4], float (&mixed)[2][4],
                                          float (*ra)[4], float (*rv)[4]) {
    const float LOG_THIRD = -1.0986122886681098f;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = ls_a[nt][j], v = ls_v[nt][j], f = a + v;
            const float mx = fmaxf(a, fmaxf(v, f));
            const float ea = Math<FAST>::exp(a - mx), ev = Math<FAST>::exp(v - mx), ef = Math<FAST>::exp(f - mx);
            const float s = ea + ev + ef;
            mixed[nt][j] = LOG_THIRD + mx + Math<FAST>::log(s);
            if (ra != nullptr) {
                ra[nt][j] = Math<FAST>::div(ea + ef, s);
                rv[nt][j] = Math<FAST>::div(ev + ef, s);
            }
        }
}

// ---- probability-domain MoPoE (bf16 policies) -----------------------------------------------------------------------------
// With pa = softmax_16(la), pv = softmax_16(lv):  exp(mixed) = (pa + pv + pa pv) / 3, so the per-group softmax of `mixed`
// (MultiOneHotFactory on the fused logits, mopoe_mrssm/core.py:241-251) is  q = s / sum_group(s),  s = pa + pv + pa pv,
// and the responsibilities are ra = (pa + pa pv) / s, rv = (pv + pa pv) / s: no logarithm and one exponential per logit
// instead of four.  If a whole group underflows (flat log-probabilities below ~ -69 in BOTH modalities) the callers fall
// back to the log-domain functions above (warp-uniform branch, never taken at sane logit scales).
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// softmax over the FLAT 16 columns
__device__ __forceinline__ void softmax_flat_fast(const float (&x)[2][4], float (&p)[2][4]) {
    constexpr float L2E = 1.4426950408889634f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float m[2] = {fmaxf(x[0][2 * h], x[0][2 * h + 1]), fmaxf(x[1][2 * h], x[1][2 * h + 1])};
        group_reduce<16, true>(m);
        const float nm = -m[0] * L2E;
        float e[2][2], s[2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            e[nt][0] = ex2_ftz(fmaf(x[nt][2 * h], L2E, nm));
            e[nt][1] = ex2_ftz(fmaf(x[nt][2 * h + 1], L2E, nm));
            s[nt] = e[nt][0] + e[nt][1];
        }
        group_reduce<16, false>(s);
        const float inv = rcp_ftz(s[0]);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) p[nt][2 * h] = e[nt][0] * inv, p[nt][2 * h + 1] = e[nt][1] * inv;
    }
}
constexpr float MOPOE_TINY = 1e-30f;
// fused posterior probabilities q (groups of K) from the two modality logits
template <int K>
__device__ __forceinline__ void mopoe_posterior_fast(const float (&la)[2][4], const float (&lv)[2][4], float (&q)[2][4]) {
    float pa[2][4], pv[2][4];
    softmax_flat_fast(la, pa);
    softmax_flat_fast(lv, pv);
    float lo = 1.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float s[2][2], g[2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            s[nt][0] = fmaf(pa[nt][2 * h], pv[nt][2 * h], pa[nt][2 * h] + pv[nt][2 * h]);
            s[nt][1] = fmaf(pa[nt][2 * h + 1], pv[nt][2 * h + 1], pa[nt][2 * h + 1] + pv[nt][2 * h + 1]);
            g[nt] = s[nt][0] + s[nt][1];
        }
        group_reduce<K, false>(g);
        lo = fminf(lo, fminf(g[0], g[1]));
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const float inv = rcp_ftz(g[nt]);
            q[nt][2 * h] = s[nt][0] * inv, q[nt][2 * h + 1] = s[nt][1] * inv;
        }
    }
    if (__any_sync(FULL, !(lo >= MOPOE_TINY))) {  // a whole group underflowed: log-domain evaluation
        float lsa[2][4], lsv[2][4], mixed[2][4];
        log_softmax_flat<true>(la, lsa);
        log_softmax_flat<true>(lv, lsv);
        mopoe_mix<true>(lsa, lsv, mixed, nullptr, nullptr);
        softmax_groups<K, true>(mixed, q);
    }
}
// backward side: softmax(la), softmax(lv) (all the flat log-softmax backward needs) and the two responsibilities
__device__ __forceinline__ void mopoe_responsibilities_fast(const float (&la)[2][4], const float (&lv)[2][4], float (&pa)[2][4],
                                                            float (&pv)[2][4], float (&ra)[2][4], float (&rv)[2][4]) {
    softmax_flat_fast(la, pa);
    softmax_flat_fast(lv, pv);
    float lo = 1.f;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float pf = pa[nt][j] * pv[nt][j], s = pa[nt][j] + pv[nt][j] + pf;
            const float inv = rcp_ftz(s);
            lo = fminf(lo, s);
            ra[nt][j] = (pa[nt][j] + pf) * inv;
            rv[nt][j] = (pv[nt][j] + pf) * inv;
        }
    if (__any_sync(FULL, !(lo >= MOPOE_TINY))) {
        float lsa[2][4], lsv[2][4], mixed[2][4];
        log_softmax_flat<true>(la, lsa);
        log_softmax_flat<true>(lv, lsv);
        mopoe_mix<true>(lsa, lsv, mixed, ra, rv);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) pa[nt][j] = Math<true>::exp(lsa[nt][j]), pv[nt][j] = Math<true>::exp(lsv[nt][j]);
    }
}

// Inverse-CDF categorical draw (A2): idx = min(K-1, #{k : cdf_k <= u}), cdf accumulated in class order.
// p: per-group probabilities (2 tiles); uA/uB: pointers to this row's uniforms u[c], c = 0..16/K-1.
// Writes the exact one-hot into z (C-tile layout).
template <int K>
__device__ __forceinline__ void sample_onehot(const float (&p)[2][4], const float* __restrict__ uA, const float* __restrict__ uB,
                                              float (&z)[2][4], int lane) {
    const int t = lane & 3, qbase = lane & ~3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float* u = h == 0 ? uA : uB;
        const float mine[4] = {p[0][2 * h], p[0][2 * h + 1], p[1][2 * h], p[1][2 * h + 1]};  // logical columns 4t .. 4t+3
        int hit[4];  // hit[i]: my i-th column is the drawn class of its group
        if constexpr (K == 2) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int idx = mine[2 * half] <= u[2 * t + half] ? 1 : 0;
                hit[2 * half] = idx == 0, hit[2 * half + 1] = idx == 1;
            }
        } else {
            constexpr int LANES = K / 4;  // quad lanes per group
            const int first = t & ~(LANES - 1), pos = 4 * (t & (LANES - 1));
            const float uu = u[t / LANES];
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

// backward of the per-group softmax: dx = p * (dp - sum_g p dp)
template <int K>
__device__ __forceinline__ void softmax_groups_bwd(const float (&p)[2][4], const float (&dp)[2][4], float (&dx)[2][4]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float s[2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) s[nt] = p[nt][2 * h] * dp[nt][2 * h] + p[nt][2 * h + 1] * dp[nt][2 * h + 1];
        group_reduce<K, false>(s);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            dx[nt][2 * h] = p[nt][2 * h] * (dp[nt][2 * h] - s[nt]);
            dx[nt][2 * h + 1] = p[nt][2 * h + 1] * (dp[nt][2 * h + 1] - s[nt]);
        }
    }
}

// backward of the flat log_softmax: dx = dls - softmax(x) * sum(dls), softmax(x) = exp(ls)
template <bool FAST>
__device__ __forceinline__ void log_softmax_flat_bwd(const float (&ls)[2][4], const float (&dls)[2][4], float (&dx)[2][4]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float s[2] = {dls[0][2 * h] + dls[0][2 * h + 1], dls[1][2 * h] + dls[1][2 * h + 1]};
        group_reduce<16, false>(s);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            dx[nt][2 * h] = dls[nt][2 * h] - Math<FAST>::exp(ls[nt][2 * h]) * s[0];
            dx[nt][2 * h + 1] = dls[nt][2 * h + 1] - Math<FAST>::exp(ls[nt][2 * h + 1]) * s[0];
        }
    }
}

// KL(q || p) summed over all 16 columns of a row (torch.distributions OneHotCategorical KL with its
// eps clamp on the logs); result valid in every lane of the quad: out[0] row A, out[1] row B
template <bool FAST>
__device__ __forceinline__ float clamp_log(float p) {
    const float eps = 1.1920928955078125e-07f;
    return Math<FAST>::log(fminf(fmaxf(p, eps), 1.f - eps));
}
template <bool FAST>
__device__ __forceinline__ void kl_rows(const float (&q)[2][4], const float (&p)[2][4], float (&out)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float s[2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            s[nt] = q[nt][2 * h] * (clamp_log<FAST>(q[nt][2 * h]) - clamp_log<FAST>(p[nt][2 * h])) +
                    q[nt][2 * h + 1] * (clamp_log<FAST>(q[nt][2 * h + 1]) - clamp_log<FAST>(p[nt][2 * h + 1]));
        }
        group_reduce<16, false>(s);
        out[h] = s[0];
    }
}

// d KL / d q (weight wq) and d KL / d p (weight wp), added into dq / dp.  dkl[h] = upstream grad of the
// row's KL.  dKL/dq_k = log q_k - log p_k + 1 ; dKL/dp_k = -q_k / p_k.
template <bool FAST>
__device__ __forceinline__ void kl_rows_bwd(const float (&q)[2][4], const float (&p)[2][4], const float (&dkl)[2], float wq, float wp,
                                            float (&dq)[2][4], float (&dp)[2][4]) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float g = dkl[j >> 1];
            if constexpr (FAST) {
                // one reciprocal serves both halves: log q - log p = log(q / p) (the upper clamp of p, 1 - eps, moves the ratio by
                // at most 1.2e-7 relative)
                const float eps = 1.1920928955078125e-07f;
                const float rp = rcp_ftz(fmaxf(p[nt][j], eps));
                const float qc = fminf(fmaxf(q[nt][j], eps), 1.f - eps);
                dq[nt][j] += g * wq * (Math<true>::log(qc * rp) + 1.f);
                dp[nt][j] -= g * wp * (q[nt][j] * rp);
            } else {
                dq[nt][j] += g * wq * (clamp_log<FAST>(q[nt][j]) - clamp_log<FAST>(p[nt][j]) + 1.f);
                dp[nt][j] -= g * wp * Math<FAST>::div(q[nt][j], fmaxf(p[nt][j], 1.1920928955078125e-07f));
            }
        }
}

}  // namespace rssm
