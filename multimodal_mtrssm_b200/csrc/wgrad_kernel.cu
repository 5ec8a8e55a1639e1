// Batched weight gradients of the rollout on the tensor cores.
//
// For every layer:  dW[n][k] += sum_{(b,t)} dY[(b,t)][n] * X[(b,t)][k],   db[n] += sum_{(b,t)} dY[(b,t)][n]
// with dY a slice of the backward kernel's per-(b,t) "dpre" record and X the layer's input (outputs of the
// forward kernel, its saved record, or kernel inputs; "previous-step" inputs are the row (b,t-1), or the
// initial state for t == 0).  The recurrence only carries data gradients, so this pass is embarrassingly
// parallel over (b,t).
//
// Design: a CTA stages ROWS = 32 (16 on the fp32 path) consecutive (b,t) rows ONCE from HBM (cp.async) into shared memory
// as bf16 (NS planes: 1 = bf16 path, 3 = hi/mid/lo split for the fp32-parity path), laid out per layer input so
// that every layer is one contiguous column range.  Eight warps then each own a fixed set of 16x8 output tiles
// (<= 20, register accumulators that live across the CTA's whole row loop) and compute
//     acc[n-tile][k-tile] += dY^T[16 n x 16 rows] * X[16 rows x 8 k]        (mma.m16n8k16, K = rows)
// with both operands fetched by ldmatrix.trans.  Bias gradients ride along as one extra MMA per n-tile
// against a constant "ones" B fragment.  At the end every CTA adds its partial sums to global memory
// (one atomicAdd per output element per CTA).  HBM traffic = each staged column read exactly once.
#include <stdint.h>

#include "frag.cuh"
#include "kernels.h"

namespace rssm {

namespace wg {

__host__ __device__ constexpr int rows_of(int ns) { return ns == 1 ? 32 : 16; }  // (b,t) rows staged per block
constexpr int MAX_TILES = 20;
constexpr int THREADS = 256;

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], const void* smem_ptr) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr));
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}

// One part: MT n-tiles of dY (16 columns each, starting at staged column colY) times NT k-tiles of X (8 columns
// each, starting at staged column colX), accumulators acc[OFF .. OFF + MT*NT), bias accumulators (if BIAS) at
// acc[OFF + MT*NT .. + MT).
template <int NS, int MT, int NT, int OFF, bool BIAS>
__device__ __forceinline__ void part(float (&acc)[MAX_TILES][4], const __nv_bfloat16* __restrict__ smY, int strideY,
                                     const __nv_bfloat16* __restrict__ smX, int strideX, int plane, int colY, int colX, int lane) {
    constexpr int ROWS = rows_of(NS);
    static_assert(OFF + MT * NT + (BIAS ? MT : 0) <= MAX_TILES, "too many tiles for one warp");
    const int lr = lane & 7, m8 = (lane >> 3) & 1, m16 = (lane >> 4) & 1;
    const uint32_t one2 = (lane >> 2) == 0 ? 0x3f803f80u : 0u;  // bf16 (1,1) in lanes holding output column 0
#pragma unroll
    for (int ks = 0; ks < ROWS / 16; ++ks) {
        const int r0 = ks * 16;
        uint32_t a[NS][MT][4];
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
                ldmatrix_x4_trans(a[s][mt], smY + (size_t)s * plane + (size_t)(r0 + lr + m16 * 8) * strideY + colY + 16 * mt + m8 * 8);
        if constexpr (BIAS) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                mma_bf16(acc[OFF + MT * NT + mt], a[0][mt], make_uint2(one2, one2));
                if constexpr (NS == 3) {
                    mma_bf16(acc[OFF + MT * NT + mt], a[1][mt], make_uint2(one2, one2));
                    mma_bf16(acc[OFF + MT * NT + mt], a[2][mt], make_uint2(one2, one2));
                }
            }
        }
#pragma unroll
        for (int np = 0; np < (NT + 1) / 2; ++np) {
            const bool pair = 2 * np + 1 < NT;
            uint32_t b[NS][4];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const __nv_bfloat16* base = smX + (size_t)s * plane + (size_t)(r0 + lr + m8 * 8) * strideX + colX + 16 * np;
                if (pair) {
                    ldmatrix_x4_trans(b[s], base + m16 * 8);
                } else {
                    uint32_t t2[2];
                    ldmatrix_x2_trans(t2, base);
                    b[s][0] = t2[0], b[s][1] = t2[1], b[s][2] = 0u, b[s][3] = 0u;
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h == 1 && !pair) break;
                const int nt = 2 * np + h;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    float (&c)[4] = acc[OFF + mt * NT + nt];
                    if constexpr (NS == 1) {
                        mma_bf16(c, a[0][mt], make_uint2(b[0][2 * h], b[0][2 * h + 1]));
                    } else {
                        mma_bf16(c, a[0][mt], make_uint2(b[2][2 * h], b[2][2 * h + 1]));
                        mma_bf16(c, a[2][mt], make_uint2(b[0][2 * h], b[0][2 * h + 1]));
                        mma_bf16(c, a[1][mt], make_uint2(b[1][2 * h], b[1][2 * h + 1]));
                        mma_bf16(c, a[0][mt], make_uint2(b[1][2 * h], b[1][2 * h + 1]));
                        mma_bf16(c, a[1][mt], make_uint2(b[0][2 * h], b[0][2 * h + 1]));
                        mma_bf16(c, a[0][mt], make_uint2(b[0][2 * h], b[0][2 * h + 1]));
                    }
                }
            }
        }
    }
}

// adds a part's accumulators to global memory
template <int MT, int NT, int OFF, bool BIAS>
__device__ __forceinline__ void flush(const float (&acc)[MAX_TILES][4], const WgradOut& o, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const float (&c)[4] = acc[OFF + mt * NT + nt];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = 16 * mt + g + (j >> 1) * 8, k = 8 * nt + 2 * t + (j & 1);
                if (k < o.kvalid) atomicAdd(o.dW + (size_t)n * o.ldw + k, c[j]);
            }
        }
        if constexpr (BIAS) {
            if (t == 0) {
                const float (&c)[4] = acc[OFF + MT * NT + mt];
                if (o.db0) atomicAdd(o.db0 + 16 * mt + g, c[0]), atomicAdd(o.db0 + 16 * mt + g + 8, c[2]);
                if (o.db1) atomicAdd(o.db1 + 16 * mt + g, c[0]), atomicAdd(o.db1 + 16 * mt + g + 8, c[2]);
            }
        }
    }
}

// ---- per-model programs: which warp owns which parts (see the layout tables in rollout_abi.cu) ---------------------
// MODE 0: accumulate one staged block; MODE 1: flush to global
#define PART(MT, NT, OFF, BIAS, COLY, COLX, OUT)                                  \
    if (MODE == 0) part<NS, MT, NT, OFF, BIAS>(acc, sm, stride, sm, stride, plane, COLY, COLX, lane); \
    else flush<MT, NT, OFF, BIAS>(acc, outs[OUT], lane);

template <int NS, int MODE>
__device__ __forceinline__ void program_mtrssm(float (&acc)[MAX_TILES][4], const __nv_bfloat16* sm, int stride, int plane,
                                               const WgradOut* outs, int warp, int lane) {
    using namespace wgl_mt;
    switch (warp) {
        case 0: PART(2, 4, 0, true, DP + 0, XLD, O_LD) PART(2, 4, 10, false, DP + 0, XLZ, O_LIZ)
                PART(2, 1, 18, false, DP + 0, XLA, O_LIA) break;
        case 1: PART(2, 4, 0, true, DP + 32, XHD, O_HD) PART(2, 2, 10, false, DP + 32, XHI, O_HI)
                PART(1, 4, 14, true, DP + 96, HID + 0, O_LP2) break;
        case 2: PART(2, 4, 0, true, DP + 64, XQ + 0, O_LP1) PART(2, 4, 10, true, DP + 112, XQ + 32, O_HP1) break;
        case 3: PART(2, 8, 0, true, DP + 160, XQ, O_HQ1) break;
        case 4: PART(1, 12, 0, true, DP + 208, XA, O_A1A) PART(1, 4, 13, true, DP + 240, HID + 96, O_A2) break;
        case 5: PART(1, 12, 0, true, DP + 224, XA, O_A1B) PART(1, 4, 13, true, DP + 144, HID + 32, O_HP2) break;
        case 6: PART(1, 12, 0, true, DP + 256, XV, O_V1A) PART(1, 4, 13, true, DP + 288, HID + 128, O_V2) break;
        default: PART(1, 12, 0, true, DP + 272, XV, O_V1B) PART(1, 4, 13, true, DP + 192, HID + 64, O_HQ2) break;
    }
}

template <int NS, int MODE>
__device__ __forceinline__ void program_mrssm(float (&acc)[MAX_TILES][4], const __nv_bfloat16* sm, int stride, int plane,
                                              const WgradOut* outs, int warp, int lane) {
    using namespace wgl_mr;
    switch (warp) {
        case 0: PART(2, 2, 0, true, DP + 0, XASPZ, O_ASP1Z) PART(2, 1, 6, false, DP + 0, XASPA, O_ASP1A)
                PART(2, 4, 8, true, DP + 32, XH1, O_ASP2) break;
        case 1: PART(2, 4, 0, true, DP + 64, XX2, O_IHR) PART(2, 4, 10, true, DP + 96, XX2, O_IHZ) break;
        case 2: PART(2, 4, 0, true, DP + 128, XX2, O_IHN) PART(2, 4, 10, true, DP + 64, XHP, O_HHR) break;
        case 3: PART(2, 4, 0, true, DP + 96, XHP, O_HHZ) PART(2, 4, 10, true, DP + 160, XHP, O_HHN) break;
        case 4: PART(2, 4, 0, true, DP + 192, XA, O_P1) PART(1, 4, 10, true, DP + 224, HID + 0, O_P2)
                PART(1, 4, 15, true, DP + 272, HID + 32, O_A2) break;
        case 5: PART(1, 12, 0, true, DP + 240, XA, O_A1A) PART(1, 4, 13, true, DP + 320, HID + 64, O_V2) break;
        case 6: PART(1, 12, 0, true, DP + 256, XA, O_A1B) PART(1, 6, 13, true, DP + 304, XV, O_V1B_L) break;
        default: PART(1, 12, 0, true, DP + 288, XV, O_V1A) PART(1, 6, 13, false, DP + 304, XV + 48, O_V1B_R) break;
    }
}
#undef PART

// ---- staging: global (fp32 or bf16) -> shared bf16 planes ----------------------------------------------------------
// The staged row is tiled by the segments' 4-element chunk ranges; a per-chunk descriptor table in shared memory
// All of a block's bytes are moved by bulk asynchronous copies (cp.async.bulk, one per (segment,row) piece, completion
// on an mbarrier): every load of the block is in flight at once and none of them holds a register.
//   * bf16 sources (the dpre / saved records on the bf16 path) land directly in their operand position;
//   * fp32 sources land in a raw staging area and are converted (and split, NS = 3) by a second smem->smem pass.
struct ChunkDesc {
    const char* ptr;  // source of this chunk in row 0 (ptr0: in batch element 0 of the initial-state tensor)
    int ld;           // row stride in bytes
    int flags;        // kind (bits 0-1) | shift (bit 2) | valid elements 0..4 (bits 4-6) | raw chunk index (bits 8-15)
};

__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc, int bytes /*4,8,16*/, int src_bytes) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    if (bytes == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
    else if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

// Issue one block: one bulk copy per (segment, row) piece -- bf16 pieces straight into their operand position, fp32
// pieces into the raw area; rows that only allow 8-byte alignment (actions) go through cp.async.  `total_tx` bytes
// are announced on the mbarrier by thread 0.  Dead rows (tail block) are zero-filled by plain stores.
template <int NS>
__device__ __forceinline__ void stage_issue(__nv_bfloat16* sm, float* raw, int stride, int raw_cpr, const WgradSeg* __restrict__ segs,
                                            const int* __restrict__ seg_slot, int nseg, int B, int T, int row_base, uint64_t* bar,
                                            int tid) {
    constexpr int ROWS = rows_of(NS);
    const int R = B * T;
    const int nlive = min(ROWS, R - row_base);
    if (tid == 0) {
        uint32_t per_row = 0;
        for (int si = 0; si < nseg; ++si)
            if (segs[si].kind != 1) per_row += segs[si].valid * (segs[si].kind == 2 ? 2 : 4);
        mbar_expect_tx(bar, per_row * nlive);
    }
    for (int job = tid; job < nseg * ROWS; job += THREADS) {
        const int si = job / ROWS, row = job - si * ROWS;
        const WgradSeg sg = segs[si];
        const int r = row_base + row;
        const int esz = sg.kind == 2 ? 2 : 4;
        char* dst = sg.kind == 2 ? reinterpret_cast<char*>(sm + (size_t)row * stride + sg.c4_begin * 4)
                                 : reinterpret_cast<char*>(raw + ((size_t)row * raw_cpr + seg_slot[si]) * 4);
        const int nbytes = ((sg.valid + 3) & ~3) * esz;  // padded piece (pad columns are zeroed once at kernel start)
        if (r >= R) {  // tail: zero the piece
            for (int o = 0; o < nbytes; o += 8) *reinterpret_cast<uint2*>(dst + o) = make_uint2(0u, 0u);
            continue;
        }
        const char* src;
        if (sg.shift) {
            const int b = r / T, t = r - b * T;
            src = t > 0 ? sg.ptr + (size_t)(r - 1) * sg.ld_bytes : sg.ptr0 + (size_t)b * sg.ld0_bytes;
        } else {
            src = sg.ptr + (size_t)r * sg.ld_bytes;
        }
        if (sg.kind == 1) {  // 8-byte aligned fp32 rows: cp.async pieces (completion via cp.async.wait_group)
            for (int e = 0; e < sg.valid; e += 2) {
                const int n = min(2, sg.valid - e) * 4;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst + e * 4)), "l"(src + e * 4), "r"(n) : "memory");
            }
        } else {
            bulk_g2s(dst, src, sg.valid * esz, bar);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// raw fp32 staging area -> bf16 operand planes
template <int NS>
__device__ __forceinline__ void stage_convert(__nv_bfloat16* sm, const float* raw, int stride, int plane, int raw_cpr,
                                              const uint8_t* __restrict__ chunk_of_raw, int tid) {
    constexpr int ROWS = rows_of(NS);
    // a warp walks whole rows, its lanes the row's chunks: no integer division by the run-time chunk count (it was a quarter of
    // the kernel's instructions, ncu)
    for (int row = tid >> 5; row < ROWS; row += THREADS / 32) {
        for (int k = tid & 31; k < raw_cpr; k += 32) {
            const float4 v = *reinterpret_cast<const float4*>(raw + ((size_t)row * raw_cpr + k) * 4);
            uint32_t lo[NS], hi[NS];
            split_pack<NS>(v.x, v.y, lo);
            split_pack<NS>(v.z, v.w, hi);
            const int off = row * stride + chunk_of_raw[k] * 4;
#pragma unroll
            for (int s = 0; s < NS; ++s) *reinterpret_cast<uint2*>(sm + (size_t)s * plane + off) = make_uint2(lo[s], hi[s]);
        }
    }
}

template <int NS, int MODEL>
__global__ void __launch_bounds__(THREADS, NS == 1 ? 2 : 1) wgrad_mma_kernel(const WgradMmaArgs a) {
    constexpr int ROWS = rows_of(NS);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __nv_bfloat16* sm = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    const int stride = a.stride, plane = ROWS * a.stride;
    float* raw = reinterpret_cast<float*>(sm + (size_t)NS * plane);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // segment table, first raw slot of every fp32 segment, raw-chunk -> staged-chunk map, one mbarrier
    __shared__ WgradSeg segs[MAX_WGRAD_SEGS];
    __shared__ int seg_slot[MAX_WGRAD_SEGS];
    __shared__ uint8_t chunk_of_raw[256];
    __shared__ int raw_cpr_s;
    __shared__ __align__(8) uint64_t bar;
    if (tid < a.nseg) segs[tid] = a.seg[tid];
    if (tid == 0) {
        int nraw = 0;
        for (int si = 0; si < a.nseg; ++si) {
            seg_slot[si] = nraw;
            if (a.seg[si].kind != 2)
                for (int c = a.seg[si].c4_begin; c < a.seg[si].c4_end; ++c) chunk_of_raw[nraw++] = (uint8_t)c;
        }
        raw_cpr_s = nraw;
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int raw_cpr = raw_cpr_s;
    // zero both areas once: pad columns are never written again (bulk copies move only the valid elements)
    for (int i = tid; i < (int)(((size_t)NS * plane * 2 + (size_t)ROWS * raw_cpr * 16) / 16); i += THREADS)
        reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    uint32_t phase = 0;
    float acc[MAX_TILES][4];
#pragma unroll
    for (int i = 0; i < MAX_TILES; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    const int nblocks = (a.B * a.T + ROWS - 1) / ROWS;
    if constexpr (NS == 3) {
        // fp32 records: every segment lands in the raw area and is converted into the operand planes, so the raw area is free as
        // soon as the conversion is done -- the NEXT block's copies are issued there and stream in under this block's MMAs
        int blk = blockIdx.x;
        if (blk < nblocks) stage_issue<NS>(sm, raw, stride, raw_cpr, segs, seg_slot, a.nseg, a.B, a.T, blk * ROWS, &bar, tid);
        for (; blk < nblocks; blk += gridDim.x) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            mbar_wait(&bar, phase);
            phase ^= 1;
            __syncthreads();
            stage_convert<NS>(sm, raw, stride, plane, raw_cpr, chunk_of_raw, tid);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // our reads of the raw area before the next bulk writes
            __syncthreads();
            if (blk + (int)gridDim.x < nblocks)
                stage_issue<NS>(sm, raw, stride, raw_cpr, segs, seg_slot, a.nseg, a.B, a.T, (blk + (int)gridDim.x) * ROWS, &bar, tid);
            if (MODEL == 0) program_mtrssm<NS, 0>(acc, sm, stride, plane, a.out, warp, lane);
            else program_mrssm<NS, 0>(acc, sm, stride, plane, a.out, warp, lane);
            __syncthreads();  // the planes may be rewritten
        }
    } else
    for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        stage_issue<NS>(sm, raw, stride, raw_cpr, segs, seg_slot, a.nseg, a.B, a.T, blk * ROWS, &bar, tid);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        mbar_wait(&bar, phase);
        phase ^= 1;
        __syncthreads();
        stage_convert<NS>(sm, raw, stride, plane, raw_cpr, chunk_of_raw, tid);
        __syncthreads();
        if (MODEL == 0) program_mtrssm<NS, 0>(acc, sm, stride, plane, a.out, warp, lane);
        else program_mrssm<NS, 0>(acc, sm, stride, plane, a.out, warp, lane);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // order our smem reads before the next bulk writes
        __syncthreads();
    }
    if (MODEL == 0) program_mtrssm<NS, 1>(acc, sm, stride, plane, a.out, warp, lane);
    else program_mrssm<NS, 1>(acc, sm, stride, plane, a.out, warp, lane);
}

// =====================================================================================================================
// MMTRSSM, bf16 path: SLAB-staged variant.
// A block's 32 (b,t) rows are consecutive rows of every [B*T, C] tensor, so each source is ONE contiguous slab and one bulk
// copy (5 per block instead of ~450 per-row pieces).  The slabs are DOUBLE-BUFFERED: block n+1 streams in while block n is
// converted and multiplied, so the kernel runs at the speed of its HBM reads.  The bf16 records (dpre 608 B, saved 416 B
// per row: whole 32-byte sectors for the kernels that write them, and 2 mod 4 sixteen-byte chunks so that ldmatrix reads
// them in place with at most 2-way bank conflicts) are used as they land; the fp32 sources (feature rows r0-1 .. r0+31,
// both embeddings) land in a raw area and one pass writes their bf16 operand columns (X plane); previous-step inputs of
// rows with t == 0 come from the initial state.
// =====================================================================================================================
namespace slab {
constexpr int ROWS = 32;
constexpr int DP_LD = MTRSSM_DPRE_FLOATS, SV_LD = MTRSSM_SAVED_FLOATS, X_LD = 392;  // row pitches (bf16 elements)
static_assert((DP_LD * 2 / 16) % 4 == 2 && (SV_LD * 2 / 16) % 4 == 2 && (X_LD * 2 / 16) % 2 == 1, "ldmatrix-friendly row pitches");
static_assert((DP_LD * 2) % 32 == 0 && (SV_LD * 2) % 32 == 0, "record rows are whole 32-byte sectors");
// X plane columns: ordered so that a thread converting the 4-column chunks sub, sub + 8, sub + 16, ... of one row meets one
// source per step of its (unrolled) loop; [d_l | d_h] and [z_l_prev | z_h_prev] stay adjacent (they are one K range each)
constexpr int XDLP = 0, XDHP = 32, XDL = 64, XDH = 96, XEA = 128, XEV = 192, XZLP = 256, XZHP = 272, XACT = 288;  // 296 columns
// shared-memory map (bytes): X plane, then two stages of slabs
constexpr int XP = 0, STAGE0 = ROWS * X_LD * 2;
constexpr int DP = 0, SV = DP + ROWS * DP_LD * 2, FEAT = SV + ROWS * SV_LD * 2, EA = FEAT + (ROWS + 1) * 384, EV = EA + ROWS * 256,
              STAGE_BYTES = EV + ROWS * 256;                 // offsets inside a stage; 61,824 bytes per stage
#ifndef RSSM_WGRAD_STAGES
#define RSSM_WGRAD_STAGES 3  // slab stages per CTA (one CTA per SM): NSTAGE - 1 blocks (62 KB each) are in flight while one is
#endif                       // multiplied; two stages leave the kernel latency x concurrency bound at ~4.4 TB/s
constexpr int NSTAGE = RSSM_WGRAD_STAGES;
static_assert(NSTAGE >= 2 && NSTAGE <= 3, "2 or 3 slab stages");
constexpr int BYTES = STAGE0 + NSTAGE * STAGE_BYTES;        // 210,560 with three stages
}  // namespace slab

template <int MODE>
__device__ __forceinline__ void program_mt_slab(float (&acc)[MAX_TILES][4], const __nv_bfloat16* dp, const __nv_bfloat16* sv,
                                                const __nv_bfloat16* xp, const WgradOut* outs, int warp, int lane) {
    using namespace slab;
    using namespace wgl_mt;
    // dY columns of the dpre record (mtrssm_common.cuh, namespace mtd) and hidden columns of the saved record (mts)
    constexpr int L = 0, H = 32, LP1 = 64, LPL = 96, HP1 = 112, HPL = 144, HQ1 = 160, HQL = 192, A1 = 208, LA = 240, V1 = 256, LV = 288;
    constexpr int LP_HID = 0, HP_HID = 32, HQ_HID = 64, A_HID = 96, V_HID = 128;
#define PX(MT, NT, OFF, BIAS, COLY, COLX) part<1, MT, NT, OFF, BIAS>(acc, dp, DP_LD, xp, X_LD, 0, COLY, COLX, lane)
#define PS(MT, NT, OFF, BIAS, COLY, COLX) part<1, MT, NT, OFF, BIAS>(acc, dp, DP_LD, sv, SV_LD, 0, COLY, COLX, lane)
#define FL(MT, NT, OFF, BIAS, OUT) flush<MT, NT, OFF, BIAS>(acc, outs[OUT], lane)
    switch (warp) {
        case 0:
            if (MODE == 0) { PX(2, 4, 0, true, L, XDLP); PX(2, 4, 10, false, L, XZLP); PX(2, 1, 18, false, L, XACT); }
            else { FL(2, 4, 0, true, O_LD); FL(2, 4, 10, false, O_LIZ); FL(2, 1, 18, false, O_LIA); }
            break;
        case 1:
            if (MODE == 0) { PX(2, 4, 0, true, H, XDHP); PX(2, 2, 10, false, H, XZHP); PS(1, 4, 14, true, LPL, LP_HID); }
            else { FL(2, 4, 0, true, O_HD); FL(2, 2, 10, false, O_HI); FL(1, 4, 14, true, O_LP2); }
            break;
        case 2:
            if (MODE == 0) { PX(2, 4, 0, true, LP1, XDL); PX(2, 4, 10, true, HP1, XDH); }
            else { FL(2, 4, 0, true, O_LP1); FL(2, 4, 10, true, O_HP1); }
            break;
        case 3:
            if (MODE == 0) { PX(2, 8, 0, true, HQ1, XDL); }  // [d_l | d_h] are adjacent
            else { FL(2, 8, 0, true, O_HQ1); }
            break;
        case 4:  // first layer of a modality head: tiles 0..3 = d_l half, 4..11 = embedding half, 12 = bias
            if (MODE == 0) { PX(1, 4, 0, false, A1, XDL); PX(1, 8, 4, true, A1, XEA); PS(1, 4, 13, true, LA, A_HID); }
            else { FL(1, 12, 0, true, O_A1A); FL(1, 4, 13, true, O_A2); }
            break;
        case 5:
            if (MODE == 0) { PX(1, 4, 0, false, A1 + 16, XDL); PX(1, 8, 4, true, A1 + 16, XEA); PS(1, 4, 13, true, HPL, HP_HID); }
            else { FL(1, 12, 0, true, O_A1B); FL(1, 4, 13, true, O_HP2); }
            break;
        case 6:
            if (MODE == 0) { PX(1, 4, 0, false, V1, XDL); PX(1, 8, 4, true, V1, XEV); PS(1, 4, 13, true, LV, V_HID); }
            else { FL(1, 12, 0, true, O_V1A); FL(1, 4, 13, true, O_V2); }
            break;
        default:
            if (MODE == 0) { PX(1, 4, 0, false, V1 + 16, XDL); PX(1, 8, 4, true, V1 + 16, XEV); PS(1, 4, 13, true, HQL, HQ_HID); }
            else { FL(1, 12, 0, true, O_V1B); FL(1, 4, 13, true, O_HQ2); }
            break;
    }
#undef PX
#undef PS
#undef FL
}

__global__ void __launch_bounds__(THREADS, 1) wgrad_mt_slab_kernel(const WgradMtSlabArgs a) {
    using namespace slab;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __nv_bfloat16* xp = reinterpret_cast<__nv_bfloat16*>(smem_raw + XP);
    __shared__ __align__(8) uint64_t bars[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bars[0], 1), mbar_init(&bars[1], 1), mbar_init(&bars[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    float acc[MAX_TILES][4];
#pragma unroll
    for (int i = 0; i < MAX_TILES; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    const int R = a.B * a.T, A = a.A, T = a.T;
    const int nblocks = (R + ROWS - 1) / ROWS;
    // one thread streams a block's five slabs into a stage
    auto issue = [&](int blk, int stage) {
        unsigned char* st = smem_raw + STAGE0 + stage * STAGE_BYTES;
        const int r0 = blk * ROWS, nlive = min(ROWS, R - r0);
        const int fslot = r0 > 0 ? 0 : 1, frow = r0 > 0 ? r0 - 1 : 0, fn = nlive + 1 - fslot;  // feature rows r0-1 .. r0+nlive-1
        mbar_expect_tx(&bars[stage], (uint32_t)nlive * (DP_LD * 2 + SV_LD * 2 + 512) + (uint32_t)fn * 384);
        bulk_g2s(st + DP, a.dpre + (size_t)r0 * DP_LD, nlive * DP_LD * 2, &bars[stage]);
        bulk_g2s(st + SV, a.saved + (size_t)r0 * SV_LD, nlive * SV_LD * 2, &bars[stage]);
        bulk_g2s(st + FEAT + fslot * 384, a.feature + (size_t)frow * 96, fn * 384, &bars[stage]);
        bulk_g2s(st + EA, a.embed_a + (size_t)r0 * 64, nlive * 256, &bars[stage]);
        bulk_g2s(st + EV, a.embed_v + (size_t)r0 * 64, nlive * 256, &bars[stage]);
    };
    if (tid == 0)  // prologue: the first NSTAGE - 1 blocks of this CTA
        for (int k = 0; k < NSTAGE - 1; ++k)
            if ((int)(blockIdx.x + k * gridDim.x) < nblocks) issue(blockIdx.x + k * gridDim.x, k);
    uint32_t phases = 0u;  // bit s = parity of stage s
    int stage = 0;
    // conversion: thread -> row crow, four-column chunks csub + 8k of that row
    const int crow = tid >> 3, csub = tid & 7;
    for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        unsigned char* st = smem_raw + STAGE0 + stage * STAGE_BYTES;
        const __nv_bfloat16* dp = reinterpret_cast<const __nv_bfloat16*>(st + DP);
        const __nv_bfloat16* sv = reinterpret_cast<const __nv_bfloat16*>(st + SV);
        const float* feat = reinterpret_cast<const float*>(st + FEAT);  // slot s = row r0 - 1 + s
        const float* ea = reinterpret_cast<const float*>(st + EA);
        const float* ev = reinterpret_cast<const float*>(st + EV);
        const int r0 = blk * ROWS, nlive = min(ROWS, R - r0);
        {  // prefetch block n + NSTAGE - 1 into the stage block n - 1 used (its readers finished before the last __syncthreads)
            const int nxt = blk + (NSTAGE - 1) * (int)gridDim.x, ns = stage == 0 ? NSTAGE - 1 : stage - 1;
            if (tid == 0 && nxt < nblocks) issue(nxt, ns);
        }
        // the action columns come straight from global memory (24-byte rows): fetch before waiting on the slabs
        float2 av0 = make_float2(0.f, 0.f), av1 = av0;
        if (csub < 2 && crow < nlive) {
            const float* ap = a.actions + (size_t)(r0 + crow) * A;
            if (4 * csub < A) av0 = *reinterpret_cast<const float2*>(ap + 4 * csub);
            if (4 * csub + 2 < A) av1 = *reinterpret_cast<const float2*>(ap + 4 * csub + 2);
        }
        const int r = r0 + crow, b = r / T;
        const bool first_step = r - b * T == 0;  // t == 0: the previous state is the initial state
        mbar_wait(&bars[stage], (phases >> stage) & 1u);
        phases ^= 1u << stage;
        if (nlive < ROWS)  // tail block: stale dY rows of an earlier block must not contribute
            for (int i = tid; i < (ROWS - nlive) * DP_LD / 8; i += THREADS)
                reinterpret_cast<uint4*>(st + DP + nlive * DP_LD * 2)[i] = make_uint4(0u, 0u, 0u, 0u);
        // ---- fp32 -> bf16 operand columns of row crow ----------------------------------------------------------------
        {
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 v[9];
            const bool live = crow < nlive;
            const float* fp = feat + crow * 96;        // row r - 1
            const float* fc = feat + (crow + 1) * 96;  // row r
            const int zoff = csub < 4 ? 80 + 4 * csub : 32 + 4 * (csub - 4);  // z_l_prev <- feature[80:96], z_h_prev <- feature[32:48]
            if (live && !first_step) {
                v[0] = *reinterpret_cast<const float4*>(fp + 48 + 4 * csub);  // d_l_prev
                v[1] = *reinterpret_cast<const float4*>(fp + 4 * csub);       // d_h_prev
                v[8] = *reinterpret_cast<const float4*>(fp + zoff);
            } else if (live) {
                v[0] = *reinterpret_cast<const float4*>(a.deter_l0 + (size_t)b * 32 + 4 * csub);
                v[1] = *reinterpret_cast<const float4*>(a.deter_h0 + (size_t)b * 32 + 4 * csub);
                v[8] = *reinterpret_cast<const float4*>(csub < 4 ? a.stoch_l0 + (size_t)b * 16 + 4 * csub : a.stoch_h0 + (size_t)b * 16 + 4 * (csub - 4));
            } else {
                v[0] = v[1] = v[8] = zero4;
            }
            v[2] = live ? *reinterpret_cast<const float4*>(fc + 48 + 4 * csub) : zero4;  // d_l
            v[3] = live ? *reinterpret_cast<const float4*>(fc + 4 * csub) : zero4;       // d_h
            v[4] = live ? *reinterpret_cast<const float4*>(ea + crow * 64 + 4 * csub) : zero4;
            v[5] = live ? *reinterpret_cast<const float4*>(ea + crow * 64 + 32 + 4 * csub) : zero4;
            v[6] = live ? *reinterpret_cast<const float4*>(ev + crow * 64 + 4 * csub) : zero4;
            v[7] = live ? *reinterpret_cast<const float4*>(ev + crow * 64 + 32 + 4 * csub) : zero4;
            __nv_bfloat16* xr = xp + crow * X_LD + 4 * csub;
#pragma unroll
            for (int k = 0; k < 9; ++k)
                *reinterpret_cast<uint2*>(xr + 32 * k) = make_uint2(pack_bf16(v[k].x, v[k].y), pack_bf16(v[k].z, v[k].w));
            if (csub < 2) *reinterpret_cast<uint2*>(xr + XACT) = make_uint2(pack_bf16(av0.x, av0.y), pack_bf16(av1.x, av1.y));
        }
        __syncthreads();
        program_mt_slab<0>(acc, dp, sv, xp, a.out, warp, lane);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // order our smem reads before the next bulk writes
        __syncthreads();
        stage = stage + 1 == NSTAGE ? 0 : stage + 1;
    }
    program_mt_slab<1>(acc, nullptr, nullptr, xp, a.out, warp, lane);
}

}  // namespace wg

cudaError_t launch_wgrad_mt_slab(const WgradMtSlabArgs& a, cudaStream_t s) {
    auto kernel = wg::wgrad_mt_slab_kernel;
    const size_t smem = wg::slab::BYTES;
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int per_sm = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, wg::THREADS, smem);
    if (err != cudaSuccess) return err;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int nblocks = (a.B * a.T + wg::slab::ROWS - 1) / wg::slab::ROWS;
    int grid = sms * (per_sm > 0 ? per_sm : 1);
    if (grid > nblocks) grid = nblocks;
    kernel<<<grid, wg::THREADS, smem, s>>>(a);
    return cudaGetLastError();
}

template <int NS, int MODEL>
static cudaError_t launch_wgrad_k(const WgradMmaArgs& a, cudaStream_t s) {
    // operand planes + raw fp32 staging of the fp32-sourced chunks
    int raw_chunks = 0;
    for (int i = 0; i < a.nseg; ++i)
        if (a.seg[i].kind != 2) raw_chunks += a.seg[i].c4_end - a.seg[i].c4_begin;
    const size_t smem = (size_t)wg::rows_of(NS) * ((size_t)NS * a.stride * sizeof(__nv_bfloat16) + (size_t)raw_chunks * 16);
    auto kernel = wg::wgrad_mma_kernel<NS, MODEL>;
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int per_sm = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, wg::THREADS, smem);
    if (err != cudaSuccess) return err;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int nblocks = (a.B * a.T + wg::rows_of(NS) - 1) / wg::rows_of(NS);
    int grid = sms * (per_sm > 0 ? per_sm : 1);
    if (grid > nblocks) grid = nblocks;
    kernel<<<grid, wg::THREADS, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_wgrad_mma(const WgradMmaArgs& a, int model, int precision, cudaStream_t s) {
    if (precision == RSSM_PRECISION_FP32) return model == 0 ? launch_wgrad_k<3, 0>(a, s) : launch_wgrad_k<3, 1>(a, s);
    return model == 0 ? launch_wgrad_k<1, 0>(a, s) : launch_wgrad_k<1, 1>(a, s);
}

}  // namespace rssm
