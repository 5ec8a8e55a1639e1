// Batched weight gradients of the rollout on the tensor cores.
//
// For every layer:  dW[n][k] += sum_{(b,t)} dY[(b,t)][n] * X[(b,t)][k],   db[n] += sum_{(b,t)} dY[(b,t)][n]
// with dY a slice of the backward kernel's per-(b,t) "dpre" record and X the layer's input (outputs of the
// forward kernel, its saved record, or kernel inputs; "previous-step" inputs are the row (b,t-1), or the
// initial state for t == 0).  The recurrence only carries data gradients, so this pass is embarrassingly
// parallel over (b,t).
//
// Design: a CTA stages ROWS = 32 consecutive (b,t) rows ONCE from HBM (coalesced float4) into shared memory
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

constexpr int ROWS = 32;
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
__device__ __forceinline__ void part(float (&acc)[MAX_TILES][4], const __nv_bfloat16* __restrict__ sm, int stride, int plane,
                                     int colY, int colX, int lane) {
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
                ldmatrix_x4_trans(a[s][mt], sm + (size_t)s * plane + (size_t)(r0 + lr + m16 * 8) * stride + colY + 16 * mt + m8 * 8);
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
                const __nv_bfloat16* base = sm + (size_t)s * plane + (size_t)(r0 + lr + m8 * 8) * stride + colX + 16 * np;
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
    if (MODE == 0) part<NS, MT, NT, OFF, BIAS>(acc, sm, stride, plane, COLY, COLX, lane); \
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
// The staged row is tiled by the segments' 4-element chunk ranges.  A per-chunk descriptor table in shared memory
// (built once per CTA) turns "chunk c of row r" into one address computation.  A warp stages whole rows: its lanes
// first ISSUE the loads of all their chunks of the row (independent 16/8-byte loads in flight together), then
// convert and store -- one DRAM round trip per row per warp, and no per-chunk integer division.
struct ChunkDesc {
    const char* ptr;  // source of this chunk in row 0 (ptr0: in batch element 0 of the initial-state tensor)
    int ld;           // row stride in bytes
    int flags;        // kind (bits 0-1) | shift (bit 2) | valid elements 0..4 (bits 4-6)
};

template <int NS>
__device__ __forceinline__ void stage_block(__nv_bfloat16* sm, int stride, int plane, const ChunkDesc* __restrict__ desc,
                                            const ChunkDesc* __restrict__ desc0, int B, int T, int row_base, int warp, int lane) {
    constexpr int MAXJ = 8;  // up to 256 chunks (1024 staged columns) per row
    const int R = B * T;
    const int cpr = stride >> 2;
    for (int row = warp; row < ROWS; row += THREADS / 32) {
        const int r = row_base + row;
        const bool live = r < R;
        const int b = r / T, t = r - b * T;
        uint4 raw[MAXJ];
        int flg[MAXJ];
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            const int c = lane + 32 * j;
            raw[j] = make_uint4(0u, 0u, 0u, 0u);
            flg[j] = 0;
            if (c < cpr) {
                ChunkDesc d = desc[c];
                const int nvalid = live ? (d.flags >> 4) & 7 : 0;
                flg[j] = (d.flags & 3) | (nvalid << 4) | 0x100;
                if (nvalid > 0) {
                    const char* src;
                    if (d.flags & 4) {
                        if (t > 0) {
                            src = d.ptr + (size_t)(r - 1) * d.ld;
                        } else {
                            const ChunkDesc d0 = desc0[c];
                            src = d0.ptr + (size_t)b * d0.ld;
                        }
                    } else {
                        src = d.ptr + (size_t)r * d.ld;
                    }
                    const int kind = d.flags & 3;
                    if (kind == 0) {
                        raw[j] = *reinterpret_cast<const uint4*>(src);
                    } else if (kind == 2) {
                        const uint2 q = *reinterpret_cast<const uint2*>(src);
                        raw[j].x = q.x, raw[j].y = q.y;
                    } else {
                        const float* f = reinterpret_cast<const float*>(src);
                        raw[j].x = __float_as_uint(f[0]);
                        if (nvalid > 1) raw[j].y = __float_as_uint(f[1]);
                        if (nvalid > 2) raw[j].z = __float_as_uint(f[2]);
                        if (nvalid > 3) raw[j].w = __float_as_uint(f[3]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            if (!(flg[j] & 0x100)) continue;
            const int off = row * stride + (lane + 32 * j) * 4;
            if ((flg[j] & 3) == 2) {  // already bf16 (bf16 path only)
                *reinterpret_cast<uint2*>(sm + off) = make_uint2(raw[j].x, raw[j].y);
            } else {
                uint32_t lo[NS], hi[NS];
                split_pack<NS>(__uint_as_float(raw[j].x), __uint_as_float(raw[j].y), lo);
                split_pack<NS>(__uint_as_float(raw[j].z), __uint_as_float(raw[j].w), hi);
#pragma unroll
                for (int s = 0; s < NS; ++s) *reinterpret_cast<uint2*>(sm + (size_t)s * plane + off) = make_uint2(lo[s], hi[s]);
            }
        }
    }
}

template <int NS, int MODEL>
__global__ void __launch_bounds__(THREADS, NS == 1 ? 2 : 1) wgrad_mma_kernel(const WgradMmaArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* sm = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    const int stride = a.stride, plane = ROWS * a.stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // per-chunk descriptors in shared memory (per-thread indexing of kernel parameters would push the whole
    // parameter block into local memory)
    __shared__ ChunkDesc desc[256], desc0[256];
    for (int c = tid; c < (stride >> 2); c += THREADS) {
        int si = 0;
        while (c >= a.seg[si].c4_end) ++si;
        const WgradSeg& sg = a.seg[si];
        const int e = (c - sg.c4_begin) * 4, esz = sg.kind == 2 ? 2 : 4;
        int nvalid = sg.valid - e;
        nvalid = nvalid > 4 ? 4 : (nvalid < 0 ? 0 : nvalid);
        desc[c].ptr = sg.ptr + (size_t)e * esz, desc[c].ld = sg.ld_bytes;
        desc[c].flags = sg.kind | (sg.shift ? 4 : 0) | (nvalid << 4);
        desc0[c].ptr = sg.ptr0 + (size_t)e * esz, desc0[c].ld = sg.ld0_bytes, desc0[c].flags = 0;
    }
    __syncthreads();
    float acc[MAX_TILES][4];
#pragma unroll
    for (int i = 0; i < MAX_TILES; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    const int nblocks = (a.B * a.T + ROWS - 1) / ROWS;
    for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        stage_block<NS>(sm, stride, plane, desc, desc0, a.B, a.T, blk * ROWS, warp, lane);
        __syncthreads();
        if (MODEL == 0) program_mtrssm<NS, 0>(acc, sm, stride, plane, a.out, warp, lane);
        else program_mrssm<NS, 0>(acc, sm, stride, plane, a.out, warp, lane);
        __syncthreads();
    }
    if (MODEL == 0) program_mtrssm<NS, 1>(acc, sm, stride, plane, a.out, warp, lane);
    else program_mrssm<NS, 1>(acc, sm, stride, plane, a.out, warp, lane);
}

}  // namespace wg

template <int NS, int MODEL>
static cudaError_t launch_wgrad_k(const WgradMmaArgs& a, cudaStream_t s) {
    const size_t smem = (size_t)NS * wg::ROWS * a.stride * sizeof(__nv_bfloat16);
    auto kernel = wg::wgrad_mma_kernel<NS, MODEL>;
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int per_sm = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, wg::THREADS, smem);
    if (err != cudaSuccess) return err;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int nblocks = (a.B * a.T + wg::ROWS - 1) / wg::ROWS;
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
