// Wide MoPoE-MRSSM: the non-recurrent half of the backward.
//
//  * wide_wgrad_kernel -- weight gradients dW[m][n] += sum over (b,t) rows of Y[row][m] * X[row][n], where Y is a plane of
//    pre-activation gradients (or a head-hidden plane for the logit layers) and X a plane of layer inputs.  The packed
//    [F/8][128][8] planes are MN-major tcgen05 operands over K = row (wide_common.cuh), so a 128-row block of a 128-feature
//    Y tile / an N-feature X tile is ONE contiguous bulk copy each, and a block is 8 MMAs (M = 128, N <= 256, K = 16 rows)
//    into a TMEM accumulator that stays resident over the CTA's whole row range.  Bias gradients ride along as one more
//    MMA against a "ones" operand.  One atomicAdd per element per CTA at the end.
//  * wide_dembed_kernel -- d embed_{a,v}[row][e] = DAH/DVH[row][:] . W1[:, D + e]  (K-major contraction over features).
#include "kernels.h"
#include "wide_common.cuh"

namespace rssm {
namespace wide {

constexpr int WG_STAGES = 2;
constexpr int WG_A_BYTES = 128 * 128 * 2;   // Y tile: 128 features x 128 rows
constexpr int WG_B_BYTES = 256 * 128 * 2;   // X tile: up to 256 features x 128 rows
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_ONES_BYTES = 2 * 128 * 8 * 2;  // two 8-feature groups of the ones block

__host__ __device__ inline size_t wg_smem_bytes() { return 128 + (size_t)WG_STAGES * WG_STAGE_BYTES + WG_ONES_BYTES + 64; }

__global__ void __launch_bounds__(AUX_THREADS, 1) wide_wgrad_kernel(const WideWgradTile* __restrict__ tiles, int nsplit, int T, int NBBT,
                                                                  const __nv_bfloat16* __restrict__ ones) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
    unsigned char* ring = base;
    unsigned char* s_ones = base + (size_t)WG_STAGES * WG_STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(s_ones + WG_ONES_BYTES);
    uint64_t* empty = full + WG_STAGES;
    uint64_t* accbar = empty + WG_STAGES;
    uint64_t* onesbar = accbar + 1;
    uint32_t* tmem_base = reinterpret_cast<uint32_t*>(onesbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const WideWgradTile tl = tiles[blockIdx.x / nsplit];
    const int split = blockIdx.x % nsplit;
    const int nblocks = T * NBBT;  // row blocks (t, bb)
    const int per = (nblocks + nsplit - 1) / nsplit;
    const int rb0 = split * per, rb1 = min(nblocks, rb0 + per);

    if (tid == 0) {
        for (int i = 0; i < WG_STAGES; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
        mbar_init(accbar, 1), mbar_init(onesbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_base;
    const bool bias = tl.db != nullptr;
    const uint32_t a_bytes = WG_A_BYTES, b_bytes = (uint32_t)tl.N * 128 * 2;

    if (warp == 4 && lane == 0) {
        if (bias) {
            mbar_expect_tx(onesbar, WG_ONES_BYTES);
            bulk_g2s(s_ones, ones, WG_ONES_BYTES, onesbar);
        }
        Ring ring_p;
        for (int rb = rb0; rb < rb1; ++rb) {
            const int t = rb / NBBT, bb = rb - t * NBBT;
            const __nv_bfloat16* ya = tl.y + (long long)t * tl.y_tstride + (long long)bb * tl.y_bstride;
            const __nv_bfloat16* xa = (tl.x_shift && t == 0) ? tl.x0 + (long long)bb * tl.x_bstride
                                                             : tl.x + (long long)(t - tl.x_shift) * tl.x_tstride + (long long)bb * tl.x_bstride;
            mbar_wait(&empty[ring_p.slot], ring_p.phase ^ 1);
            mbar_expect_tx(&full[ring_p.slot], a_bytes + b_bytes);
            unsigned char* st = ring + (size_t)ring_p.slot * WG_STAGE_BYTES;
            bulk_g2s(st, ya, a_bytes, &full[ring_p.slot]);
            bulk_g2s(st + WG_A_BYTES, xa, b_bytes, &full[ring_p.slot]);
            ring_p.advance(WG_STAGES);
        }
    } else if (warp == 5 && lane == 0) {
        if (bias) mbar_wait(onesbar, 0);
        Ring ring_c;
        const uint32_t idesc = idesc_bf16(tl.N, 1, 1), idesc1 = idesc_bf16(16, 1, 1);
        for (int rb = rb0; rb < rb1; ++rb) {
            mbar_wait(&full[ring_c.slot], ring_c.phase);
            tc_fence_after();
            const uint32_t a0 = smem_u32(ring + (size_t)ring_c.slot * WG_STAGE_BYTES), b0 = a0 + WG_A_BYTES, o0 = smem_u32(s_ones);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {  // 16 rows per MMA: two 8-row groups of 128 B
                const uint32_t acc = (rb == rb0 && kk == 0) ? 0u : 1u;
                const uint64_t da = smem_desc(a0 + kk * 256, 128, 2048);
                umma(tmem, da, smem_desc(b0 + kk * 256, 128, 2048), idesc, acc);
                if (bias) umma(tmem + 256, da, smem_desc(o0 + kk * 256, 128, 2048), idesc1, acc);
            }
            umma_commit(&empty[ring_c.slot]);
            ring_c.advance(WG_STAGES);
        }
        umma_commit(accbar);
    } else if (warp < 4) {
        if (rb1 > rb0) {
            mbar_wait(accbar, 0);
            tc_fence_after();
            const int m = warp * 32 + lane;  // Y feature inside the tile
            const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
            const bool mine = m < tl.mvalid;  // the TMEM loads are warp-collective: every lane executes them
            for (int n0 = 0; n0 < tl.N; n0 += 8) {
                float v[8];
                tmem_ld8(tlane + n0, v);
                if (mine) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (n0 + i < tl.nvalid) atomicAdd(tl.dW + (long long)m * tl.sm + (long long)(n0 + i) * tl.sn, v[i]);
                }
            }
            if (bias) {
                float v[8];
                tmem_ld8(tlane + 256, v);
                if (mine) atomicAdd(tl.db + m, v[0]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    __syncwarp();
    if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// d embed: grid (T * NBBT, 2).  A = DAH / DVH plane chunks (K-major), B = transposed embedding columns of W1 [64 x D].
constexpr int DE_STAGES = 4;
constexpr int DE_STAGE_BYTES = A_BYTES + 64 * 64 * 2;
__host__ __device__ inline size_t de_smem_bytes() { return 128 + (size_t)DE_STAGES * DE_STAGE_BYTES + 128; }

__global__ void __launch_bounds__(AUX_THREADS, 1) wide_dembed_kernel(const WideDembedArgs p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)DE_STAGES * DE_STAGE_BYTES);
    uint64_t* empty = full + DE_STAGES;
    uint64_t* accbar = empty + DE_STAGES;
    uint32_t* tmem_base = reinterpret_cast<uint32_t*>(accbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = blockIdx.x / p.NBBT, bb = blockIdx.x - t * p.NBBT, mod = blockIdx.y;
    const int KC = p.D >> 6;
    if (tid == 0) {
        for (int i = 0; i < DE_STAGES; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
        mbar_init(accbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_base;
    const __nv_bfloat16* ya = (mod == 0 ? p.dah : p.dvh) + (long long)t * p.dt_stride + (long long)bb * p.D * BM;
    const __nv_bfloat16* wt = mod == 0 ? p.pWaeT : p.pWveT;
    if (warp == 4 && lane == 0) {
        Ring r;
        for (int c = 0; c < KC; ++c) {
            mbar_wait(&empty[r.slot], r.phase ^ 1);
            mbar_expect_tx(&full[r.slot], A_BYTES + 64 * 64 * 2);
            unsigned char* st = ring + (size_t)r.slot * DE_STAGE_BYTES;
            bulk_g2s(st, ya + (long long)c * (BM * 64), A_BYTES, &full[r.slot]);
            bulk_g2s(st + A_BYTES, wt + (long long)c * (64 * 64), 64 * 64 * 2, &full[r.slot]);
            r.advance(DE_STAGES);
        }
    } else if (warp == 5 && lane == 0) {
        Ring r;
        const uint32_t idesc = idesc_bf16(64, 0, 0);
        for (int c = 0; c < KC; ++c) {
            mbar_wait(&full[r.slot], r.phase);
            tc_fence_after();
            const uint32_t a0 = smem_u32(ring + (size_t)r.slot * DE_STAGE_BYTES), b0 = a0 + A_BYTES;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                umma(tmem, smem_desc(a0 + kk * 2 * (BM * 16), BM * 16, 128), smem_desc(b0 + kk * 2 * (64 * 16), 64 * 16, 128), idesc,
                     (c == 0 && kk == 0) ? 0u : 1u);
            umma_commit(&empty[r.slot]);
            r.advance(DE_STAGES);
        }
        umma_commit(accbar);
    } else if (warp < 4) {
        mbar_wait(accbar, 0);
        tc_fence_after();
        const int row = bb * BM + warp * 32 + lane;
        float* out = (mod == 0 ? p.d_embed_a : p.d_embed_v) + ((long long)row * p.T + t) * 64;
        const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            float v[16];
            tmem_ld16(tlane + q * 16, v);
            if (row < p.B) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(out + q * 16 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    __syncwarp();
    if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

// transposed embedding columns: dst[mod][c][kg][e (64)][8] = W1_mod[k][D + e]; ones block [16][128][8]: feature 0 = 1
__global__ void wide_pack_dembed_kernel(const float* __restrict__ au_w1, const float* __restrict__ vi_w1, int D, __nv_bfloat16* __restrict__ dstA,
                                        __nv_bfloat16* __restrict__ dstV, __nv_bfloat16* __restrict__ ones) {
    const int KC = D >> 6;
    const long long total = (long long)KC * 8 * 64;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * total; i += (long long)gridDim.x * blockDim.x) {
        const int mod = i >= total;
        const long long ii = i - mod * total;
        const int e = (int)(ii % 64), kg = (int)((ii / 64) % 8), c = (int)(ii / 512);
        const float* w = mod ? vi_w1 : au_w1;
        float v[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) v[x] = w[(long long)(c * 64 + kg * 8 + x) * (D + 64) + D + e];
        *reinterpret_cast<uint4*>((mod ? dstV : dstA) + ii * 8) = pack8(v);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 16 * 128; i += gridDim.x * blockDim.x) {
        float v[8] = {i < 128 ? 1.f : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        *reinterpret_cast<uint4*>(ones + (long long)i * 8) = pack8(v);
    }
}

__global__ void wide_wgrad_tiles_upload_kernel(const __grid_constant__ WideWgradTileTable tb, WideWgradTile* __restrict__ dst) {
    for (int i = threadIdx.x; i < tb.n; i += blockDim.x) dst[i] = tb.t[i];
}

}  // namespace wide

cudaError_t launch_wide_wgrad_tiles_upload(const WideWgradTileTable& table, WideWgradTile* dst, cudaStream_t s) {
    wide::wide_wgrad_tiles_upload_kernel<<<1, 128, 0, s>>>(table, dst);
    return cudaGetLastError();
}

cudaError_t launch_wide_wgrad(const WideWgradTile* tiles_dev, int ntiles, int nsplit, int T, int NBBT, const __nv_bfloat16* ones, cudaStream_t s) {
    const size_t smem = wide::wg_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(wide::wide_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    wide::wide_wgrad_kernel<<<ntiles * nsplit, wide::AUX_THREADS, smem, s>>>(tiles_dev, nsplit, T, NBBT, ones);
    return cudaGetLastError();
}

cudaError_t launch_wide_dembed(const WideDembedArgs& a, cudaStream_t s) {
    const size_t smem = wide::de_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(wide::wide_dembed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    wide::wide_dembed_kernel<<<dim3(a.T * a.NBBT, 2), wide::AUX_THREADS, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_wide_pack_dembed(const float* au_w1, const float* vi_w1, int D, __nv_bfloat16* dstA, __nv_bfloat16* dstV, __nv_bfloat16* ones,
                                    cudaStream_t s) {
    wide::wide_pack_dembed_kernel<<<32, 256, 0, s>>>(au_w1, vi_w1, D, dstA, dstV, ones);
    return cudaGetLastError();
}

}  // namespace rssm
