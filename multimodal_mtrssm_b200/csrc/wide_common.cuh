// Shared pieces of the WIDE MoPoE-MRSSM rollout kernels (deterministic_size = hidden_size = D, D % 64 == 0, 64 <= D <= 512:
// BASELINE.json cfg3, "hidden 512").
//
// At D = 512 the weights (2.7 M scalars) no longer fit one SM and one step is a chain of dense [B x D] x [D x 3D] contractions,
// so the design of the default-size kernels (one warp walks 16 sequences in registers) does not apply.  Here
//   * ONE persistent cooperative kernel walks all T steps; CTA (bb, s) owns batch block bb (128 sequences) x slice s
//     (32 hidden units / head features) for every phase of a step,
//   * the contractions run on tcgen05 (cta_group::1, kind::f16, M = 128 batch rows, N = 32 / 96 features, K = 16) with fp32
//     TMEM accumulators; the operands are staged by 1-D TMA bulk copies (cp.async.bulk + mbarrier ring),
//   * between the dependent phases of a step the CTAs exchange activations through L2 in the tensor core's own operand
//     layout and meet at a grid barrier.
//
// PACKED OPERAND LAYOUT ("k8-interleaved").  A [rows x F] bf16 matrix is stored per 128-row block as [F/8][128][8]:
//     element (r, f) of block bb at  ((bb * F/8 + f/8) * 128 + r) * 8 + f % 8.
// A 64-column K chunk of a block is one contiguous 16 KB piece = one bulk copy, and it is at the same time
//   - the canonical no-swizzle K-MAJOR tcgen05 operand with K = feature (LBO = 128 rows * 16 B between 8-column groups,
//     SBO = 128 B between 8-row groups): used by the forward / backward contractions over features, and
//   - the canonical no-swizzle MN-MAJOR operand with K = row (LBO = 128 B between 8-row groups, SBO = 2048 B between
//     8-feature groups): used by the weight-gradient contractions over (b,t) rows (layout validated by profiles/src/umma_probe.cu).
// Weights are packed once per call the same way, per slice: [slice][K/64][8][N][8] (N = 32 or 96 rows).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "frag.cuh"

namespace rssm {
namespace wide {

constexpr int BM = 128;                // batch rows per block = MMA M
#ifndef RSSM_WIDE_ISSUERS
#define RSSM_WIDE_ISSUERS 2
#endif
constexpr int N_ISSUERS = RSSM_WIDE_ISSUERS;  // MMA issuer threads (1 or 2; build-time switch -DRSSM_WIDE_ISSUERS=2), see below
constexpr int NTHREADS = 320 + 32 * (N_ISSUERS - 1);  // warps 0-7: epilogue (TMEM lane quadrant = warp & 3, column half = warp >> 2), warp 8: producer, warps 9(-10): MMA issuer(s)
constexpr int PRODUCER_WARP = 8, MMA_WARP = 9, MMA_WARP2 = 10, EPI_THREADS = 256;
// Two issuer threads take alternate operand chunks: one thread's wait -> 4 MMAs -> commit loop costs ~470 cycles per chunk
// (mbarrier ops ~130 cycles each and ~47 cycles per tcgen05.mma issue, all serial in the thread: profiles/src/umma_ring*.cu) against
// 268 cycles of tensor-pipe time, so a single issuer leaves the pipe idle 40 % of the time (cfg3, same box: 6.70 -> 6.29 ms).
// Each issuer accumulates its chunks (even / odd) into ITS OWN copy of the accumulator (TMEM columns + acc_off); the epilogue
// adds the two copies.  So every accumulator is written by one thread in a fixed order and the results stay bit-reproducible
// (two threads accumulating into one tile would make the summation order a race).
constexpr int AUX_THREADS = 192;       // non-recurrent kernels: warps 0-3 epilogue, warp 4 producer, warp 5 MMA issuer
constexpr int A_BYTES = BM * 64 * 2;   // one K chunk (64 columns) of an activation block
constexpr int NPLANES = 10;            // record planes per step
enum Plane { P_HID1 = 0, P_X2, P_R, P_Z, P_N, P_HN, P_HB, P_PH, P_AH, P_VH };
// gradient planes per step of the backward workspace, followed by two narrow planes: d logits [48] and [a_t ; z_{t-1}] [32]
constexpr int NDPLANES = 9;
enum DPlane { DP_PH = 0, DP_AH, DP_VH, DP_GR, DP_GZ, DP_GIN, DP_GHN, DP_X2, DP_H1 };

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// element offset of (row r of block bb, feature f) in a packed [*, F] matrix
__host__ __device__ inline long long pk_off(int bb, int r, int f, int F) {
    return (((long long)bb * (F >> 3) + (f >> 3)) * BM + r) * 8 + (f & 7);
}

// ---- tcgen05 plumbing ---------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) |
           ((uint64_t)1 << 46);
}
// instruction descriptor: bf16 x bf16 -> fp32, M = 128; a_mn / b_mn = 1: operand is MN-major (else K-major)
__host__ __device__ constexpr uint32_t idesc_bf16(int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// multicast variants for CTA clusters: the slice CTAs of a batch block read the SAME activation chunks, so each CTA of a cluster
// fetches 1/CS of a chunk and the copy lands in every CTA of `mask` (same CTA-relative offsets, each CTA's own mbarrier gets the
// bytes); a slot is reusable once EVERY CTA's MMAs have read it, so the commit arrives on the `empty` barrier of all of them.
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async;" ::: "memory"); }

// named barrier of the 8 epilogue warps
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 8 consecutive accumulator columns of this thread's TMEM lane (32 * (warp & 3) + lane)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ void unpack8(const uint4 q, float (&v)[8]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// this thread's 8 / 16 accumulator columns, summed over the issuers' copies
__device__ __forceinline__ void acc_ld8(uint32_t taddr, uint32_t acc_off, float (&v)[8]) {
    tmem_ld8(taddr, v);
    if (N_ISSUERS > 1) {
        float w[8];
        tmem_ld8(taddr + acc_off, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += w[i];
    }
}
__device__ __forceinline__ void acc_ld16(uint32_t taddr, uint32_t acc_off, float (&v)[16]) {
    tmem_ld16(taddr, v);
    if (N_ISSUERS > 1) {
        float w[16];
        tmem_ld16(taddr + acc_off, w);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += w[i];
    }
}

// ---- block-group barrier ---------------------------------------------------------------------------------------------
// All CTAs are co-resident (cooperative launch).  Only the `nparts` slice CTAs of one batch block exchange data, so each batch
// block has its own arrival counter (128 concurrent atomics on ONE address cost ~1.8 us on B200, 16 cost ~0.25 us).  `ctr`
// counts arrivals monotonically; `epoch` is this thread's copy of the target.  Generic-proxy global writes made before the barrier are ordered before the async-proxy (bulk copy) reads other CTAs
// issue after it (fence.proxy.async on both sides), TMEM accesses likewise (tcgen05 fences).
__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned& epoch, int* status, unsigned nparts) {
    proxy_fence();  // this thread's generic-proxy writes -> async proxy (other CTAs read them with bulk copies)
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        // one cumulative release for the whole CTA (the CTA barrier above ordered the other threads' writes before it)
        epoch += nparts;
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned seen, spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
            if (++spins > (1u << 27)) {  // a lost CTA must not hang the GPU
                *status = 2;
                __trap();
            }
        } while ((int)(seen - epoch) < 0);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
    tc_fence_after();
    proxy_fence();
}

// ---- bulk-copy ring shared by the producer and the MMA issuer -----------------------------------------------------------
struct Ring {
    uint32_t slot = 0, phase = 0;
    __device__ __forceinline__ void advance(int stages) {
        if (++slot == (uint32_t)stages) slot = 0, phase ^= 1;
    }
};

}  // namespace wide
}  // namespace rssm
