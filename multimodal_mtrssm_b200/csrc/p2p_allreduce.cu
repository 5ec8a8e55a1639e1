// One-shot mean-allreduce of the rollout's weight-gradient bucket over NVLink / NVSwitch PEER MEMORY.
//
// The only exchange of the batch-sharded rollout is one small bucket after the fused backward (66 KB for the default sizes).  NCCL
// costs ~37 us per step for it at 8 GPUs (profiles/r1_n_bench_8gpu.json); running it asynchronously under the next step is worse
// (profiles/r2_u_bench_8gpu.json: the persistent rollout kernels own every SM, the collective's CTAs displace ours and a statically
// partitioned persistent kernel finishes late by the whole delay).  At this size the collective is pure latency, so:
//   every rank keeps its bucket in a peer-mapped allocation (cudaIpc); after its backward a rank (1) tells every peer "bucket of
//   epoch e is complete" (one release store per peer into that peer's flag row), (2) waits for the same word from every peer,
//   (3) reads ALL buckets straight from peer memory, sums them in rank order (every rank computes the identical sum) and writes the
//   mean into its own local output.  One kernel, no staging copy, no second phase: N x 66 KB of NVLink reads per rank.
// Buckets are double-buffered by the caller (epoch parity): rank A rewrites its bucket of epoch e only after its allreduce of
// epoch e+1 completed, which needed every peer's e+1 flag, which a peer raises (stream order) only after its reads of epoch e.
#include <cstdint>
#include <cstdio>

#include "kernels.h"

namespace rssm {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// grid: any; every CTA waits for the flags itself (no grid-wide sync), CTA 0 raises this rank's flag at the peers
__global__ void __launch_bounds__(256) p2p_allreduce_mean_kernel(const P2pAllreduceArgs a) {
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    if (blockIdx.x == 0 && threadIdx.x < a.world) {
        // the bucket was completed by earlier kernels of this stream: visible device-wide at this kernel's start; the release
        // makes it visible to the peer that acquires the flag
        __threadfence_system();
        st_release_sys(a.flags[threadIdx.x] + a.slot * P2P_MAX_RANKS + a.rank, a.epoch);
    }
    __syncthreads();
    if (threadIdx.x < a.world) {
        const uint32_t* f = a.flags[a.rank] + a.slot * P2P_MAX_RANKS + threadIdx.x;
        const long long t0 = clock64();
        // epochs are monotonic per slot: a peer that is already one step ahead has raised a larger epoch
        while ((int32_t)(ld_acquire_sys(f) - a.epoch) < 0) {
            if (clock64() - t0 > a.timeout_cycles) {
                timed_out = 1;
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (timed_out) {  // a peer never arrived: report instead of hanging the GPU (the result is not written)
        if (threadIdx.x == 0 && blockIdx.x == 0) {
            printf("rssm p2p allreduce: rank %d timed out waiting for its peers (epoch %u)\n", a.rank, a.epoch);
            *a.status = 1;
        }
        return;
    }
    const float inv = 1.f / (float)a.world;
    const size_t n4 = a.n >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int r0 = 0; r0 < a.world; r0 += 4) {  // four peer loads in flight
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (r0 + k < a.world) v[k] = __ldcg(reinterpret_cast<const float4*>(a.data[r0 + k]) + i);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (r0 + k < a.world) acc.x += v[k].x, acc.y += v[k].y, acc.z += v[k].z, acc.w += v[k].w;
        }
        reinterpret_cast<float4*>(a.out)[i] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    }
    for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (size_t)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int r0 = 0; r0 < a.world; ++r0) acc += __ldcg(a.data[r0] + i);
        a.out[i] = acc * inv;
    }
}

cudaError_t launch_p2p_allreduce_mean(const P2pAllreduceArgs& a, cudaStream_t s) {
    const size_t n4 = a.n >> 2;
    int grid = (int)((n4 + 255) / 256);
    grid = grid < 1 ? 1 : grid > 64 ? 64 : grid;
    p2p_allreduce_mean_kernel<<<grid, 256, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace rssm
