// Per-chunk cost of the producer / MMA-issuer handshake used by the wide kernels, without any data movement:
// warp 1 lane 0 = producer (wait empty -> arrive full), warp 2 lane 0 = MMA issuer (wait full -> 4 MMAs -> commit empty).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o profiles/src/umma_ring profiles/src/umma_ring.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void arrive_tx0(uint64_t* bar) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(0) : "memory"); }
// variant 0: try_wait loop (as in the kernels); variant 1: test_wait busy poll
template <int V>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        if (V == 0) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        else asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
template <int V>
__global__ void __launch_bounds__(192, 1) ring(int N, int stages, int R, int do_mma, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t full[8], empty[8], fin;
    __shared__ uint32_t tmem_s;
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[i])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[i])), "r"(1));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&fin)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 4 && lane == 0) {
        uint32_t slot = 0, ph = 0;
        for (int r = 0; r < R; ++r) {
            mbar_wait<V>(&empty[slot], ph ^ 1);
            arrive_tx0(&full[slot]);
            if (++slot == (uint32_t)stages) slot = 0, ph ^= 1;
        }
    } else if (warp == 5 && lane == 0) {
        uint32_t slot = 0, ph = 0;
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 16384, idesc = make_idesc(128, N);
        const long long t0 = clock64();
        for (int r = 0; r < R; ++r) {
            mbar_wait<V>(&full[slot], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (do_mma) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma(tmem, make_desc(a0 + kk * 4096, 2048, 128), make_desc(b0 + kk * 2 * N * 16, N * 16, 128), idesc, 1);
            }
            commit(&empty[slot]);
            if (++slot == (uint32_t)stages) slot = 0, ph ^= 1;
        }
        commit(&fin);
        mbar_wait<V>(&fin, 0);
        out[0] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}
int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(ring<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ring<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int R = 512;
    for (int v = 0; v < 2; ++v)
        for (int mma = 0; mma < 2; ++mma)
            for (int stages : {1, 2, 4, 6})
                for (int N : {32, 96}) {
                    long long h = 0;
                    for (int rep = 0; rep < 2; ++rep) {
                        if (v == 0) ring<0><<<1, 192, 64 * 1024>>>(N, stages, R, mma, d);
                        else ring<1><<<1, 192, 64 * 1024>>>(N, stages, R, mma, d);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                    }
                    printf("wait=%s mma=%d stages=%d N=%2d: %7.1f cycles per chunk (4 MMAs; MMA floor 268)\n", v ? "test_wait" : "try_wait ", mma, stages, N, (double)h / R);
                }
    return 0;
}
