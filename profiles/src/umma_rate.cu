// tcgen05.mma issue/execution rate by N and shared-memory operand layout (bf16, M = 128, K = 16 per instruction, cta_group::1).
// Times R chunks of 4 back-to-back MMAs (one K = 64 operand chunk each), a commit per chunk, one wait at the end.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o profiles/src/umma_rate profiles/src/umma_rate.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// mode 0: no swizzle, K-major (LBO = rows*16, SBO = 128); mode 1: SWIZZLE_128B K-major (SBO = 1024, +32 B per K step)
__global__ void __launch_bounds__(128, 1) rate(int N, int mode, int R, int commit_each, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t tmem_s;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(R + 8));
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
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(base), b0 = a0 + 16384;
        const uint32_t idesc = make_idesc(128, N);
        const long long t0 = clock64();
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint64_t da, db;
                if (mode == 0) da = make_desc(a0 + kk * 2 * 2048, 2048, 128, 0), db = make_desc(b0 + kk * 2 * N * 16, N * 16, 128, 0);
                else da = make_desc(a0 + kk * 32, 16, 1024, 2), db = make_desc(b0 + kk * 32, 16, 1024, 2);
                umma(tmem, da, db, idesc, 1);
            }
            if (commit_each) commit(&bar[1]);
        }
        const long long t1 = clock64();
        commit(&bar[0]);
        mbar_wait(&bar[0], 0);
        const long long t2 = clock64();
        out[0] = t1 - t0, out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}
int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    const int R = 512;
    for (int mode = 0; mode < 2; ++mode)
        for (int ce = 0; ce < 2; ++ce)
            for (int N : {32, 64, 96, 128, 256}) {
                long long h[2];
                for (int rep = 0; rep < 2; ++rep) {
                    rate<<<1, 128, 60 * 1024>>>(N, mode, R, ce, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                }
                printf("layout %s commit/chunk %d N=%3d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (floor %d)\n", mode ? "SW128 " : "noswz ", ce, N,
                       (double)h[0] / (4.0 * R), (double)h[1] / (4.0 * R), 128 * N / 256);
            }
    return 0;
}
