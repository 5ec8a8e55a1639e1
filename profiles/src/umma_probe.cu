// Standalone probe of tcgen05.mma (kind::f16, bf16 operands, MN-major / no-swizzle smem operands, fp32 TMEM accumulator).
// Validates on a B200 the descriptor conventions the fused wgrad uses:
//   operand element (mn, k) at  (mn/8)*SBO + (k/8)*LBO + (k%8)*16 + (mn%8)*2   bytes   [to be confirmed: which field is which]
// and whether MMAs issued by different warps may accumulate into the same TMEM tile.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o profiles/src/umma_probe profiles/src/umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
    return d;                // layout_type = 0 (no swizzle), base_offset = 0, lbo_mode = 0
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4)      // c_format = F32
           | (1u << 7)    // a_format = BF16
           | (1u << 10)   // b_format = BF16
           | (1u << 15)   // a_major = MN
           | (1u << 16)   // b_major = MN
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(a), "r"(parity)
                     : "memory");
        if (done) return true;
    }
    return false;
}

constexpr int KROWS = 16;         // MMA K (bf16)
constexpr int A_CHUNKS = 20;      // 160 MN columns staged for A (window of 16 chunks can start at 0..4)
constexpr int B_CHUNKS = 8;       // 64 MN columns staged for B
constexpr int CHUNK_BYTES = 256;  // [2 kgroups][8 rows][8 cols] bf16

// element (mn, k) of an operand staged as [chunk][kgroup][8][8]
__host__ __device__ inline int op_off(int mn, int k) { return (mn >> 3) * 128 + (k >> 3) * 64 + (k & 7) * 8 + (mn & 7); }

struct Params {
    const __nv_bfloat16* A;  // [KROWS][A_CHUNKS*8] row-major (k, mn)
    const __nv_bfloat16* B;  // [KROWS][B_CHUNKS*8]
    float* D;                // [variants][128][64]
    int* status;
};

// variant v: bit0 swaps the (LBO,SBO) roles; N = 16 * (1 + (v >> 1) % 4); a_start chunk = (v >> 3)
__global__ void __launch_bounds__(256, 1) probe(Params p, int nvariants, int stress_iters) {
    __shared__ __align__(128) __nv_bfloat16 sA[A_CHUNKS * 128];
    __shared__ __align__(128) __nv_bfloat16 sB[B_CHUNKS * 128];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < KROWS * A_CHUNKS * 8; i += blockDim.x) {
        const int k = i / (A_CHUNKS * 8), mn = i % (A_CHUNKS * 8);
        sA[op_off(mn, k)] = p.A[i];
    }
    for (int i = tid; i < KROWS * B_CHUNKS * 8; i += blockDim.x) {
        const int k = i / (B_CHUNKS * 8), mn = i % (B_CHUNKS * 8);
        sB[op_off(mn, k)] = p.B[i];
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    uint32_t phase = 0;
    int ok = 1;

    for (int v = 0; v < nvariants; ++v) {
        const int swap = v & 1, N = 16 * (1 + ((v >> 1) & 3)), a0 = v >> 3;
        // chunk stride 256 B, k-group stride 128 B
        const uint32_t lbo = swap ? 256 : 128, sbo = swap ? 128 : 256;
        if (tid == 0) {
            const uint64_t da = make_desc(smem_u32(sA) + a0 * CHUNK_BYTES, lbo, sbo);
            const uint64_t db = make_desc(smem_u32(sB), lbo, sbo);
            umma(tmem, da, db, make_idesc(128, N), 0);
            umma_commit(&bar);
        }
        if (!mbar_wait(&bar, phase)) ok = 0;
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp < 4) {  // warp w reads TMEM lanes 32w .. 32w+31; thread = one accumulator row
            for (int c0 = 0; c0 < 64; c0 += 16) {
                uint32_t r[16];
                const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int j = 0; j < 16; ++j) p.D[((size_t)v * 128 + 32 * warp + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    // ---- stress: all 8 warps' lane 0 issue accumulating MMAs into the SAME tile concurrently ----------------------
    {
        __shared__ __align__(8) uint64_t bars[8];
        if (tid == 0) {
            for (int w = 0; w < 8; ++w) mbar_init(&bars[w], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            // zero-init the tile with one non-accumulating MMA of B against itself?  simpler: first MMA overwrites
            const uint64_t da = make_desc(smem_u32(sA), 128, 256), db = make_desc(smem_u32(sB), 128, 256);
            umma(tmem + 64, da, db, make_idesc(128, 32), 0);
            umma_commit(&bar);
        }
        if (!mbar_wait(&bar, phase)) ok = 0;
        phase ^= 1;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t ph = 0;
        for (int it = 0; it < stress_iters; ++it) {
            if (lane == 0) {
                const uint64_t da = make_desc(smem_u32(sA), 128, 256), db = make_desc(smem_u32(sB), 128, 256);
                umma(tmem + 64, da, db, make_idesc(128, 32), 1);
                umma_commit(&bars[warp]);
            }
            __syncwarp();
            if (!mbar_wait(&bars[warp], ph)) ok = 0;
            ph ^= 1;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp < 4) {
            for (int c0 = 0; c0 < 32; c0 += 16) {
                uint32_t r[16];
                const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + 64 + c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int j = 0; j < 16; ++j) p.D[((size_t)nvariants * 128 + 32 * warp + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
    if (tid == 0) *p.status = ok;
}

int main() {
    const int AC = A_CHUNKS * 8, BC = B_CHUNKS * 8, NV = 8 * 5, STRESS = 200;
    std::vector<float> hA(KROWS * AC), hB(KROWS * BC);
    std::vector<__nv_bfloat16> bA(hA.size()), bB(hB.size());
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) hA[i] = (float)(rand() % 7 - 3), bA[i] = __float2bfloat16(hA[i]);
    for (size_t i = 0; i < hB.size(); ++i) hB[i] = (float)(rand() % 5 - 2), bB[i] = __float2bfloat16(hB[i]);
    __nv_bfloat16 *dA, *dB;
    float* dD;
    int* dS;
    cudaMalloc(&dA, bA.size() * 2), cudaMalloc(&dB, bB.size() * 2), cudaMalloc(&dD, (NV + 1) * 128 * 64 * 4), cudaMalloc(&dS, 4);
    cudaMemcpy(dA, bA.data(), bA.size() * 2, cudaMemcpyHostToDevice), cudaMemcpy(dB, bB.data(), bB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, (NV + 1) * 128 * 64 * 4);
    Params p{dA, dB, dD, dS};
    probe<<<1, 256>>>(p, NV, STRESS);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> hD((NV + 1) * 128 * 64);
    int st = 0;
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost), cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    printf("barrier waits ok: %d\n", st);
    for (int v = 0; v < NV; ++v) {
        const int swap = v & 1, N = 16 * (1 + ((v >> 1) & 3)), a0 = v >> 3;
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                float ref = 0.f;
                for (int k = 0; k < KROWS; ++k) ref += hA[k * AC + a0 * 8 + m] * hB[k * BC + n];
                if (hD[((size_t)v * 128 + m) * 64 + n] != ref) ++bad;
            }
        printf("variant %2d: swap=%d N=%d a_start_chunk=%d  mismatches=%d / %d\n", v, swap, N, a0, bad, 128 * N);
    }
    {
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 32; ++n) {
                float ref = 0.f;
                for (int k = 0; k < KROWS; ++k) ref += hA[k * AC + m] * hB[k * BC + n];
                ref *= (float)(1 + 8 * STRESS);
                if (hD[((size_t)NV * 128 + m) * 64 + n] != ref) ++bad;
            }
        printf("8-warp concurrent accumulate (%d MMAs into one tile): mismatches=%d / %d\n", 1 + 8 * STRESS, bad, 128 * 32);
    }
    return 0;
}
