// What does the rollout kernels' OUTPUT pattern cost?  Each warp owns 16 rows of several [B, T, C] tensors and, per step, writes
// one row segment per tensor per row (forward kernel: 384 + 384 + 128 + 128 + 4 x 64 bytes = 1280 B per row-step).
//   mode 0: mma-fragment pattern straight to global: lane (g, t) stores 16 bytes of rows g and g + 8 (8 rows x 64 B per instruction)
//   mode 1: the same fragment stores go to shared memory; one asynchronous bulk store (cp.async.bulk) per row segment
//   mode 2: staged in shared memory; warp-coalesced 16-byte global stores (consecutive lanes -> consecutive addresses of a row)
//   mode 3: no stores (compute stand-in only)
// A dependent-FMA loop stands in for the step's math.  Prints ms per launch and written GB/s at 4 and 8 warps per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/src/store_bw profiles/src/store_bw.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

constexpr int NSEG = 8;
__constant__ int c_seg_bytes[NSEG] = {384, 384, 128, 128, 64, 64, 64, 64};
constexpr int ROW_BYTES = 1280;       // sum of the segments
constexpr int PITCH = ROW_BYTES + 64;  // staged row pitch (64 mod 128: conflict-free fragment stores)

struct Ptrs {
    char* t[NSEG];
};

template <int MODE>
__global__ void __launch_bounds__(128) store_kernel(Ptrs out, int B, int T, int compute_iters, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    unsigned char* stage = smem + (size_t)warp * 16 * PITCH;
    float acc = (float)lane;
    const int nwarps_total = gridDim.x * 4;
    for (int tile = blockIdx.x * 4 + warp; tile * 16 < B; tile += nwarps_total) {
        const int row0 = tile * 16;
        for (int t = 0; t < T; ++t) {
            for (int it = 0; it < compute_iters; ++it) acc = acc * 1.0001f + 0.5f;
            const float4 v = make_float4(acc, acc + 1.f, acc + 2.f, acc + 3.f);
            if (MODE == 1) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
            int off = 0;
#pragma unroll
            for (int s = 0; s < NSEG; ++s) {
                const int sb = c_seg_bytes[s];
                for (int c = tq * 16; c < sb; c += 64) {  // 64-byte column blocks: lane owns 16 bytes of rows g and g + 8
                    if (MODE == 0) {
                        *reinterpret_cast<float4*>(out.t[s] + ((size_t)(row0 + g) * T + t) * sb + c) = v;
                        *reinterpret_cast<float4*>(out.t[s] + ((size_t)(row0 + g + 8) * T + t) * sb + c) = v;
                    } else if (MODE == 4) {
                        *reinterpret_cast<float4*>(out.t[0] + ((size_t)(row0 + g) * T + t) * ROW_BYTES + off + c) = v;
                        *reinterpret_cast<float4*>(out.t[0] + ((size_t)(row0 + g + 8) * T + t) * ROW_BYTES + off + c) = v;
                    } else if (MODE == 1 || MODE == 2) {
                        *reinterpret_cast<float4*>(stage + g * PITCH + off + c) = v;
                        *reinterpret_cast<float4*>(stage + (g + 8) * PITCH + off + c) = v;
                    }
                }
                off += sb;
            }
            if (MODE == 1) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                int o = 0;
#pragma unroll
                for (int s = 0; s < NSEG; ++s) {
                    const int sb = c_seg_bytes[s];
                    if (lane < 16)
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out.t[s] + ((size_t)(row0 + lane) * T + t) * sb),
                                     "r"(smem_u32(stage + lane * PITCH + o)), "r"(sb)
                                     : "memory");
                    o += sb;
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            } else if (MODE == 2) {
                __syncwarp();
                int o = 0;
#pragma unroll
                for (int s = 0; s < NSEG; ++s) {
                    const int sb = c_seg_bytes[s], c16 = sb / 16;
                    for (int i = lane; i < 16 * c16; i += 32) {
                        const int r = i / c16, c = i - r * c16;
                        *reinterpret_cast<float4*>(out.t[s] + ((size_t)(row0 + r) * T + t) * sb + c * 16) =
                            *reinterpret_cast<const float4*>(stage + r * PITCH + o + c * 16);
                    }
                    o += sb;
                }
                __syncwarp();
            }
        }
    }
    if (MODE == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (acc == 1.2345e-30f) *sink = acc;
}

template <int MODE>
static void run(Ptrs out, int B, int T, int ctas_per_sm, int ci, float* sink) {
    const size_t smem = (size_t)4 * 16 * PITCH;
    cudaFuncSetAttribute(store_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        store_kernel<MODE><<<148 * ctas_per_sm, 128, smem>>>(out, B, T, ci, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const char* names[] = {"fragment STG -> global      ", "smem + bulk store per row   ", "smem + coalesced STG        ", "no stores                   ",
                           "fragment STG -> ONE packed  "};
    printf("%s ctas/sm=%d compute=%4d: %.3f ms  %7.1f GB/s written  (%s)\n", names[MODE], ctas_per_sm, ci, best,
           MODE == 3 ? 0.0 : (double)B * T * ROW_BYTES / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int B = 37888, T = 30;
    const int seg_bytes[NSEG] = {384, 384, 128, 128, 64, 64, 64, 64};
    Ptrs out;
    for (int s = 0; s < NSEG; ++s) cudaMalloc(&out.t[s], (size_t)B * T * (s == 0 ? ROW_BYTES : seg_bytes[s]));
    float* sink;
    cudaMalloc(&sink, 4);
    for (int ci : {0, 1000, 2000})
        for (int cps : {1, 2}) {
            run<3>(out, B, T, cps, ci, sink);
            run<0>(out, B, T, cps, ci, sink);
            run<1>(out, B, T, cps, ci, sink);
            run<2>(out, B, T, cps, ci, sink);
            run<4>(out, B, T, cps, ci, sink);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
