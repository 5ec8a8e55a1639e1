// How much HBM bandwidth does the rollout's access pattern allow?  Each warp owns 16 rows of a [B, T, seg] tensor and marches
// over t, touching one `seg`-byte segment per row per step (stride T*seg between rows), exactly like the rollout kernels.
// Modes: copy (read + write), read-only, write-only; NT independent tensors are touched per step (the kernels touch ~10).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/src/seg_bw profiles/src/seg_bw.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

template <int MODE>  // 0 copy, 1 read, 2 write
__global__ void __launch_bounds__(128) seg_kernel(const float4* __restrict__ in, float4* __restrict__ out, int B, int T, int seg16, int TB,
                                                   float* sink) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int row0 = warp * 16;
    if (row0 >= B) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // per step block of TB steps: 16 rows x (TB*seg16) float4 per row, contiguous per row
    const int per_row = TB * seg16;
    for (int t0 = 0; t0 < T; t0 += TB) {
        for (int i = lane; i < 16 * per_row; i += 32) {
            const int r = i / per_row, c = i - r * per_row;
            const size_t idx = ((size_t)(row0 + r) * T + t0) * seg16 + c;
            if (MODE == 0) {
                out[idx] = in[idx];
            } else if (MODE == 1) {
                const float4 v = in[idx];
                acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
            } else {
                out[idx] = make_float4((float)i, 1.f, 2.f, 3.f);
            }
        }
    }
    if (MODE == 1 && acc.x + acc.y + acc.z + acc.w == 1.2345e-30f) *sink = acc.x;
}

int main() {
    const size_t total = (size_t)3 << 30;  // bytes per tensor
    float4 *in, *out;
    float* sink;
    cudaMalloc(&in, total), cudaMalloc(&out, total), cudaMalloc(&sink, 4);
    cudaMemset(in, 1, total), cudaMemset(out, 0, total);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const int segs[] = {64, 128, 256, 384, 512, 1024, 2048, 4096};
    printf("pattern: warp = 16 rows, one seg per row per step, row stride T*seg; B = 37888 (2368 warps), TB = steps fetched at once\n");
    for (int mode = 0; mode < 3; ++mode)
        for (int seg : segs)
            for (int TB : {1, 2, 4, 8}) {
                const int B = 37888;
                int T = (int)(total / ((size_t)B * seg));
                T = T / 8 * 8;
                if (T > 240) T = 240;
                const int seg16 = seg / 16;
                const int warps = B / 16, blocks = (warps + 3) / 4;
                float best = 1e30f;
                for (int rep = 0; rep < 3; ++rep) {
                    cudaEventRecord(e0);
                    if (mode == 0) seg_kernel<0><<<blocks, 128>>>(in, out, B, T, seg16, TB, sink);
                    if (mode == 1) seg_kernel<1><<<blocks, 128>>>(in, out, B, T, seg16, TB, sink);
                    if (mode == 2) seg_kernel<2><<<blocks, 128>>>(in, out, B, T, seg16, TB, sink);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (ms < best) best = ms;
                }
                const double bytes = (double)B * T * seg * (mode == 0 ? 2 : 1);
                printf("%s seg=%4d TB=%d T=%3d: %7.1f GB/s\n", mode == 0 ? "copy " : mode == 1 ? "read " : "write", seg, TB, T, bytes / best / 1e6);
            }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
