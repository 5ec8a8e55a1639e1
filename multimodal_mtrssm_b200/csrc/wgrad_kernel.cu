// Batched weight gradients for the rollout: for every layer ("job")
//     dW[n][k] += sum over (b,t) of dY[(b,t)][n] * X[(b,t) or (b,t-1)][k],   db[n] += sum dY[(b,t)][n]
// where dY is a slice of the backward kernel's per-(b,t) "dpre" record and X a slice of the forward
// kernel's outputs / saved record / inputs.  The recurrence carries only data gradients, so this pass
// is embarrassingly parallel over (b,t): grid = (row chunks, jobs).
//
// v1: fp32 FFMA, 4x4 register micro-tiles, operands streamed through L1 (each CTA re-reads a row only
// from L1), one atomicAdd per output element per CTA.  Accumulates into dW/db (caller zero-fills).
#include <stdint.h>

#include "kernels.h"

namespace rssm {

__global__ void __launch_bounds__(256) wgrad_kernel(const WgradArgs a) {
    const WgradJob& j = a.jobs[blockIdx.y];
    const int tiles_k = (j.K + 3) >> 2, tiles_n = (j.N + 3) >> 2, ntiles = tiles_n * tiles_k;
    const int nslices = 256 / ntiles > 0 ? 256 / ntiles : 1;
    const int tid = threadIdx.x;
    const int R = a.B * a.T, T = a.T;
    const int per = (R + gridDim.x - 1) / gridDim.x;
    const int c0 = blockIdx.x * per, c1 = min(R, c0 + per);

    const int slice = tid / ntiles, tile = tid % ntiles;
    if (slice >= nslices) return;  // no block-level synchronisation below
    {
        const int tn = tile / tiles_k, tk = tile % tiles_k;
        const int n0 = tn * 4, k0 = tk * 4;
        const bool vec = (j.N % 4 == 0) && (j.K % 4 == 0) && (j.ldy % 4 == 0) && (j.ldx % 4 == 0) && (j.ldx0 % 4 == 0) &&
                         ((reinterpret_cast<uintptr_t>(j.dY) | reinterpret_cast<uintptr_t>(j.X) |
                           reinterpret_cast<uintptr_t>(j.X0)) % 16 == 0);
        float acc[4][4] = {};
        float bsum[4] = {};
        for (int row = c0 + slice; row < c1; row += nslices) {
            const float* dy = j.dY + (size_t)row * j.ldy + n0;
            const float* x;
            if (j.shift) {
                const int b = row / T, t = row - b * T;
                x = (t > 0 ? j.X + (size_t)(row - 1) * j.ldx : j.X0 + (size_t)b * j.ldx0) + k0;
            } else {
                x = j.X + (size_t)row * j.ldx + k0;
            }
            float d[4], v[4];
            if (vec) {
                const float4 d4 = *reinterpret_cast<const float4*>(dy), v4 = *reinterpret_cast<const float4*>(x);
                d[0] = d4.x, d[1] = d4.y, d[2] = d4.z, d[3] = d4.w;
                v[0] = v4.x, v[1] = v4.y, v[2] = v4.z, v[3] = v4.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    d[i] = n0 + i < j.N ? dy[i] : 0.f;
                    v[i] = k0 + i < j.K ? x[i] : 0.f;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                bsum[i] += d[i];
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[i][k] = fmaf(d[i], v[k], acc[i][k]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (n0 + i >= j.N) continue;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k0 + k < j.K) atomicAdd(j.dW + (size_t)(n0 + i) * j.ldw + k0 + k, acc[i][k]);
            if (j.db != nullptr && tk == 0) atomicAdd(j.db + n0 + i, bsum[i]);
        }
    }
}

cudaError_t launch_wgrad(const WgradArgs& a, cudaStream_t s) {
    if (a.njobs <= 0) return cudaSuccess;
    const int R = a.B * a.T;
    int chunks = (R + 255) / 256;
    if (chunks > 74) chunks = 74;
    if (chunks < 1) chunks = 1;
    wgrad_kernel<<<dim3(chunks, a.njobs), 256, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace rssm
