// Fused Gaussian reconstruction likelihood.
//
// Reference: models/objective.py:7-23 -- `-Independent(Normal(prediction, scale), event_ndims).log_prob(target).mean()`,
// called once per modality by `compute_reconstruction_loss` (mrssm/mopoe_mrssm/core.py:294-303) on the decoder outputs
// [B,T,1,32,32]: the largest tensors of a training step (4 KB per (b,t) per modality).
//
// Closed form: loss = 0.5 / (scale^2 * n_batch) * sum_i (target_i - prediction_i)^2 + n_event * (log(scale) + 0.5 log(2 pi)).
// Both directions are pure streaming passes, bound by HBM: forward reads prediction + target (8 B per element for fp32
// predictions), backward reads both and writes d prediction (12 B per element).  Up to RSSM_NLL_MAX_SEGMENTS (prediction,
// target) pairs -- the modalities -- share ONE launch (blockIdx.y = pair).
//
// Forward: grid-stride over 16-byte vectors, UNROLL loads of each tensor issued before the first use (2 * UNROLL LDG.128 in
// flight per thread, a CTA touches contiguous 16 KB runs), fp32 partial per thread (4 independent accumulators), fp64 from
// the warp reduction on.  Each CTA writes one fp64 partial; the last CTA to arrive (ticket) adds the partials in index order,
// so the result does not depend on the arrival order (bit-reproducible for a given grid).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"

namespace rssm {
namespace {

constexpr int NLL_THREADS = 256;
constexpr int NLL_UNROLL = 4;

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ld_stream_u2(const uint2* p) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

// four consecutive predictions as fp32
template <typename PT>
struct Vec4;
template <>
struct Vec4<float> {
    using raw = float4;
    static __device__ __forceinline__ raw load(const void* base, size_t i) { return ld_stream_f4(static_cast<const float4*>(base) + i); }
    static __device__ __forceinline__ float4 widen(raw r) { return r; }
    static __device__ __forceinline__ void store(void* base, size_t i, float4 v) { static_cast<float4*>(base)[i] = v; }
    static __device__ __forceinline__ float scalar(const void* base, size_t i) { return static_cast<const float*>(base)[i]; }
    static __device__ __forceinline__ void store_scalar(void* base, size_t i, float v) { static_cast<float*>(base)[i] = v; }
};
template <>
struct Vec4<__nv_bfloat16> {
    using raw = uint2;
    static __device__ __forceinline__ raw load(const void* base, size_t i) { return ld_stream_u2(static_cast<const uint2*>(base) + i); }
    static __device__ __forceinline__ float4 widen(raw r) {
        // bf16 -> fp32 is a 16-bit shift
        return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                           __uint_as_float(r.y & 0xffff0000u));
    }
    static __device__ __forceinline__ void store(void* base, size_t i, float4 v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 r;
        r.x = *reinterpret_cast<uint32_t*>(&a), r.y = *reinterpret_cast<uint32_t*>(&b);
        static_cast<uint2*>(base)[i] = r;
    }
    static __device__ __forceinline__ float scalar(const void* base, size_t i) {
        return __bfloat162float(static_cast<const __nv_bfloat16*>(base)[i]);
    }
    static __device__ __forceinline__ void store_scalar(void* base, size_t i, float v) {
        static_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v);
    }
};
template <>
struct Vec4<__half> {
    using raw = uint2;
    static __device__ __forceinline__ raw load(const void* base, size_t i) { return ld_stream_u2(static_cast<const uint2*>(base) + i); }
    static __device__ __forceinline__ float4 widen(raw r) {
        float2 a = __half22float2(*reinterpret_cast<__half2*>(&r.x)), b = __half22float2(*reinterpret_cast<__half2*>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ void store(void* base, size_t i, float4 v) {
        __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 r;
        r.x = *reinterpret_cast<uint32_t*>(&a), r.y = *reinterpret_cast<uint32_t*>(&b);
        static_cast<uint2*>(base)[i] = r;
    }
    static __device__ __forceinline__ float scalar(const void* base, size_t i) { return __half2float(static_cast<const __half*>(base)[i]); }
    static __device__ __forceinline__ void store_scalar(void* base, size_t i, float v) { static_cast<__half*>(base)[i] = __float2half_rn(v); }
};

template <typename PT>
__global__ void __launch_bounds__(NLL_THREADS, 4) gaussian_nll_fwd_kernel(const NllArgs a) {
    using V = Vec4<PT>;
    const NllSeg s = a.seg[blockIdx.y];
    const size_t n4 = s.n / 4;
    const float4* tgt = reinterpret_cast<const float4*>(s.target);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const size_t span = (size_t)NLL_THREADS * NLL_UNROLL;
    for (size_t base = (size_t)blockIdx.x * span + threadIdx.x; base < n4; base += (size_t)gridDim.x * span) {
        typename V::raw p[NLL_UNROLL];
        float4 t[NLL_UNROLL];
        bool ok[NLL_UNROLL];
#pragma unroll
        for (int u = 0; u < NLL_UNROLL; ++u) {
            const size_t i = base + (size_t)u * NLL_THREADS;
            ok[u] = i < n4;
            if (ok[u]) p[u] = V::load(s.prediction, i), t[u] = ld_stream_f4(tgt + i);
        }
#pragma unroll
        for (int u = 0; u < NLL_UNROLL; ++u) {
            if (!ok[u]) continue;
            const float4 q = V::widen(p[u]);
            const float dx = t[u].x - q.x, dy = t[u].y - q.y, dz = t[u].z - q.z, dw = t[u].w - q.w;
            acc[0] = fmaf(dx, dx, acc[0]), acc[1] = fmaf(dy, dy, acc[1]), acc[2] = fmaf(dz, dz, acc[2]), acc[3] = fmaf(dw, dw, acc[3]);
        }
    }
    // ragged tail (n % 4 elements), first CTA of the pair
    if (blockIdx.x == 0 && threadIdx.x < (s.n & 3)) {
        const size_t i = n4 * 4 + threadIdx.x;
        const float d = s.target[i] - V::scalar(s.prediction, i);
        acc[0] = fmaf(d, d, acc[0]);
    }
    double sum = ((double)acc[0] + (double)acc[1]) + ((double)acc[2] + (double)acc[3]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __shared__ double warp_sum[NLL_THREADS / 32];
    __shared__ bool last;
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = sum;
    __syncthreads();
    double* partials = a.partials + (size_t)blockIdx.y * gridDim.x;
    if (threadIdx.x == 0) {
        double cta = 0.0;
#pragma unroll
        for (int w = 0; w < NLL_THREADS / 32; ++w) cta += warp_sum[w];
        partials[blockIdx.x] = cta;
        __threadfence();
        last = atomicAdd(a.tickets + blockIdx.y, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last || threadIdx.x >= 32) return;
    __threadfence();
    // fixed-order final sum: lane l adds partials l, l+32, ...; then a fixed butterfly
    double tot = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) tot += __ldcg(partials + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (threadIdx.x == 0) {
        *s.loss = (float)(tot * s.sq_coeff + s.constant);
        a.tickets[blockIdx.y] = 0;  // the workspace is reusable without a memset
    }
}

template <typename PT>
__global__ void __launch_bounds__(NLL_THREADS, 4) gaussian_nll_bwd_kernel(const NllArgs a) {
    using V = Vec4<PT>;
    const NllSeg s = a.seg[blockIdx.y];
    const size_t n4 = s.n / 4;
    const float4* tgt = reinterpret_cast<const float4*>(s.target);
    // d loss / d prediction_i = 2 * sq_coeff * (prediction_i - target_i) * upstream
    const float g = 2.f * (float)s.sq_coeff * (s.d_loss ? __ldg(s.d_loss) : 1.f);
    const size_t span = (size_t)NLL_THREADS * NLL_UNROLL;
    for (size_t base = (size_t)blockIdx.x * span + threadIdx.x; base < n4; base += (size_t)gridDim.x * span) {
        typename V::raw p[NLL_UNROLL];
        float4 t[NLL_UNROLL];
        bool ok[NLL_UNROLL];
#pragma unroll
        for (int u = 0; u < NLL_UNROLL; ++u) {
            const size_t i = base + (size_t)u * NLL_THREADS;
            ok[u] = i < n4;
            if (ok[u]) p[u] = V::load(s.prediction, i), t[u] = ld_stream_f4(tgt + i);
        }
#pragma unroll
        for (int u = 0; u < NLL_UNROLL; ++u) {
            if (!ok[u]) continue;
            const size_t i = base + (size_t)u * NLL_THREADS;
            const float4 q = V::widen(p[u]);
            const float4 d = make_float4(g * (q.x - t[u].x), g * (q.y - t[u].y), g * (q.z - t[u].z), g * (q.w - t[u].w));
            V::store(s.d_prediction, i, d);
            if (s.d_target) reinterpret_cast<float4*>(s.d_target)[i] = make_float4(-d.x, -d.y, -d.z, -d.w);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (s.n & 3)) {
        const size_t i = n4 * 4 + threadIdx.x;
        const float d = g * (V::scalar(s.prediction, i) - s.target[i]);
        V::store_scalar(s.d_prediction, i, d);
        if (s.d_target) s.d_target[i] = -d;
    }
}

template <typename PT>
cudaError_t launch(const NllArgs& a, bool backward, int ctas, cudaStream_t st) {
    dim3 grid(ctas, a.nseg);
    if (backward)
        gaussian_nll_bwd_kernel<PT><<<grid, NLL_THREADS, 0, st>>>(a);
    else
        gaussian_nll_fwd_kernel<PT><<<grid, NLL_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace

int nll_ctas_per_segment(int nseg) {
    // 4 resident CTAs of 256 threads per SM (<= 64 registers: __launch_bounds__), whole waves over the 148 SMs, split over the pairs
    int per = (148 * 4) / (nseg < 1 ? 1 : nseg);
    return per < 1 ? 1 : per;
}

cudaError_t launch_gaussian_nll(const NllArgs& a, bool backward, cudaStream_t s) {
    size_t nmax = 0;
    for (int i = 0; i < a.nseg; ++i) nmax = a.seg[i].n > nmax ? a.seg[i].n : nmax;
    // no more CTAs than 16 KB runs of the largest pair; never more than the workspace was sized for
    size_t runs = (nmax / 4 + (size_t)NLL_THREADS * NLL_UNROLL - 1) / ((size_t)NLL_THREADS * NLL_UNROLL);
    int ctas = nll_ctas_per_segment(a.nseg);
    if ((size_t)ctas > runs) ctas = runs < 1 ? 1 : (int)runs;
    switch (a.pred_dtype) {
        case 0: return launch<float>(a, backward, ctas, s);
        case 1: return launch<__nv_bfloat16>(a, backward, ctas, s);
        case 2: return launch<__half>(a, backward, ctas, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace rssm
