// Epilogue helpers shared by the 1-CTA and 2-CTA tcgen05 GEMM kernels.
#pragma once

#include "hvlm_internal.cuh"
#include "hvlm_ptx.cuh"

namespace hvlm {

__device__ __forceinline__ float quick_gelu(float x) {
    // x * sigmoid(1.702 x)  (HF QuickGELUActivation); sigmoid(z) = 0.5 tanh(z/2) + 0.5 -> one MUFU op, no divide
    const float hx = 0.5f * x;
    return fmaf(hx, tanh_approx(0.851f * x), hx);
}

template <int EPI>
__host__ __device__ constexpr bool epi_is_staged() {
    return EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_F32 || EPI == EPI_GELU_BF16 || EPI == EPI_GELU_F32 ||
           EPI == EPI_RESID_F32 || EPI == EPI_QKV_HM;
}
template <int EPI>
__host__ __device__ constexpr bool epi_out_f32() {
    return EPI == EPI_BIAS_F32 || EPI == EPI_GELU_F32 || EPI == EPI_RESID_F32;
}

// acc (32 fp32 columns of this thread's row) -> +bias -> (quick-GELU) -> v
template <int EPI>
__device__ __forceinline__ void epilogue_math(const uint32_t (&acc)[32], const float* __restrict__ bias, int n0,
                                              float (&v)[32]) {
    if (bias != nullptr) {
        const float4* b4 = reinterpret_cast<const float4*>(bias + n0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            v[4 * j + 0] = __uint_as_float(acc[4 * j + 0]) + b.x;
            v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + b.y;
            v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + b.z;
            v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + b.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
    }
    if constexpr (EPI == EPI_GELU_BF16 || EPI == EPI_GELU_F32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
    }
}

// LayerNorm folded into the GEMM (see EpiArgs::ln_stats): v = rstd * acc + (bias[n] - mean * rstd * c[n]) -> (quick-GELU)
// b4 / c4: the 32 columns' folded bias and column sums, staged in SHARED memory by the caller (every thread of the group reads
// the same addresses: broadcast).  From global memory these warp-uniform loads cost an L2 round trip per 32-column chunk --
// a pair meets each column block only once or twice per launch -- which ncu showed as a quarter of the epilogue's time.
template <int EPI>
__device__ __forceinline__ void epilogue_math_fold(const uint32_t (&acc)[32], const float4* b4, const float4* c4, float rstd,
                                                   float nmr, float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 b = b4[j];
        const float4 c = c4[j];
        // packed fp32x2 FMAs: the epilogue warps are alone on their schedulers, instruction count is what they pay for
        float t0, t1, t2, t3;
        ffma2(t0, t1, nmr, nmr, c.x, c.y, b.x, b.y);
        ffma2(t2, t3, nmr, nmr, c.z, c.w, b.z, b.w);
        ffma2(v[4 * j + 0], v[4 * j + 1], __uint_as_float(acc[4 * j + 0]), __uint_as_float(acc[4 * j + 1]), rstd, rstd, t0, t1);
        ffma2(v[4 * j + 2], v[4 * j + 3], __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]), rstd, rstd, t2, t3);
    }
    if constexpr (EPI == EPI_GELU_BF16 || EPI == EPI_GELU_F32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
    }
}

// (mean, rstd) of a 1024-wide row from its eight (sum, sum of squares) partials; returns rstd, nmr = -mean * rstd and mean
__device__ __forceinline__ void fold_row_stats(const float4 (&st)[4], float eps, float& rstd, float& nmr, float& mean_out) {
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s += st[j].x + st[j].z;
        ss += st[j].y + st[j].w;
    }
    const float mean = s * (1.0f / 1024.0f);
    const float var = fmaxf(fmaf(-mean, mean, ss * (1.0f / 1024.0f)), 0.f);
    rstd = rsqrtf(var + eps);
    nmr = -mean * rstd;
    mean_out = mean;
}

}  // namespace hvlm
