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

}  // namespace hvlm
