// Internal (non-ABI) declarations shared by the translation units of libhvlm_b200.so.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hvlm_b200.h"

namespace hvlm {

// ---- GEMM epilogue selectors (internal superset of hvlm_epilogue) ----------------------------------
enum Epi : int {
    EPI_BIAS_BF16 = 0,   // out bf16 [M,N] = acc + bias
    EPI_BIAS_F32 = 1,    // out f32  [M,N] = acc + bias
    EPI_GELU_BF16 = 2,   // out bf16 [M,N] = quick_gelu(acc + bias)
    EPI_RESID_F32 = 3,   // out f32  [M,N] += acc + bias  (TMA reduce-add into the residual stream, in place)
    EPI_QKV_HM = 4,      // ViT: bf16 q|k|v written column-block-major [48][M][64] (3-D TMA stores)
    EPI_PATCH = 5,       // ViT: x0[f*257+1+p, n] = acc   (f32; the position embedding is added by the pre-LayerNorm)
    EPI_GELU_F32 = 6     // out f32 = quick_gelu(acc + bias)            (tests only)
};

struct EpiArgs {
    const float* bias = nullptr;
    const float* resid = nullptr;
    void* out = nullptr;
    const float* pos = nullptr;    // unused since the position embedding moved into the pre-LayerNorm kernel
    int reverse = 0;   // walk the tiles last-to-first (start with what the producer kernel left in L2)
    // fused LayerNorm of the updated residual rows (EPI_RESID_F32, N == 1024, 2-CTA kernel only): extra warps
    // normalise each 128-row block as soon as all of its column tiles have been reduced into `out`
    const float* ln_gamma = nullptr;
    const float* ln_beta = nullptr;
    void* ln_out = nullptr;        // bf16 [M,1024]
    int32_t* ln_count = nullptr;   // int32 [ceil(M/128)], zero on entry, zero again on exit
    // LayerNorm FOLDED into the GEMM that consumes it (2-CTA kernel; EPI_QKV_HM / EPI_BIAS_BF16 / EPI_GELU_BF16, K == 1024):
    //   LN(x) W^T + b  =  rstd_i * (bf16(x) (gamma*W)^T - mean_i * c) + (b + W beta),   c[n] = sum_k (gamma*W)[n,k]
    // A = bf16(x) (un-normalised rows), B = bf16(gamma*W), bias = b + W beta, and the epilogue applies the row's
    // (mean, rstd) computed from `ln_stats`.
    const float* ln_c = nullptr;       // f32 [N]
    const float* ln_stats = nullptr;   // f32 [M][8][2]: (sum, sum of squares) of x over each 128-column block of the row
    float ln_eps = 1e-5f;
    // the producer side of the fold (EPI_RESID_F32, N == 1024): the epilogue LOADS the residual tile, so it can also emit
    // the bf16 copy of the updated rows and their per-128-column statistics for the next folded GEMM
    void* xb_out = nullptr;            // bf16 [M,1024]
    float* stats_out = nullptr;        // f32 [M][8][2]
    // Row centring.  LayerNorm does not see a constant added to a row, so the producer may subtract ANY per-row value before
    // it rounds the row to bf16 and accumulates its statistics: xb = bf16(x - shift_i), stats over (x - shift_i), and the
    // consumer's formula is unchanged.  With shift_i ~ the row mean the bf16 rounding acts on the centred row (as in
    // LayerNorm-then-round) and the one-pass variance has nothing to cancel.  The consumer's n_blk == 0 tiles keep the
    // running mean up to date for the next producer: shift_io[i] += mean of the centred row.
    const float* shift_in = nullptr;   // producer: f32 [M] or null (= 0)
    float* shift_io = nullptr;         // consumer: f32 [M] or null
};

// C = A[M,K] * B[N,K]^T with the chosen epilogue.  Returns an hvlm_status.
int launch_gemm(int epi, const void* A, const void* B, int M, int N, int K, const EpiArgs& ep, cudaStream_t s);

// ---- TMA descriptor helper -----------------------------------------------------------------------
// 2D/3D bf16 tensor map with SWIZZLE_128B and a [box_rows x 64] (x1) box.  dims/strides innermost first.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);
// same for fp32 elements (box inner dimension 32 elements = 128 B)
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

int make_qkv_hm_tmap(CUtensorMap* out, const void* qkv_hm, int n_rows, int box_rows);
int num_sms();
int check_last(const char* what);
void count_launch();

// RAII per-stage timing (no-op unless hvlm_profile_enable(1))
struct StageTimer {
    StageTimer(int stage, cudaStream_t s);
    ~StageTimer();
    int idx_;
    cudaStream_t s_;
};

// Programmatic dependent launch (PDL): the kernel may be scheduled while its predecessor on the stream drains; it
// runs its prologue (barrier init, TMEM alloc, descriptor prefetch) and then blocks in pdl_wait() until the
// predecessor has completed and flushed.  Every kernel launched through this helper MUST call pdl_wait() before
// touching global memory.  On for small batches, off for large ones, HVLM_PDL=1/0 overrides (policy and numbers in profile.cu).
bool pdl_enabled(int cls = 1);   // cls: 0 = small streaming kernels (LayerNorm, im2col), 1 = GEMM, 2 = attention
// RAII hint for the launches issued by the current host thread while it is alive (see profile.cu)
struct PdlScope {
    explicit PdlScope(bool on);
    ~PdlScope();
    int prev_;
};
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cls(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                  Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled(cls) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    return launch_pdl_cls(1, kernel, grid, block, smem, s, args...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace hvlm
