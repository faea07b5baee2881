// Bandwidth-bound pieces of the ViT-L/14 tower and of the training-shaped variant:
//   * LayerNorm over 1024 channels (fp32 residual stream in; bf16 GEMM operand or fp32 out)
//   * patch extraction (im2col) for the 14x14/stride-14 convolution + CLS row initialisation
//   * feature_select (drop CLS, cast)            clip_encoder.py:29-37,49
//   * transpose-to-bf16 and column sums for the projector wgrad / bias grad
#include "hvlm_internal.cuh"
#include "hvlm_ptx.cuh"
#include "hvlm_vec.cuh"

namespace hvlm {

// ------------------------------------------------------------------------------------------------
// LayerNorm(1024), eps inside the sqrt, biased variance (torch.nn.LayerNorm).  One warp per row: the row
// (4 KB fp32) lives in registers, two-pass mean / variance in fp32 with warp shuffles.
// ------------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void __launch_bounds__(256) layernorm1024_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, TOut* __restrict__ out,
                                                            int rows, float eps, int reverse,
                                                            const float* __restrict__ add_rows, int add_period,
                                                            __nv_bfloat16* __restrict__ xb_out,
                                                            float* __restrict__ stats_out, float* __restrict__ shift_out) {
    // the dependent launch is triggered at the END of this kernel: its successor is a GEMM whose CTAs would otherwise become
    // resident next to the LayerNorm CTAs right away (no shared memory to wait for) and halve their occupancy
    pdl_wait();
    int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) {
        pdl_launch_dependents();
        return;
    }
    if (reverse) row = rows - 1 - row;   // start with the rows the producer kernel touched last (still in L2)
    const int lane = threadIdx.x & 31;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * 1024);
    float4 v[8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = xr[lane + 32 * j];
    if (add_rows != nullptr) {
        // pre-LayerNorm of the tower: the patch-embedding GEMM stored the bare convolution output, the position
        // embedding of token t = row % 257 is added here with coalesced reads (token 0 = CLS already carries pos[0])
        const int t = row % add_period;
        if (t > 0) {
            const float4* pr = reinterpret_cast<const float4*>(add_rows + static_cast<size_t>(t) * 1024);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 e = __ldg(pr + lane + 32 * j);
                v[j].x += e.x;
                v[j].y += e.y;
                v[j].z += e.z;
                v[j].w += e.w;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / 1024.0f);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.0f / 1024.0f) + eps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    [[maybe_unused]] float ys = 0.f, yss = 0.f;   // (sum, sum of squares) of the centred OUTPUT row (folded-LayerNorm feed)
    [[maybe_unused]] float ymean = 0.f;           // the output row's mean = the centre the bf16 copy and the statistics use
    if constexpr (sizeof(TOut) == 4) {
        if (xb_out != nullptr) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 g = __ldg(g4 + lane + 32 * j);
                const float4 b = __ldg(b4 + lane + 32 * j);
                t += ((v[j].x - mean) * rstd * g.x + b.x) + ((v[j].y - mean) * rstd * g.y + b.y) +
                     ((v[j].z - mean) * rstd * g.z + b.z) + ((v[j].w - mean) * rstd * g.w + b.w);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            ymean = t * (1.0f / 1024.0f);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 g = __ldg(g4 + lane + 32 * j);
        const float4 b = __ldg(b4 + lane + 32 * j);
        float y[4];
        y[0] = (v[j].x - mean) * rstd * g.x + b.x;
        y[1] = (v[j].y - mean) * rstd * g.y + b.y;
        y[2] = (v[j].z - mean) * rstd * g.z + b.z;
        y[3] = (v[j].w - mean) * rstd * g.w + b.w;
        TOut* o = out + static_cast<size_t>(row) * 1024 + (lane + 32 * j) * 4;
        if constexpr (sizeof(TOut) == 4) {
            *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
            if (xb_out != nullptr) {
                // the tower's pre-LayerNorm feeds a GEMM with the next LayerNorm folded in: bf16 copy of the CENTRED row + its
                // statistics in the same [8][2] layout the residual GEMMs write (block 0 carries the whole row)
                const float d0 = y[0] - ymean, d1 = y[1] - ymean, d2 = y[2] - ymean, d3 = y[3] - ymean;
                ys += (d0 + d1) + (d2 + d3);
                yss = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, yss))));
                __nv_bfloat162 lo = __floats2bfloat162_rn(d0, d1);
                __nv_bfloat162 hi = __floats2bfloat162_rn(d2, d3);
                uint2 w;
                w.x = *reinterpret_cast<uint32_t*>(&lo);
                w.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(xb_out + static_cast<size_t>(row) * 1024 + (lane + 32 * j) * 4) = w;
            }
        } else {
            __nv_bfloat162 lo = __floats2bfloat162_rn(y[0], y[1]);
            __nv_bfloat162 hi = __floats2bfloat162_rn(y[2], y[3]);
            uint2 w;
            w.x = *reinterpret_cast<uint32_t*>(&lo);
            w.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(o) = w;
        }
    }
    if constexpr (sizeof(TOut) == 4) {
        if (xb_out != nullptr) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ys += __shfl_xor_sync(0xffffffffu, ys, o);
                yss += __shfl_xor_sync(0xffffffffu, yss, o);
            }
            if (lane < 8)
                *reinterpret_cast<float2*>(stats_out + static_cast<size_t>(row) * 16 + lane * 2) =
                    lane == 0 ? make_float2(ys, yss) : make_float2(0.f, 0.f);
            if (lane == 0 && shift_out != nullptr) shift_out[row] = ymean;
        }
    }
    pdl_launch_dependents();
}

int launch_layernorm(const float* x, const float* g, const float* b, void* out, int rows, int out_dtype, float eps,
                     cudaStream_t s, int reverse, const float* add_rows, int add_period, void* xb_out, float* stats_out,
                     float* shift_out) {
    const int grid = (rows + 7) / 8;
    if ((xb_out != nullptr) != (stats_out != nullptr) || (xb_out != nullptr && out_dtype != HVLM_F32)) return HVLM_ERR_BAD_ARG;
    if (out_dtype == HVLM_F32)
        launch_pdl_cls(0, layernorm1024_kernel<float>, dim3(grid), dim3(256), 0, s, x, g, b, static_cast<float*>(out), rows, eps, reverse, add_rows,
                   add_period, static_cast<__nv_bfloat16*>(xb_out), stats_out, shift_out);
    else if (out_dtype == HVLM_BF16)
        launch_pdl_cls(0, layernorm1024_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, x, g, b, static_cast<__nv_bfloat16*>(out), rows, eps,
                   reverse, add_rows, add_period, static_cast<__nv_bfloat16*>(nullptr), static_cast<float*>(nullptr),
                   static_cast<float*>(nullptr));
    else
        return HVLM_ERR_BAD_DTYPE;
    return check_last("layernorm");
}

// ------------------------------------------------------------------------------------------------
// im2col for Conv2d(3,1024,k=14,s=14,bias=False): CTA = (patch row gy, frame).  The [3,14,224] pixel slab is
// staged in shared memory with coalesced reads, then the 16 patch rows [640] (588 + zero pad) are written
// contiguously as bf16.  The gy==0 CTA also writes the CLS row  x0[f,0,:] = class_embedding + pos[0].
// Column order (c, i, j) == conv.weight.reshape(1024, -1).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

template <typename TPix>
__global__ void __launch_bounds__(256) im2col_kernel(const TPix* __restrict__ pix, __nv_bfloat16* __restrict__ A,
                                                     const float* __restrict__ cls, const float* __restrict__ pos0,
                                                     float* __restrict__ x0) {
    __shared__ float slab[3 * 14 * 224];
    pdl_launch_dependents();
    pdl_wait();
    const int gy = blockIdx.x, f = blockIdx.y;
    const TPix* base = pix + static_cast<size_t>(f) * 3 * 224 * 224;
    // the 14 image rows of one channel are one contiguous run of 3136 pixels: 16-byte vector loads
    constexpr int V = Vec16<TPix>::N;
    constexpr int kRun = 14 * 224;
    for (int v = threadIdx.x; v < 3 * kRun / V; v += 256) {
        const int c = v / (kRun / V);
        const int off = (v - c * (kRun / V)) * V;
        float fv[V];
        unpack16<TPix>(ld_stream16(base + static_cast<size_t>(c) * 224 * 224 + gy * kRun + off), fv);
#pragma unroll
        for (int e = 0; e < V; ++e) slab[c * kRun + off + e] = fv[e];
    }
    __syncthreads();
    __nv_bfloat16* dst = A + (static_cast<size_t>(f) * 256 + gy * 16) * HVLM_VIT_PATCH_KPAD;
    // one 16-byte store = 8 consecutive columns (c, i, j) of one patch row
    for (int v = threadIdx.x; v < 16 * (HVLM_VIT_PATCH_KPAD / 8); v += 256) {
        const int gx = v / (HVLM_VIT_PATCH_KPAD / 8);
        const int col0 = (v - gx * (HVLM_VIT_PATCH_KPAD / 8)) * 8;
        int c = col0 / 196, r = col0 - c * 196;
        int i = r / 14, j = r - i * 14;
        float o8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            o8[e] = (col0 + e < HVLM_VIT_PATCH_K) ? slab[c * kRun + i * 224 + gx * 14 + j] : 0.f;
            if (++j == 14) {
                j = 0;
                if (++i == 14) {
                    i = 0;
                    ++c;
                }
            }
        }
        uint4 w;
        w.x = pack_bf16x2(o8[0], o8[1]);
        w.y = pack_bf16x2(o8[2], o8[3]);
        w.z = pack_bf16x2(o8[4], o8[5]);
        w.w = pack_bf16x2(o8[6], o8[7]);
        *reinterpret_cast<uint4*>(dst + static_cast<size_t>(gx) * HVLM_VIT_PATCH_KPAD + col0) = w;
    }
    if (gy == 0) {
        float* o = x0 + static_cast<size_t>(f) * HVLM_VIT_TOKENS * 1024;
        for (int c = threadIdx.x; c < 1024; c += 256) o[c] = cls[c] + pos0[c];
    }
}

// uint8 NHWC frames (what a video decoder / PIL hands over) with the CLIPImageProcessor rescale + normalise fused in:
// value = (u8 / 255 - mean[c]) / std[c].  Same slab staging and output layout as im2col_kernel.
struct PixelNorm {
    float scale[3];   // 1 / (255 * std[c])
    float shift[3];   // -mean[c] / std[c]
};

__global__ void __launch_bounds__(256) im2col_u8_kernel(const uint8_t* __restrict__ pix, __nv_bfloat16* __restrict__ A,
                                                        const float* __restrict__ cls, const float* __restrict__ pos0,
                                                        float* __restrict__ x0, PixelNorm nrm) {
    __shared__ float slab[3 * 14 * 224];
    pdl_launch_dependents();
    pdl_wait();
    const int gy = blockIdx.x, f = blockIdx.y;
    // rows gy*14 .. gy*14+13 of frame f are one contiguous run of 14*224*3 bytes in NHWC
    const uint8_t* base = pix + (static_cast<size_t>(f) * 224 + gy * 14) * 224 * 3;
    const uint32_t* base4 = reinterpret_cast<const uint32_t*>(base);   // 14*224*3 = 9408 bytes, 4-byte aligned
    for (int w = threadIdx.x; w < 9408 / 4; w += 256) {
        const uint32_t v = base4[w];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = 4 * w + k;            // (i*224 + x)*3 + c
            const int c = idx % 3;
            const int ix = idx / 3;               // i*224 + x
            slab[c * 14 * 224 + ix] = static_cast<float>((v >> (8 * k)) & 0xFFu) * nrm.scale[c] + nrm.shift[c];
        }
    }
    __syncthreads();
    __nv_bfloat16* dst = A + (static_cast<size_t>(f) * 256 + gy * 16) * HVLM_VIT_PATCH_KPAD;
    for (int idx = threadIdx.x; idx < 16 * (HVLM_VIT_PATCH_KPAD / 2); idx += 256) {
        const int gx = idx / (HVLM_VIT_PATCH_KPAD / 2);
        const int col = (idx - gx * (HVLM_VIT_PATCH_KPAD / 2)) * 2;
        float v0 = 0.f, v1 = 0.f;
        if (col < HVLM_VIT_PATCH_K) {
            int c = col / 196, r = col - c * 196;
            int i = r / 14, j = r - i * 14;
            v0 = slab[c * 14 * 224 + i * 224 + gx * 14 + j];
            const int col1 = col + 1;
            c = col1 / 196;
            r = col1 - c * 196;
            i = r / 14;
            j = r - i * 14;
            v1 = slab[c * 14 * 224 + i * 224 + gx * 14 + j];
        }
        *reinterpret_cast<__nv_bfloat162*>(dst + static_cast<size_t>(gx) * HVLM_VIT_PATCH_KPAD + col) =
            __floats2bfloat162_rn(v0, v1);
    }
    if (gy == 0) {
        float* o = x0 + static_cast<size_t>(f) * HVLM_VIT_TOKENS * 1024;
        for (int c = threadIdx.x; c < 1024; c += 256) o[c] = cls[c] + pos0[c];
    }
}

int launch_im2col_u8(const uint8_t* frames, const float* mean, const float* stdv, int n_frames, void* A, const float* cls,
                     const float* pos, float* x0, cudaStream_t s) {
    PixelNorm nrm;
    for (int c = 0; c < 3; ++c) {
        nrm.scale[c] = 1.0f / (255.0f * stdv[c]);
        nrm.shift[c] = -mean[c] / stdv[c];
    }
    launch_pdl_cls(0, im2col_u8_kernel, dim3(16, n_frames), dim3(256), 0, s, frames, static_cast<__nv_bfloat16*>(A), cls, pos, x0, nrm);
    return check_last("im2col_u8");
}

int launch_im2col(const void* pixels, int pix_dtype, int n_frames, void* A, const float* cls, const float* pos,
                  float* x0, cudaStream_t s) {
    dim3 grid(16, n_frames);
    HVLM_DISPATCH_DTYPE(pix_dtype, TT, {
        launch_pdl_cls(0, im2col_kernel<TT>, grid, dim3(256), 0, s, static_cast<const TT*>(pixels), static_cast<__nv_bfloat16*>(A), cls, pos, x0);
    });
    return check_last("im2col");
}

// ------------------------------------------------------------------------------------------------
// feature_select: hidden f32 [n,257,1024] -> [n,256(+1),1024] in out dtype
// ------------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void __launch_bounds__(256) feature_select_kernel(const float* __restrict__ hidden, TOut* __restrict__ out,
                                                             int rows_out, int keep_cls) {
    const int row = blockIdx.x;   // output row
    const int per = keep_cls ? 257 : 256;
    const int f = row / per, t = row - f * per;
    const float4 v = reinterpret_cast<const float4*>(hidden + (static_cast<size_t>(f) * 257 + t + (keep_cls ? 0 : 1)) * 1024)[threadIdx.x];
    TOut* o = out + static_cast<size_t>(row) * 1024 + threadIdx.x * 4;
    if constexpr (sizeof(TOut) == 4) {
        *reinterpret_cast<float4*>(o) = v;
    } else {
        o[0] = from_float<TOut>(v.x);
        o[1] = from_float<TOut>(v.y);
        o[2] = from_float<TOut>(v.z);
        o[3] = from_float<TOut>(v.w);
    }
}

// ------------------------------------------------------------------------------------------------
// out bf16 [C, R_pad] = in[R, C]^T, zero padded; 32x32 smem tile transpose
// ------------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(256) transpose_to_bf16_kernel(const TIn* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                                int R, int C, int R_pad) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + tx;
        tile[ty + 8 * i][tx] = (r < R && c < C) ? to_float<TIn>(in[static_cast<size_t>(r) * C + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, r = r0 + tx;
        if (c < C && r < R_pad) out[static_cast<size_t>(c) * R_pad + r] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
    }
}

// db[n] = sum_m dy[m,n]: CTA = 32 columns x 8 row-lanes, then smem reduce
template <typename TIn>
__global__ void __launch_bounds__(256) colsum_kernel(const TIn* __restrict__ dy, float* __restrict__ db, int M, int N) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx;
    float acc = 0.f;
    if (n < N)
        for (int m = ty; m < M; m += 8) acc += to_float<TIn>(dy[static_cast<size_t>(m) * N + n]);
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][tx];
        db[n] = s;
    }
}

}  // namespace hvlm

extern "C" int hvlm_layernorm_1024(const float* x, const float* gamma, const float* beta, void* out, int rows,
                                   int out_dtype, float eps, void* stream) {
    using namespace hvlm;
    if (!x || !gamma || !beta || !out || rows <= 0) return HVLM_ERR_BAD_ARG;
    if (!aligned16(x) || !aligned16(gamma) || !aligned16(beta) || !aligned16(out)) return HVLM_ERR_ALIGN;
    return launch_layernorm(x, gamma, beta, out, rows, out_dtype, eps, static_cast<cudaStream_t>(stream), 0, nullptr, 1,
                            nullptr, nullptr, nullptr);
}

extern "C" int hvlm_layernorm_1024_stats(const float* x, const float* gamma, const float* beta, float* out, void* xb_out,
                                         float* stats_out, float* shift_out, int rows, float eps, void* stream) {
    using namespace hvlm;
    if (!x || !gamma || !beta || !out || !xb_out || !stats_out || !shift_out || rows <= 0) return HVLM_ERR_BAD_ARG;
    if (!aligned16(x) || !aligned16(gamma) || !aligned16(beta) || !aligned16(out) || !aligned16(xb_out) || !aligned16(stats_out))
        return HVLM_ERR_ALIGN;
    return launch_layernorm(x, gamma, beta, out, rows, HVLM_F32, eps, static_cast<cudaStream_t>(stream), 0, nullptr, 1, xb_out,
                            stats_out, shift_out);
}

extern "C" int hvlm_feature_select(const float* hidden, void* feats, int n_frames, int out_dtype, int keep_cls,
                                   void* stream) {
    using namespace hvlm;
    if (!hidden || !feats || n_frames <= 0) return HVLM_ERR_BAD_ARG;
    if (!aligned16(hidden) || !aligned16(feats)) return HVLM_ERR_ALIGN;
    const int rows = n_frames * (keep_cls ? 257 : 256);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HVLM_DISPATCH_DTYPE(out_dtype, TT, {
        feature_select_kernel<TT><<<rows, 256, 0, s>>>(hidden, static_cast<TT*>(feats), rows, keep_cls);
    });
    return check_last("feature_select");
}

extern "C" int hvlm_transpose_to_bf16(const void* in, int in_dtype, void* out, int R, int C, int R_pad, void* stream) {
    using namespace hvlm;
    if (!in || !out || R <= 0 || C <= 0 || R_pad < R || (R_pad % 8) != 0) return HVLM_ERR_BAD_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    dim3 grid((R_pad + 31) / 32, (C + 31) / 32);
    HVLM_DISPATCH_DTYPE(in_dtype, TT, {
        transpose_to_bf16_kernel<TT><<<grid, 256, 0, s>>>(static_cast<const TT*>(in), static_cast<__nv_bfloat16*>(out), R, C, R_pad);
    });
    return check_last("transpose");
}

extern "C" int hvlm_colsum(const void* dy, int dtype, float* db, int M, int N, void* stream) {
    using namespace hvlm;
    if (!dy || !db || M <= 0 || N <= 0) return HVLM_ERR_BAD_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HVLM_DISPATCH_DTYPE(dtype, TT, {
        colsum_kernel<TT><<<(N + 31) / 32, 256, 0, s>>>(static_cast<const TT*>(dy), db, M, N);
    });
    return check_last("colsum");
}
