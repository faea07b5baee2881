// LITA slow-fast token pooling (forward + backward), HBM-bound.
//
// Replaces the pooling branch of LitaMetaForCausalLM.videos_to_tokens (lita/model/lita_arch.py:41-73) and
// VisualToTokenHelper.compress_tokens (hoi_forecast/model/visual_to_tokens.py:230-272):
//   fast[b,f,:]        = mean over the 256 tokens of frame f
//   slow[b,64k+8h+w,:] = mean of the 2x2 cell (h,w) of frame sel[k], sel = round(linspace(0,t-1,4))
// The reference gathers 4 frames, permutes to NCHW, runs avg_pool2d, permutes back, runs a second
// reduction over the full tensor and concatenates.  Here ONE pass reads every input element exactly once
// (slow frames are a subset of what the fast mean reads) and writes the concatenated layout directly.
//
// Mapping: CTA = (channel slab, frame, clip), 8 warps.  A half-warp covers one token's slab with 16-byte
// vectors (16 lanes x 16 B = 256 B contiguous); the two half-warps take the left/right token of a 2x2 cell,
// each lane adds the cell's upper and lower token, a warp shuffle (xor 16) completes the cell sum.  Cell sums
// feed the slow token (x 1/4) and the per-warp fast accumulator; the 8 warps are combined through
// shared memory.
#include "hvlm_internal.cuh"
#include "hvlm_vec.cuh"

namespace hvlm {

struct SelFrames {
    int sel[4];
};

static SelFrames selected_frames(int t) {
    // np.round(np.linspace(0, t-1, 4)).astype(int): k*(t-1)/3 is never a .5 tie for integers.
    SelFrames s;
    for (int k = 0; k < 4; ++k) s.sel[k] = static_cast<int>(__builtin_rint(static_cast<double>(k) * (t - 1) / 3.0));
    return s;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) pool_slowfast_fwd_kernel(const TIn* __restrict__ tok, int64_t frame_stride,
                                                                TOut* __restrict__ out, int t, int C, int n_out,
                                                                int fast_row0 /* -1: no fast rows */,
                                                                int slow_row0 /* -1: no slow rows */, SelFrames sf,
                                                                int frames_from_sel,
                                                                const int32_t* __restrict__ frame_map) {
    constexpr int V = Vec16<TIn>::N;         // channels per lane
    constexpr int SLAB = 16 * V;             // channels per CTA
    __shared__ float red[8][16][V];

    const int b = blockIdx.z;
    const int f = frames_from_sel ? sf.sel[blockIdx.y] : static_cast<int>(blockIdx.y);
    const int c0 = blockIdx.x * SLAB;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int half = lane >> 4;              // 0: left token of the cell, 1: right token
    const int cl = (lane & 15) * V;          // channel offset inside the slab
    const bool ch_ok = (c0 + cl) < C;        // C % V == 0 is guaranteed by the host

    // which slow slots does this frame feed?  (duplicates when t < 4)
    int kmask = 0;
    if (slow_row0 >= 0) {
        if (frames_from_sel) {
            kmask = 1 << blockIdx.y;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) kmask |= (sf.sel[k] == f) << k;
        }
    }

    // optional indirection (frame de-duplication): logical frame (b,f) lives at row block frame_map[b*t+f]
    const int64_t fsrc = frame_map ? static_cast<int64_t>(frame_map[b * t + f]) : static_cast<int64_t>(b) * t + f;
    const TIn* src = tok + fsrc * frame_stride * C + c0 + cl;
    float facc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) facc[i] = 0.f;

    // each warp owns 8 of the 64 cells; issue all 16 loads of the warp first (memory-level parallelism)
    uint4 raw[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int cell = warp * 8 + j;
        const int h = cell >> 3, w = cell & 7;
        const int s_top = (2 * h) * 16 + 2 * w + half;
        if (ch_ok) {
            raw[j][0] = ld_stream16(src + static_cast<int64_t>(s_top) * C);
            raw[j][1] = ld_stream16(src + static_cast<int64_t>(s_top + 16) * C);
        } else {
            raw[j][0] = make_uint4(0, 0, 0, 0);
            raw[j][1] = make_uint4(0, 0, 0, 0);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float a[V], bb[V];
        unpack16<TIn>(raw[j][0], a);
        unpack16<TIn>(raw[j][1], bb);
        float cs[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            cs[i] = a[i] + bb[i];
            cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 16);   // + the other half of the cell
            facc[i] += cs[i];
        }
        if (kmask && half == 0 && ch_ok) {
            const int cell = warp * 8 + j;
            float o[V];
#pragma unroll
            for (int i = 0; i < V; ++i) o[i] = cs[i] * 0.25f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (kmask & (1 << k)) {
                    TOut* dst = out + (static_cast<int64_t>(b) * n_out + slow_row0 + 64 * k + cell) * C + c0 + cl;
                    if constexpr (sizeof(TOut) * V == 16) {
                        *reinterpret_cast<uint4*>(dst) = pack16<TOut>(o);
                    } else {
#pragma unroll
                        for (int i = 0; i < V; ++i) dst[i] = from_float<TOut>(o[i]);
                    }
                }
            }
        }
    }

    if (fast_row0 >= 0) {
        if (half == 0) {
#pragma unroll
            for (int i = 0; i < V; ++i) red[warp][lane][i] = facc[i];
        }
        __syncthreads();
        if (warp == 0 && half == 0 && ch_ok) {
            float o[V];
#pragma unroll
            for (int i = 0; i < V; ++i) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += red[w][lane][i];
                o[i] = s * (1.0f / 256.0f);
            }
            TOut* dst = out + (static_cast<int64_t>(b) * n_out + fast_row0 + f) * C + c0 + cl;
            if constexpr (sizeof(TOut) * V == 16) {
                *reinterpret_cast<uint4*>(dst) = pack16<TOut>(o);
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) dst[i] = from_float<TOut>(o[i]);
            }
        }
    }
}

// 'spatial' arch: mean over frames.  CTA = (slab, token, clip) loops over t (ablation modes only).
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(128) pool_spatial_mean_kernel(const TIn* __restrict__ tok, int64_t frame_stride,
                                                                TOut* __restrict__ out, int t, int C, int n_out,
                                                                int row0) {
    constexpr int V = Vec16<TIn>::N;
    const int b = blockIdx.z, s = blockIdx.y;
    const int c = (blockIdx.x * 128 + threadIdx.x) * V;
    if (c >= C) return;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    for (int f = 0; f < t; ++f) {
        float a[V];
        unpack16<TIn>(ld_stream16(tok + ((static_cast<int64_t>(b) * t + f) * frame_stride + s) * C + c), a);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] += a[i];
    }
    TOut* dst = out + (static_cast<int64_t>(b) * n_out + row0 + s) * C + c;
#pragma unroll
    for (int i = 0; i < V; ++i) dst[i] = from_float<TOut>(acc[i] / static_cast<float>(t));
}

// backward: d_tok[b,f,s,:] = d_fast[b,f,:]/256 + sum_{k: sel[k]==f} d_slow[b,64k+cell(s),:]/4
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) pool_slowfast_bwd_kernel(const TIn* __restrict__ dout, TOut* __restrict__ dtok,
                                                                int t, int C, int n_out, int fast_row0, int slow_row0,
                                                                SelFrames sf) {
    constexpr int V = Vec16<TIn>::N;
    constexpr int SLAB = 16 * V;
    const int b = blockIdx.z, f = blockIdx.y;
    const int c0 = blockIdx.x * SLAB;
    const int lane16 = threadIdx.x & 15;
    const int grp = threadIdx.x >> 4;        // 16 token groups per CTA
    const int c = c0 + lane16 * V;
    if (c >= C) return;
    int kmask = 0;
    if (slow_row0 >= 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) kmask |= (sf.sel[k] == f) << k;
    }
    float gf[V];
#pragma unroll
    for (int i = 0; i < V; ++i) gf[i] = 0.f;
    if (fast_row0 >= 0) {
        float a[V];
        unpack16<TIn>(*reinterpret_cast<const uint4*>(dout + (static_cast<int64_t>(b) * n_out + fast_row0 + f) * C + c), a);
#pragma unroll
        for (int i = 0; i < V; ++i) gf[i] = a[i] * (1.0f / 256.0f);
    }
    for (int s = grp; s < 256; s += 16) {
        float g[V];
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] = gf[i];
        if (kmask) {
            const int cell = ((s >> 4) >> 1) * 8 + ((s & 15) >> 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (kmask & (1 << k)) {
                    float a[V];
                    unpack16<TIn>(*reinterpret_cast<const uint4*>(
                                      dout + (static_cast<int64_t>(b) * n_out + slow_row0 + 64 * k + cell) * C + c),
                                  a);
#pragma unroll
                    for (int i = 0; i < V; ++i) g[i] += a[i] * 0.25f;
                }
            }
        }
        TOut* dst = dtok + ((static_cast<int64_t>(b) * t + f) * 256 + s) * C + c;
        if constexpr (sizeof(TOut) * V == 16) {
            st_stream16(dst, pack16<TOut>(g));
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i) dst[i] = from_float<TOut>(g[i]);
        }
    }
}

template <typename TIn, typename TOut>
static int pool_fwd_typed(const void* tok, int64_t frame_stride, void* out, int B, int t, int C, int mode,
                          cudaStream_t s, const int32_t* frame_map = nullptr) {
    constexpr int V = Vec16<TIn>::N;
    if (C % V != 0) return HVLM_ERR_BAD_SHAPE;
    const int n_out = hvlm_pool_out_tokens(t, mode);
    const SelFrames sf = selected_frames(t);
    const int slabs = (C + 16 * V - 1) / (16 * V);
    const TIn* in = static_cast<const TIn*>(tok);
    TOut* o = static_cast<TOut*>(out);
    switch (mode) {
        case HVLM_POOL_TEMPORAL_SPATIAL_POOL:
            pool_slowfast_fwd_kernel<TIn, TOut><<<dim3(slabs, t, B), 256, 0, s>>>(in, frame_stride, o, t, C, n_out, 0, t, sf, 0, frame_map);
            break;
        case HVLM_POOL_SPATIAL_POOL:
            pool_slowfast_fwd_kernel<TIn, TOut><<<dim3(slabs, 4, B), 256, 0, s>>>(in, frame_stride, o, t, C, n_out, -1, 0, sf, 1, frame_map);
            break;
        case HVLM_POOL_TEMPORAL:
            pool_slowfast_fwd_kernel<TIn, TOut><<<dim3(slabs, t, B), 256, 0, s>>>(in, frame_stride, o, t, C, n_out, 0, -1, sf, 0, frame_map);
            break;
        case HVLM_POOL_SPATIAL:
        case HVLM_POOL_TEMPORAL_SPATIAL: {
            if (frame_map) return HVLM_ERR_UNSUPPORTED;
            int row0 = 0;
            if (mode == HVLM_POOL_TEMPORAL_SPATIAL) {
                pool_slowfast_fwd_kernel<TIn, TOut><<<dim3(slabs, t, B), 256, 0, s>>>(in, frame_stride, o, t, C, n_out, 0, -1, sf, 0, frame_map);
                row0 = t;
            }
            const int gx = (C / V + 127) / 128;
            pool_spatial_mean_kernel<TIn, TOut><<<dim3(gx, 256, B), 128, 0, s>>>(in, frame_stride, o, t, C, n_out, row0);
        } break;
        default:
            return HVLM_ERR_BAD_ARG;
    }
    return check_last("pool_fwd");
}

template <typename TIn, typename TOut>
static int pool_bwd_typed(const void* dout, void* dtok, int B, int t, int C, int mode, cudaStream_t s) {
    constexpr int V = Vec16<TIn>::N;
    if (C % V != 0) return HVLM_ERR_BAD_SHAPE;
    const int n_out = hvlm_pool_out_tokens(t, mode);
    const SelFrames sf = selected_frames(t);
    const int slabs = (C + 16 * V - 1) / (16 * V);
    int fast_row0, slow_row0;
    switch (mode) {
        case HVLM_POOL_TEMPORAL_SPATIAL_POOL: fast_row0 = 0; slow_row0 = t; break;
        case HVLM_POOL_SPATIAL_POOL: fast_row0 = -1; slow_row0 = 0; break;
        case HVLM_POOL_TEMPORAL: fast_row0 = 0; slow_row0 = -1; break;
        default: return HVLM_ERR_UNSUPPORTED;
    }
    pool_slowfast_bwd_kernel<TIn, TOut><<<dim3(slabs, t, B), 256, 0, s>>>(
        static_cast<const TIn*>(dout), static_cast<TOut*>(dtok), t, C, n_out, fast_row0, slow_row0, sf);
    return check_last("pool_bwd");
}

}  // namespace hvlm

extern "C" int hvlm_pool_out_tokens(int t, int mode) {
    switch (mode) {
        case HVLM_POOL_TEMPORAL_SPATIAL_POOL: return t + 256;
        case HVLM_POOL_SPATIAL_POOL: return 256;
        case HVLM_POOL_TEMPORAL: return t;
        case HVLM_POOL_SPATIAL: return 256;
        case HVLM_POOL_TEMPORAL_SPATIAL: return t + 256;
        default: return HVLM_ERR_BAD_ARG;
    }
}

extern "C" int hvlm_pool_slowfast_fwd(const void* tok, int in_dtype, int64_t frame_stride, void* out, int out_dtype,
                                      int B, int t, int C, int mode, void* stream) {
    using namespace hvlm;
    if (!tok || !out || B <= 0 || t <= 0 || C <= 0) return HVLM_ERR_BAD_ARG;
    if (frame_stride < 256) return HVLM_ERR_BAD_SHAPE;
    if (!aligned16(tok) || !aligned16(out)) return HVLM_ERR_ALIGN;
    if (in_dtype == HVLM_F16 || out_dtype == HVLM_F16) return HVLM_ERR_BAD_DTYPE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    StageTimer st(HVLM_STAGE_POOL, s);
    if (in_dtype == HVLM_F32 && out_dtype == HVLM_F32) return pool_fwd_typed<float, float>(tok, frame_stride, out, B, t, C, mode, s);
    if (in_dtype == HVLM_F32 && out_dtype == HVLM_BF16) return pool_fwd_typed<float, __nv_bfloat16>(tok, frame_stride, out, B, t, C, mode, s);
    if (in_dtype == HVLM_BF16 && out_dtype == HVLM_BF16) return pool_fwd_typed<__nv_bfloat16, __nv_bfloat16>(tok, frame_stride, out, B, t, C, mode, s);
    if (in_dtype == HVLM_BF16 && out_dtype == HVLM_F32) return pool_fwd_typed<__nv_bfloat16, float>(tok, frame_stride, out, B, t, C, mode, s);
    return HVLM_ERR_BAD_DTYPE;
}

extern "C" int hvlm_pool_slowfast_bwd(const void* dout, int dout_dtype, void* dtok, int dtok_dtype, int B, int t, int C,
                                      int mode, void* stream) {
    using namespace hvlm;
    if (!dout || !dtok || B <= 0 || t <= 0 || C <= 0) return HVLM_ERR_BAD_ARG;
    if (!aligned16(dout) || !aligned16(dtok)) return HVLM_ERR_ALIGN;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dout_dtype == HVLM_F32 && dtok_dtype == HVLM_F32) return pool_bwd_typed<float, float>(dout, dtok, B, t, C, mode, s);
    if (dout_dtype == HVLM_BF16 && dtok_dtype == HVLM_BF16) return pool_bwd_typed<__nv_bfloat16, __nv_bfloat16>(dout, dtok, B, t, C, mode, s);
    if (dout_dtype == HVLM_F32 && dtok_dtype == HVLM_BF16) return pool_bwd_typed<float, __nv_bfloat16>(dout, dtok, B, t, C, mode, s);
    if (dout_dtype == HVLM_BF16 && dtok_dtype == HVLM_F32) return pool_bwd_typed<__nv_bfloat16, float>(dout, dtok, B, t, C, mode, s);
    return HVLM_ERR_BAD_DTYPE;
}

extern "C" int hvlm_pool_slowfast_fwd_mapped(const void* tok, int in_dtype, int64_t frame_stride, const int32_t* frame_map,
                                             void* out, int out_dtype, int B, int t, int C, int mode, void* stream) {
    using namespace hvlm;
    if (!tok || !out || !frame_map || B <= 0 || t <= 0 || C <= 0) return HVLM_ERR_BAD_ARG;
    if (frame_stride < 256) return HVLM_ERR_BAD_SHAPE;
    if (!aligned16(tok) || !aligned16(out)) return HVLM_ERR_ALIGN;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    StageTimer st(HVLM_STAGE_POOL, s);
    if (in_dtype == HVLM_F32 && out_dtype == HVLM_F32) return pool_fwd_typed<float, float>(tok, frame_stride, out, B, t, C, mode, s, frame_map);
    if (in_dtype == HVLM_F32 && out_dtype == HVLM_BF16) return pool_fwd_typed<float, __nv_bfloat16>(tok, frame_stride, out, B, t, C, mode, s, frame_map);
    if (in_dtype == HVLM_BF16 && out_dtype == HVLM_BF16) return pool_fwd_typed<__nv_bfloat16, __nv_bfloat16>(tok, frame_stride, out, B, t, C, mode, s, frame_map);
    if (in_dtype == HVLM_BF16 && out_dtype == HVLM_F32) return pool_fwd_typed<__nv_bfloat16, float>(tok, frame_stride, out, B, t, C, mode, s, frame_map);
    return HVLM_ERR_BAD_DTYPE;
}
