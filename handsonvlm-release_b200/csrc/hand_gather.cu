// <hand_traj> hidden-state gather (forward, backward, generation step).
//
// Replaces the per-sample Python loop inside HandsOnVLMForCausalLM.forward
// (handsonvlm/model/language_model/handsonvlm.py:146-187: bool-mask index + reshape(4,D/2,2).permute(2,0,1),
// one host sync per sample) and the generation-time gather (:609-622).  Four CTAs per sample: a block scan over
// the shifted label mask finds the (up to) 4 predictor rows, CTA k copies row k with the even/odd channel
// de-interleave  out[b,h,k,j] = hidden[b,row_k,2j+h].  No host sync; counts[] lets the caller enforce the
// reference's "0 or 4" contract.
#include "hvlm_internal.cuh"
#include "hvlm_scan.cuh"
#include "hvlm_vec.cuh"

namespace hvlm {

// grid = (B, 4): every CTA scans its sample's labels (cheap), CTA (b,k) then copies predictor row k with 16-byte
// vector loads, de-interleaving even/odd channels in registers.
template <typename T>
__global__ void __launch_bounds__(kPlanThreads)
hand_gather_fwd_kernel(const T* __restrict__ hidden, const int64_t* __restrict__ labels, int64_t hand_id, int L, int D,
                       T* __restrict__ out, uint8_t* __restrict__ valid, int32_t* __restrict__ rows,
                       int32_t* __restrict__ counts) {
    __shared__ int scan_smem[9];
    __shared__ int srow[4];
    const int b = blockIdx.x;
    const int k = blockIdx.y;
    const int64_t* lab = labels + static_cast<int64_t>(b) * L;
    if (threadIdx.x < 4) srow[threadIdx.x] = -1;
    int seen = 0;
    // position i predicts label i+1:  m_shift[i] = (labels[i+1] == hand_id), i in [0, L-1)
    for (int i0 = 0; i0 < L - 1; i0 += kPlanThreads) {
        const int i = i0 + threadIdx.x;
        const int hit = (i < L - 1) && (lab[i + 1] == hand_id);
        int total;
        const int ord = seen + block_excl_scan(hit, &total, scan_smem);
        if (hit && ord < 4) srow[ord] = i;
        seen += total;
    }
    __syncthreads();
    if (k == 0) {
        if (threadIdx.x == 0) {
            counts[b] = seen;
            valid[b] = seen > 0;
        }
        if (threadIdx.x < 4) rows[b * 4 + threadIdx.x] = srow[threadIdx.x];
    }
    const int half = D >> 1;
    const int r = srow[k];
    T* dst_e = out + ((static_cast<int64_t>(b) * 2 + 0) * 4 + k) * half;
    T* dst_o = out + ((static_cast<int64_t>(b) * 2 + 1) * 4 + k) * half;
    const T* src = hidden + (static_cast<int64_t>(b) * L + (r >= 0 ? r : 0)) * D;
    constexpr int V = 16 / sizeof(T);                  // elements per 16-byte vector
    if ((D % (2 * V)) == 0 && (reinterpret_cast<uintptr_t>(hidden) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
        // two input vectors (2V elements = V pairs) -> one even vector + one odd vector
        for (int p = threadIdx.x * V; p < half; p += blockDim.x * V) {
            T in[2 * V], ev[V], od[V];
            if (r >= 0) {
                *reinterpret_cast<uint4*>(&in[0]) = *reinterpret_cast<const uint4*>(src + 2 * p);
                *reinterpret_cast<uint4*>(&in[V]) = *reinterpret_cast<const uint4*>(src + 2 * p + V);
            } else {
#pragma unroll
                for (int i = 0; i < 2 * V; ++i) in[i] = from_float<T>(0.f);
            }
#pragma unroll
            for (int i = 0; i < V; ++i) {
                ev[i] = in[2 * i];
                od[i] = in[2 * i + 1];
            }
            *reinterpret_cast<uint4*>(dst_e + p) = *reinterpret_cast<uint4*>(&ev[0]);
            *reinterpret_cast<uint4*>(dst_o + p) = *reinterpret_cast<uint4*>(&od[0]);
        }
    } else {
        for (int j = threadIdx.x; j < half; j += blockDim.x) {
            dst_e[j] = r >= 0 ? src[2 * j] : from_float<T>(0.f);
            dst_o[j] = r >= 0 ? src[2 * j + 1] : from_float<T>(0.f);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
hand_gather_bwd_kernel(const T* __restrict__ dout, const int32_t* __restrict__ rows, int L, int D,
                       float* __restrict__ d_hidden) {
    const int b = blockIdx.x;
    const int half = D >> 1;
    for (int idx = threadIdx.x; idx < 4 * half; idx += blockDim.x) {
        const int k = idx / half;
        const int j = idx - k * half;
        const int r = rows[b * 4 + k];
        if (r < 0) continue;
        const T* src = dout + static_cast<int64_t>(b) * 2 * 4 * half;
        float* dst = d_hidden + (static_cast<int64_t>(b) * L + r) * D + 2 * j;
        // rows of one sample are distinct, so plain read-modify-write is race free
        dst[0] += to_float<T>(src[(0 * 4 + k) * half + j]);
        dst[1] += to_float<T>(src[(1 * 4 + k) * half + j]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) hand_gather_step_kernel(const T* __restrict__ h, int D, T* __restrict__ out) {
    const int b = blockIdx.x;
    const int half = D >> 1;
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
        out[(static_cast<int64_t>(b) * 2 + 0) * half + j] = h[static_cast<int64_t>(b) * D + 2 * j];
        out[(static_cast<int64_t>(b) * 2 + 1) * half + j] = h[static_cast<int64_t>(b) * D + 2 * j + 1];
    }
}

}  // namespace hvlm

extern "C" int hvlm_hand_gather_fwd(const void* hidden, int dtype, const int64_t* labels, int64_t hand_id, int B, int L,
                                    int D, void* out, uint8_t* valid, int32_t* rows, int32_t* counts, void* stream) {
    using namespace hvlm;
    if (!hidden || !labels || !out || !valid || !rows || !counts) return HVLM_ERR_BAD_ARG;
    if (B <= 0 || L <= 0 || D <= 0 || (D & 1)) return HVLM_ERR_BAD_SHAPE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    StageTimer st(HVLM_STAGE_GATHER, s);
    HVLM_DISPATCH_DTYPE(dtype, TT, {
        hand_gather_fwd_kernel<TT><<<dim3(B, 4), kPlanThreads, 0, s>>>(static_cast<const TT*>(hidden), labels, hand_id, L, D,
                                                             static_cast<TT*>(out), valid, rows, counts);
    });
    return check_last("hand_gather_fwd");
}

extern "C" int hvlm_hand_gather_bwd(const void* dout, int dtype, const int32_t* rows, int B, int L, int D,
                                    float* d_hidden, void* stream) {
    using namespace hvlm;
    if (!dout || !rows || !d_hidden) return HVLM_ERR_BAD_ARG;
    if (B <= 0 || L <= 0 || D <= 0 || (D & 1)) return HVLM_ERR_BAD_SHAPE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HVLM_DISPATCH_DTYPE(dtype, TT, {
        hand_gather_bwd_kernel<TT><<<B, 256, 0, s>>>(static_cast<const TT*>(dout), rows, L, D, d_hidden);
    });
    return check_last("hand_gather_bwd");
}

extern "C" int hvlm_hand_gather_step(const void* hidden_last, int dtype, int B, int D, void* out, void* stream) {
    using namespace hvlm;
    if (!hidden_last || !out) return HVLM_ERR_BAD_ARG;
    if (B <= 0 || D <= 0 || (D & 1)) return HVLM_ERR_BAD_SHAPE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HVLM_DISPATCH_DTYPE(dtype, TT, {
        hand_gather_step_kernel<TT><<<B, 256, 0, s>>>(static_cast<const TT*>(hidden_last), D, static_cast<TT*>(out));
    });
    return check_last("hand_gather_step");
}
