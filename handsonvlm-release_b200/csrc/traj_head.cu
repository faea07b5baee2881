// Trajectory head after the <hand_traj> gather, generation side.
//
// Replaces, in one launch, what the reference's sampling loop does per <hand_traj> token
// (handsonvlm/model/language_model/handsonvlm.py:609-622): slice the last hidden row, de-interleave it into the two
// hands, TrajDecoder.inference (handsonvlm/model/language_model/traj_decoder.py:39-47) -> TrajCVAE.inference
// (hoi_forecast/architecture/traj_decoder.py:75-91, condition_contact=False) -> VAE.inference
// (hoi_forecast/architecture/decoder_modules.py:56-60):
//
//     x[r]   = cat(z[r] (L), cond[r] (Dc))                          r = (b, hand[, k])
//     h[r]   = ELU(W1 x[r] + b1)        W1 [H, L+Dc]                (dec_MLP.0, dec_MLP.1)
//     out[r] = W2 h[r] + b2             W2 [2, H]                   (dec_MLP.2)
//
// The work is one skinny GEMV pass over W1 (H x (L+Dc): 2.4 MB in bf16 at the 7B size), so the kernel is a
// bandwidth/latency problem: H/4 CTAs, one hidden unit per warp, 16-byte streaming loads of the weight row, the (few)
// activation rows staged once per CTA in shared memory as fp32 (with the even/odd de-interleave fused into the
// staging when the caller passes the raw hidden row), fp32 accumulation, ELU, and the 2-wide second layer folded in as
// per-CTA partial sums.  The last CTA to finish (threadfence + counter) adds the partials in a FIXED order, so the
// result is bit-reproducible run to run; the counter resets itself for the next call.
#include "hvlm_internal.cuh"
#include "hvlm_vec.cuh"

namespace hvlm {
namespace traj {

constexpr int kRows = 4;        // activation rows per CTA pass (grid.y covers ceil(R / kRows))
constexpr int kWarps = 4;       // hidden units per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kInFlight = 12;   // 16-byte weight vectors each lane keeps in flight (covers K <= 3072 in bf16 in one batch)

template <typename T>
__global__ void __launch_bounds__(kThreads)
traj_decode_kernel(const T* __restrict__ cond, int64_t ld_cond, int interleaved, const T* __restrict__ z,
                   const T* __restrict__ W1, const T* __restrict__ b1, const T* __restrict__ W2,
                   const T* __restrict__ b2, float* __restrict__ out, float* __restrict__ partial,
                   unsigned int* __restrict__ counter, int R, int Dc, int L, int H, int vec_ok) {
    extern __shared__ __align__(16) float xs[];                       // [kRows][K] fp32
    __shared__ float part_s[kWarps][kRows][2];
    __shared__ bool is_last;
    constexpr int V = Vec16<T>::N;
    const int K = L + Dc;
    const int r0 = blockIdx.y * kRows;
    const int nrows = min(kRows, R - r0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * kWarps + warp;                         // this warp's hidden unit
    const bool live = j < H;
    const T* wrow = W1 + static_cast<int64_t>(live ? j : 0) * K;

    // ---- issue the first batch of weight loads before anything else: their latency overlaps the staging below
    uint4 wv[kInFlight];
#pragma unroll
    for (int it = 0; it < kInFlight; ++it) {
        const int i0 = (it * 32 + lane) * V;
        if (live && i0 < K) wv[it] = ld_stream16(wrow + i0);
    }

    // ---- stage the activation rows as fp32: xs[rr][0:L) = z[r], xs[rr][L:K) = cond[r]
    if (vec_ok) {
        const int zv = L / V;                                         // vectors per z row
        for (int v = threadIdx.x; v < nrows * zv; v += kThreads) {
            const int rr = v / zv, c = (v - rr * zv) * V;
            float f[V];
            unpack16<T>(*reinterpret_cast<const uint4*>(z + static_cast<int64_t>(r0 + rr) * L + c), f);
#pragma unroll
            for (int e = 0; e < V; ++e) xs[rr * K + c + e] = f[e];
        }
        if (interleaved) {
            // one 16-byte vector of the hidden row holds V/2 (even, odd) pairs: evens -> hand 0, odds -> hand 1
            // (handsonvlm.py:615-616)
            const int hv = 2 * Dc / V;
            for (int v = threadIdx.x; v < (nrows >> 1) * hv; v += kThreads) {
                const int bb = v / hv, c = (v - bb * hv) * V;
                float f[V];
                unpack16<T>(*reinterpret_cast<const uint4*>(cond + static_cast<int64_t>((r0 >> 1) + bb) * ld_cond + c), f);
#pragma unroll
                for (int e = 0; e < V; e += 2) {
                    xs[(2 * bb) * K + L + ((c + e) >> 1)] = f[e];
                    xs[(2 * bb + 1) * K + L + ((c + e) >> 1)] = f[e + 1];
                }
            }
        } else {
            const int cv = Dc / V;
            for (int v = threadIdx.x; v < nrows * cv; v += kThreads) {
                const int rr = v / cv, c = (v - rr * cv) * V;
                float f[V];
                unpack16<T>(*reinterpret_cast<const uint4*>(cond + static_cast<int64_t>(r0 + rr) * ld_cond + c), f);
#pragma unroll
                for (int e = 0; e < V; ++e) xs[rr * K + L + c + e] = f[e];
            }
        }
    } else {
        for (int idx = threadIdx.x; idx < nrows * K; idx += kThreads) {
            const int rr = idx / K;
            const int i = idx - rr * K;
            const int r = r0 + rr;
            float v;
            if (i < L) {
                v = to_float<T>(z[static_cast<int64_t>(r) * L + i]);
            } else {
                const int jj = i - L;
                v = interleaved ? to_float<T>(cond[static_cast<int64_t>(r >> 1) * ld_cond + 2 * jj + (r & 1)])
                                : to_float<T>(cond[static_cast<int64_t>(r) * ld_cond + jj]);
            }
            xs[idx] = v;
        }
    }
    __syncthreads();

    // ---- layer 1: one hidden unit per warp, fp32 accumulate
    float acc[kRows];
#pragma unroll
    for (int rr = 0; rr < kRows; ++rr) acc[rr] = 0.f;
    if (live) {
        for (int base = 0; base < K; base += kInFlight * 32 * V) {
            if (base > 0) {
#pragma unroll
                for (int it = 0; it < kInFlight; ++it) {
                    const int i0 = base + (it * 32 + lane) * V;
                    if (i0 < K) wv[it] = ld_stream16(wrow + i0);
                }
            }
#pragma unroll
            for (int it = 0; it < kInFlight; ++it) {
                const int i0 = base + (it * 32 + lane) * V;
                if (i0 < K) {
                    float w[V];
                    unpack16<T>(wv[it], w);
#pragma unroll
                    for (int rr = 0; rr < kRows; ++rr) {
                        if (rr < nrows) {
                            const float* xr = xs + rr * K + i0;
#pragma unroll
                            for (int e = 0; e < V; e += 4) {
                                const float4 xv = *reinterpret_cast<const float4*>(xr + e);
                                acc[rr] = fmaf(w[e], xv.x, acc[rr]);
                                acc[rr] = fmaf(w[e + 1], xv.y, acc[rr]);
                                acc[rr] = fmaf(w[e + 2], xv.z, acc[rr]);
                                acc[rr] = fmaf(w[e + 3], xv.w, acc[rr]);
                            }
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int rr = 0; rr < kRows; ++rr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[rr] += __shfl_xor_sync(0xffffffffu, acc[rr], o);
    }
    if (lane == 0) {
        const float bj = live ? to_float<T>(b1[j]) : 0.f;
        const float w20 = live ? to_float<T>(W2[j]) : 0.f;
        const float w21 = live ? to_float<T>(W2[H + j]) : 0.f;
#pragma unroll
        for (int rr = 0; rr < kRows; ++rr) {
            const float a = acc[rr] + bj;
            const float h = a > 0.f ? a : expm1f(a);    // nn.ELU(alpha=1)
            part_s[warp][rr][0] = w20 * h;
            part_s[warp][rr][1] = w21 * h;
        }
    }
    __syncthreads();

    // ---- layer 2: per-CTA partial (fixed order over the warps), then the last CTA of this row chunk reduces all
    float* my_part = partial + (static_cast<int64_t>(blockIdx.y) * gridDim.x + blockIdx.x) * (kRows * 2);
    if (threadIdx.x < kRows * 2) {
        const int rr = threadIdx.x >> 1, o = threadIdx.x & 1;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += part_s[w][rr][o];
        my_part[threadIdx.x] = s;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(&counter[blockIdx.y], 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // warp w sums outputs w and w + kWarps: lanes take CTAs lane, lane+32, ... in order, then a fixed shuffle tree
    const float* pbase = partial + static_cast<int64_t>(blockIdx.y) * gridDim.x * (kRows * 2);
    for (int q = warp; q < kRows * 2; q += kWarps) {
        float s = 0.f;
        for (unsigned int c = lane; c < gridDim.x; c += 32) s += __ldcg(pbase + static_cast<int64_t>(c) * (kRows * 2) + q);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const int rr = q >> 1, o = q & 1;
        if (lane == 0 && rr < nrows) out[static_cast<int64_t>(r0 + rr) * 2 + o] = s + to_float<T>(b2[o]);
    }
    if (threadIdx.x == 0) counter[blockIdx.y] = 0;      // ready for the next call on this stream
}

// out[r, j] = act(b[j] + W[j, :] . x[r, :]) for a handful of rows: one output unit per warp, all 16-byte weight loads of the
// row in flight, activations staged once per CTA as fp32.  The building block of the MLP trajectory decoder
// (hoi_forecast/architecture/traj_decoder.py:94-104: Linear-ReLU-Linear-ReLU-Linear on R = 2B rows).
template <typename T>
__global__ void __launch_bounds__(kThreads)
skinny_linear_kernel(const T* __restrict__ x, int64_t ld_x, const T* __restrict__ W, const T* __restrict__ b,
                     T* __restrict__ out, int R, int K, int H, int act, int vec_ok) {
    extern __shared__ __align__(16) float xs[];                       // [kRows][K] fp32
    constexpr int V = Vec16<T>::N;
    const int r0 = blockIdx.y * kRows;
    const int nrows = min(kRows, R - r0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * kWarps + warp;
    const bool live = j < H;
    const T* wrow = W + static_cast<int64_t>(live ? j : 0) * K;
    uint4 wv[kInFlight];
#pragma unroll
    for (int it = 0; it < kInFlight; ++it) {
        const int i0 = (it * 32 + lane) * V;
        if (live && i0 < K) wv[it] = ld_stream16(wrow + i0);
    }
    if (vec_ok) {
        const int kv = K / V;
        for (int v = threadIdx.x; v < nrows * kv; v += kThreads) {
            const int rr = v / kv, c = (v - rr * kv) * V;
            float f[V];
            unpack16<T>(*reinterpret_cast<const uint4*>(x + static_cast<int64_t>(r0 + rr) * ld_x + c), f);
#pragma unroll
            for (int e = 0; e < V; ++e) xs[rr * K + c + e] = f[e];
        }
    } else {
        for (int idx = threadIdx.x; idx < nrows * K; idx += kThreads) {
            const int rr = idx / K;
            xs[idx] = to_float<T>(x[static_cast<int64_t>(r0 + rr) * ld_x + (idx - rr * K)]);
        }
    }
    __syncthreads();
    float acc[kRows];
#pragma unroll
    for (int rr = 0; rr < kRows; ++rr) acc[rr] = 0.f;
    if (live) {
        for (int base = 0; base < K; base += kInFlight * 32 * V) {
            if (base > 0) {
#pragma unroll
                for (int it = 0; it < kInFlight; ++it) {
                    const int i0 = base + (it * 32 + lane) * V;
                    if (i0 < K) wv[it] = ld_stream16(wrow + i0);
                }
            }
#pragma unroll
            for (int it = 0; it < kInFlight; ++it) {
                const int i0 = base + (it * 32 + lane) * V;
                if (i0 < K) {
                    float w[V];
                    unpack16<T>(wv[it], w);
#pragma unroll
                    for (int rr = 0; rr < kRows; ++rr) {
                        if (rr < nrows) {
#pragma unroll
                            for (int e = 0; e < V; ++e) acc[rr] = fmaf(w[e], xs[rr * K + i0 + e], acc[rr]);
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int rr = 0; rr < kRows; ++rr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[rr] += __shfl_xor_sync(0xffffffffu, acc[rr], o);
    }
    if (lane == 0 && live) {
        const float bj = b ? to_float<T>(b[j]) : 0.f;
#pragma unroll
        for (int rr = 0; rr < kRows; ++rr) {
            if (rr < nrows) {
                float a = acc[rr] + bj;
                if (act == 1) a = fmaxf(a, 0.f);
                else if (act == 2) a = a > 0.f ? a : expm1f(a);
                out[static_cast<int64_t>(r0 + rr) * H + j] = from_float<T>(a);
            }
        }
    }
}

inline int grid_x(int H) { return (H + kWarps - 1) / kWarps; }
inline int grid_y(int R) { return (R + kRows - 1) / kRows; }
// workspace layout: [0, kCounterBytes) one arrival counter per row chunk (always at the same place, so the "left
// zeroed" invariant survives calls with different R / H), then the per-CTA partial sums.
constexpr int kMaxChunks = 1024;                        // row chunks per launch (4096 rows); more rows -> more launches
constexpr size_t kCounterBytes = kMaxChunks * sizeof(unsigned int);

}  // namespace traj
}  // namespace hvlm

extern "C" size_t hvlm_traj_decode_workspace_bytes(int R, int H) {
    using namespace hvlm::traj;
    if (R <= 0 || H <= 0) return 0;
    const int chunks = grid_y(R) < kMaxChunks ? grid_y(R) : kMaxChunks;
    return kCounterBytes + static_cast<size_t>(grid_x(H)) * chunks * kRows * 2 * sizeof(float);
}

extern "C" int hvlm_traj_decode(const void* cond, int64_t ld_cond, int interleaved, const void* z, const void* W1,
                                const void* b1, const void* W2, const void* b2, int dtype, int R, int Dc, int L, int H,
                                float* out, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace hvlm;
    using namespace hvlm::traj;
    if (!cond || !z || !W1 || !b1 || !W2 || !b2 || !out || !workspace) return HVLM_ERR_BAD_ARG;
    if (R <= 0 || Dc <= 0 || L <= 0 || H <= 0) return HVLM_ERR_BAD_SHAPE;
    if (interleaved && (R & 1)) return HVLM_ERR_BAD_SHAPE;
    const int elt = dtype == HVLM_F32 ? 4 : 2;
    const int K = L + Dc;
    // 16-byte weight-row loads: K and the base pointers must keep every row 16-byte aligned
    if ((K * elt) % 16 != 0 || (reinterpret_cast<uintptr_t>(W1) & 15u)) return HVLM_ERR_ALIGN;
    if (interleaved ? ld_cond < 2 * static_cast<int64_t>(Dc) : ld_cond < Dc) return HVLM_ERR_BAD_SHAPE;
    if (workspace_bytes < hvlm_traj_decode_workspace_bytes(R, H)) return HVLM_ERR_WORKSPACE;
    const size_t smem = static_cast<size_t>(kRows) * K * sizeof(float);
    if (smem > 200 * 1024) return HVLM_ERR_UNSUPPORTED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned int* counter = static_cast<unsigned int*>(workspace);
    float* partial = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + kCounterBytes);
    // vectorised staging needs 16-byte aligned rows everywhere; anything else takes the scalar staging loop
    const int V = 16 / elt;
    const int vec_ok = (L % V == 0) && (Dc % V == 0) && (ld_cond % V == 0) && ((reinterpret_cast<uintptr_t>(cond) & 15u) == 0) &&
                       ((reinterpret_cast<uintptr_t>(z) & 15u) == 0) && (!interleaved || (R % 2 == 0));
    StageTimer st(HVLM_STAGE_GATHER, s);
    constexpr int kRowsPerLaunch = kMaxChunks * kRows;
    HVLM_DISPATCH_DTYPE(dtype, TT, {
        auto kern = traj_decode_kernel<TT>;
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
            return HVLM_ERR_CUDA;
        for (int rb = 0; rb < R; rb += kRowsPerLaunch) {                      // one launch unless R > 4096
            const int rows = R - rb < kRowsPerLaunch ? R - rb : kRowsPerLaunch;
            const TT* c = static_cast<const TT*>(cond) + (interleaved ? (rb >> 1) : rb) * ld_cond;
            kern<<<dim3(grid_x(H), grid_y(rows)), kThreads, smem, s>>>(
                c, ld_cond, interleaved, static_cast<const TT*>(z) + static_cast<int64_t>(rb) * L,
                static_cast<const TT*>(W1), static_cast<const TT*>(b1), static_cast<const TT*>(W2),
                static_cast<const TT*>(b2), out + static_cast<int64_t>(rb) * 2, partial, counter, rows, Dc, L, H, vec_ok);
            if (rb + kRowsPerLaunch < R) count_launch();
        }
    });
    return check_last("traj_decode");
}

extern "C" int hvlm_skinny_linear(const void* x, int64_t ld_x, const void* W, const void* b, int act, int dtype, int R,
                                  int K, int H, void* out, void* stream) {
    using namespace hvlm;
    using namespace hvlm::traj;
    if (!x || !W || !out) return HVLM_ERR_BAD_ARG;
    if (R <= 0 || K <= 0 || H <= 0 || ld_x < K || act < 0 || act > 2) return HVLM_ERR_BAD_SHAPE;
    const int elt = dtype == HVLM_F32 ? 4 : 2;
    if ((K * elt) % 16 != 0 || (reinterpret_cast<uintptr_t>(W) & 15u)) return HVLM_ERR_ALIGN;
    const size_t smem = static_cast<size_t>(kRows) * K * sizeof(float);
    if (smem > 200 * 1024) return HVLM_ERR_UNSUPPORTED;
    const int V = 16 / elt;
    const int vec_ok = (ld_x % V == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    StageTimer st(HVLM_STAGE_GATHER, s);
    HVLM_DISPATCH_DTYPE(dtype, TT, {
        auto kern = skinny_linear_kernel<TT>;
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
            return HVLM_ERR_CUDA;
        kern<<<dim3(grid_x(H), grid_y(R)), kThreads, smem, s>>>(static_cast<const TT*>(x), ld_x, static_cast<const TT*>(W),
                                                              static_cast<const TT*>(b), static_cast<TT*>(out), R, K, H,
                                                              act, vec_ok);
    });
    return check_last("skinny_linear");
}
