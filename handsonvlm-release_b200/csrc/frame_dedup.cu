// Frame de-duplication in front of the ViT tower and the row gather that compacts the distinct frames.
//
// Real clips repeat frames -- EPIC clips are 10 distinct frames tiled x10 (handsonvlm/dataset/epic_dataset.py:90-95,
// handsonvlm/evaluation/handsonvlm_inference.py:205), single images are tiled x100 (hybrid_dataset.py:141-142) -- and the
// reference encodes every copy.  Three launches, no host synchronisation, every result stays on the device:
//   1. frame_hash_kernel    one streaming pass: 64-bit multiply-add checksum per frame.  A frame is cut into kSlices
//                           slices (CTA = (slice, frame)) whose partial sums are combined with a 64-bit atomic add --
//                           integer addition is associative, so the result does not depend on the order.
//   2. frame_confirm_kernel candidate = first earlier frame with the same checksum; the bytes are compared against it
//                           (second pass, over the candidate duplicates only); any difference marks the frame as distinct
//                           (a checksum collision costs the de-duplication of that frame, never correctness).
//   3. frame_finalize_kernel one CTA: flags -> exclusive scan -> frame_map / rep / n_unique.
// All HBM traffic is 16-byte vector loads; grid = kSlices x n_frames CTAs.
#include "hvlm_internal.cuh"
#include "hvlm_scan.cuh"
#include "hvlm_vec.cuh"

namespace hvlm {
namespace dedup {

constexpr int kSlices = 8;
constexpr int kThreads = 256;

struct Ws {
    unsigned long long* hash;   // [n]
    int32_t* cand;              // [n]  candidate representative (first earlier frame with the same checksum, else self)
    int32_t* differs;           // [n]  != 0: the bytes differ from the candidate's
};

__host__ __device__ inline size_t ws_bytes(int n) {
    const size_t a = (static_cast<size_t>(n) * 8 + 15) / 16 * 16;
    const size_t b = (static_cast<size_t>(n) * 4 + 15) / 16 * 16;
    return a + 2 * b;
}
static Ws carve(void* ws, int n) {
    uint8_t* p = static_cast<uint8_t*>(ws);
    const size_t a = (static_cast<size_t>(n) * 8 + 15) / 16 * 16;
    const size_t b = (static_cast<size_t>(n) * 4 + 15) / 16 * 16;
    return Ws{reinterpret_cast<unsigned long long*>(p), reinterpret_cast<int32_t*>(p + a), reinterpret_cast<int32_t*>(p + a + b)};
}

// splitmix64 finaliser: per-position odd multiplier, so that permuted or shifted content changes the checksum
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(kThreads)
frame_hash_kernel(const uint4* __restrict__ frames, size_t vec_per_frame, unsigned long long* __restrict__ hash) {
    __shared__ unsigned long long wsum[kThreads / 32];
    const int f = blockIdx.y, slice = blockIdx.x;
    const size_t per = (vec_per_frame + kSlices - 1) / kSlices;
    const size_t v0 = static_cast<size_t>(slice) * per;
    const size_t v1 = v0 + per < vec_per_frame ? v0 + per : vec_per_frame;
    const uint4* src = frames + static_cast<size_t>(f) * vec_per_frame;
    unsigned long long acc = 0;
    for (size_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
        const uint4 w = ld_stream16(src + v);
        const unsigned long long k = mix64(v) | 1ull;
        const unsigned long long lo = (static_cast<unsigned long long>(w.y) << 32) | w.x;
        const unsigned long long hi = (static_cast<unsigned long long>(w.w) << 32) | w.z;
        acc += lo * k + (hi ^ (lo >> 7)) * (k * 0xD6E8FEB86659FD93ull | 1ull);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) s += wsum[w];
        atomicAdd(hash + f, s);
    }
}

__global__ void __launch_bounds__(kThreads)
frame_confirm_kernel(const uint4* __restrict__ frames, size_t vec_per_frame, const unsigned long long* __restrict__ hash,
                     int32_t* __restrict__ cand, int32_t* __restrict__ differs) {
    __shared__ int s_cand;
    const int f = blockIdx.y, slice = blockIdx.x;
    if (threadIdx.x == 0) s_cand = f;
    __syncthreads();
    const unsigned long long h = hash[f];
    for (int j = threadIdx.x; j < f; j += kThreads)
        if (hash[j] == h) atomicMin(&s_cand, j);
    __syncthreads();
    const int c = s_cand;
    if (slice == 0 && threadIdx.x == 0) cand[f] = c;
    if (c == f) return;
    const size_t per = (vec_per_frame + kSlices - 1) / kSlices;
    const size_t v0 = static_cast<size_t>(slice) * per;
    const size_t v1 = v0 + per < vec_per_frame ? v0 + per : vec_per_frame;
    const uint4* a = frames + static_cast<size_t>(f) * vec_per_frame;
    const uint4* b = frames + static_cast<size_t>(c) * vec_per_frame;
    int diff = 0;
    for (size_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
        const uint4 x = ld_stream16(a + v), y = ld_stream16(b + v);
        diff |= (x.x != y.x) | (x.y != y.y) | (x.z != y.z) | (x.w != y.w);
    }
    if (__any_sync(0xffffffffu, diff) && (threadIdx.x & 31) == 0) atomicOr(differs + f, 1);
}

// one CTA of kPlanThreads: unique flags -> exclusive scan -> outputs
__global__ void __launch_bounds__(kPlanThreads)
frame_finalize_kernel(const int32_t* __restrict__ cand, const int32_t* __restrict__ differs, int n, int capacity,
                      int32_t* __restrict__ frame_map, int32_t* __restrict__ rep, int32_t* __restrict__ n_unique) {
    __shared__ int scan_smem[9];
    int base = 0;
    // pass 1: unique index of every distinct frame (frame_map of the distinct frames), rep list
    for (int i0 = 0; i0 < n; i0 += kPlanThreads) {
        const int i = i0 + threadIdx.x;
        const int uniq = (i < n) && (cand[i] == i || differs[i] != 0);
        int total;
        const int u = base + block_excl_scan(uniq, &total, scan_smem);
        if (uniq) {
            frame_map[i] = u;
            rep[u] = i;
        }
        base += total;
    }
    __syncthreads();   // frame_map of the distinct frames is visible to the whole CTA (global writes + barrier)
    // pass 2: duplicates point at their candidate's unique index (a candidate is always distinct: it is the FIRST frame
    // with that checksum, so its own candidate is itself)
    for (int i = threadIdx.x; i < n; i += kPlanThreads) {
        const int c = cand[i];
        if (c != i && differs[i] == 0) frame_map[i] = frame_map[c];
        if (i >= base) rep[i] = 0;
    }
    if (capacity > 0 && base > capacity) {
        // the caller sized its buffers for `capacity` distinct frames and there are more: keep every index in bounds
        // (the results of this batch are wrong -- the caller sees *n_unique > capacity)
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += kPlanThreads)
            if (frame_map[i] >= capacity) frame_map[i] = capacity - 1;
    }
    if (threadIdx.x == 0) *n_unique = base;
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint4* __restrict__ src, size_t vec_per_row, int n_src, const int32_t* __restrict__ idx,
                   uint4* __restrict__ dst) {
    const int r = blockIdx.y;
    int s = idx[r];
    s = s < 0 ? 0 : (s >= n_src ? n_src - 1 : s);
    const uint4* a = src + static_cast<size_t>(s) * vec_per_row;
    uint4* b = dst + static_cast<size_t>(r) * vec_per_row;
    for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < vec_per_row;
         v += static_cast<size_t>(gridDim.x) * blockDim.x)
        st_stream16(b + v, ld_stream16(a + v));
}

}  // namespace dedup
}  // namespace hvlm

extern "C" size_t hvlm_frame_dedup_workspace_bytes(int n_frames) {
    return n_frames > 0 ? hvlm::dedup::ws_bytes(n_frames) : 0;
}

extern "C" int hvlm_frame_dedup(const void* frames, size_t frame_bytes, int n_frames, int capacity, int32_t* frame_map,
                                int32_t* rep, int32_t* n_unique, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace hvlm;
    using namespace hvlm::dedup;
    if (!frames || !frame_map || !rep || !n_unique || !workspace || n_frames <= 0 || frame_bytes == 0 || capacity < 0)
        return HVLM_ERR_BAD_ARG;
    if (frame_bytes % 16 != 0) return HVLM_ERR_BAD_SHAPE;
    if (!aligned16(frames) || !aligned16(workspace)) return HVLM_ERR_ALIGN;
    if (workspace_bytes < ws_bytes(n_frames)) return HVLM_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const Ws w = carve(workspace, n_frames);
    if (cudaMemsetAsync(workspace, 0, ws_bytes(n_frames), s) != cudaSuccess) return HVLM_ERR_CUDA;
    const size_t vec = frame_bytes / 16;
    StageTimer st(HVLM_STAGE_OTHER, s);
    const dim3 grid(kSlices, n_frames);
    frame_hash_kernel<<<grid, kThreads, 0, s>>>(static_cast<const uint4*>(frames), vec, w.hash);
    int rc = check_last("frame_hash");
    if (rc) return rc;
    frame_confirm_kernel<<<grid, kThreads, 0, s>>>(static_cast<const uint4*>(frames), vec, w.hash, w.cand, w.differs);
    rc = check_last("frame_confirm");
    if (rc) return rc;
    frame_finalize_kernel<<<1, kPlanThreads, 0, s>>>(w.cand, w.differs, n_frames, capacity, frame_map, rep, n_unique);
    return check_last("frame_finalize");
}

extern "C" int hvlm_gather_rows(const void* src, size_t row_bytes, int n_src, const int32_t* idx, int n_out, void* dst,
                                void* stream) {
    using namespace hvlm;
    if (!src || !idx || !dst || n_src <= 0 || n_out <= 0 || row_bytes == 0) return HVLM_ERR_BAD_ARG;
    if (row_bytes % 16 != 0) return HVLM_ERR_BAD_SHAPE;
    if (!aligned16(src) || !aligned16(dst)) return HVLM_ERR_ALIGN;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t vec = row_bytes / 16;
    int gx = static_cast<int>((vec + 256 * 8 - 1) / (256 * 8));    // ~8 vectors per thread
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    StageTimer st(HVLM_STAGE_OTHER, s);
    dedup::gather_rows_kernel<<<dim3(gx, n_out), 256, 0, s>>>(static_cast<const uint4*>(src), vec, n_src, idx,
                                                              static_cast<uint4*>(dst));
    return check_last("gather_rows");
}
