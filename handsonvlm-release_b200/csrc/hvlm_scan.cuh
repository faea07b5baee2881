// Block-wide exclusive scan used by the index-plan kernels (256 threads per CTA).
#pragma once

namespace hvlm {

constexpr int kPlanThreads = 256;

// block-wide exclusive scan of one int per thread (256 threads); returns exclusive prefix, *total = block sum
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* smem /*[9]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();   // protect smem reuse across calls
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (kPlanThreads / 32) ? smem[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        if (lane < (kPlanThreads / 32)) smem[lane] = winc - w;   // exclusive warp offsets
        if (lane == (kPlanThreads / 32) - 1) smem[8] = winc;
    }
    __syncthreads();
    *total = smem[8];
    return smem[warp] + inc - v;
}

}  // namespace hvlm
