// ViT-L/14 self-attention on tcgen05 / TMEM for sm_100a:  out = softmax(q k^T) v  per (frame, head),
// 257 tokens x 64 dims, no mask, no dropout (HF CLIPAttention in eval; q already carries the 64^-1/2 scale).
//
// Input  : qkv bf16, COLUMN-BLOCK-MAJOR [48 column blocks: q0..q15 | k0..k15 | v0..v15][M = frames*257][64] as written by
//          the QKV GEMM's epilogue -- every (frame, head) operand is one contiguous 32.9 KB block.
// Output : bf16 [frame*257 + token][head*64 + d]  (the K-major A operand of out_proj).
//
// Design (sized so that TWO CTAs are resident per SM: one CTA's MMAs overlap the other's softmax):
//   * 257 = 256 + 1.  The tensor cores handle the 256 x 256 block (queries/keys 0..255) as two 128-row tiles.
//     The 257th KEY is folded in on CUDA cores by the softmax threads (one 64-MAC dot product per row and one extra
//     term in max / sum / output).  The 257th QUERY row is computed on CUDA cores: its 257 scores by the softmax
//     threads (two dot products each, in the shadow of the first S MMA), its softmax and P*V by a dedicated warp.
//   * TMEM: 256 columns per CTA.  S[128x256] fp32 fills them; the un-normalised probabilities are written back
//     IN PLACE as packed bf16 (P aliases columns 0..127, tcgen05.st) and feed the second MMA as a TMEM A-operand;
//     O[128x64] accumulates in columns 128..191.
//   * smem (109 KB): Q (2 tiles) | K[256x64] | V[256x64] | row 256 of q,k,v | tail-query scratch | per-warp store
//     staging.  V is consumed as an MN-major B operand (no transpose anywhere).
//   * warps: 0 = TMA + MMA issue (warp-uniform control flow, one elected lane issues), 1..4 = softmax + epilogue
//     (one query row per thread, fp32), 5 = tail query.  Output rows are transposed through a small per-warp smem
//     buffer so that every global store instruction writes four full 128-byte lines.
//   * next item's Q/K (V) loads are issued as soon as the current item's last S (PV) MMA has retired and the
//     CUDA-core readers of that buffer have signalled.
#include "hvlm_internal.cuh"
#include "hvlm_ptx.cuh"

namespace hvlm {
namespace attn {

constexpr int kThreads = 192;
constexpr int kS = HVLM_VIT_TOKENS;        // 257
constexpr int kTile = 128 * 64 * 2;        // 16384 : one [128 x 64] bf16 operand tile
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + 2 * kTile;
constexpr int kOffV = kOffK + 2 * kTile;
constexpr int kOffQT = kOffV + 2 * kTile;  // q row 256 (128 B used; 1 KB slot keeps the swizzle phase at 0)
constexpr int kOffKT = kOffQT + 1024;      // k row 256
constexpr int kOffVT = kOffKT + 1024;      // v row 256
constexpr int kOffStage = kOffVT + 1024;   // 4 warps x 2 KB output staging
constexpr int kOffPT = kOffStage + 4 * 2048;   // 2 x float[272]: scores / probabilities of the tail query
constexpr int kOffBar = kOffPT + 2 * 272 * 4;
constexpr int kSmem = kOffBar + 128 + 1024;    // + barriers + alignment slack
constexpr uint32_t kQKTx = 4 * kTile + 2 * 128;
constexpr uint32_t kVTx = 2 * kTile + 128;
constexpr int kPCol = 0;                   // P (bf16 pairs) : TMEM columns [0,128)
constexpr int kOCol = 128;                 // O accumulator  : TMEM columns [128,192)

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 8 bf16 (one 16-byte chunk) -> 8 floats
__device__ __forceinline__ void unpack8(const uint4& w, float* f) {
    const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(u[i] << 16);
        f[2 * i + 1] = __uint_as_float(u[i] & 0xFFFF0000u);
    }
}
// dot product of row `row` of a SWIZZLE_128B [rows x 64] bf16 tile with a 64-float vector held in registers
__device__ __forceinline__ float dot_row64(const uint8_t* tile, int row, const float (&vec)[64]) {
    const uint8_t* base = tile + row * 128;
    const int sw = row & 7;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(base + ((j ^ sw) << 4)), f);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            acc0 = fmaf(f[i], vec[8 * j + i], acc0);
            acc1 = fmaf(f[i + 1], vec[8 * j + i + 1], acc1);
        }
    }
    return acc0 + acc1;
}
// a single 128-byte row stored at a 1024-aligned address (swizzle phase 0 -> chunks in natural order) -> 64 floats
__device__ __forceinline__ void load_row0(const uint8_t* row, float (&vec)[64]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) unpack8(*reinterpret_cast<const uint4*>(row + (j << 4)), &vec[8 * j]);
}

// debug: per-phase timestamps (trace == nullptr in production)
#define ATTN_TRACE(role, slot)                                                                          \
    do {                                                                                                \
        if (trace != nullptr && it < 4)                                                                 \
            trace[((static_cast<size_t>(blockIdx.x) * 2 + (role)) * 4 + it) * 16 + (slot)] = clock64();   \
    } while (0)

__global__ void __launch_bounds__(kThreads, 2)
attn_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_tail,
                    __nv_bfloat16* __restrict__ out, int n_items, long long* __restrict__ trace) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint8_t* sQ = smem + kOffQ;
    uint8_t* sK = smem + kOffK;
    uint8_t* sV = smem + kOffV;
    uint8_t* sQT = smem + kOffQT;
    uint8_t* sKT = smem + kOffKT;
    uint8_t* sVT = smem + kOffVT;
    uint8_t* sStage = smem + kOffStage;
    float* sPT = reinterpret_cast<float*>(smem + kOffPT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint64_t* qk_full = bars + 0;    // TMA  -> everyone        (once per item)
    uint64_t* v_full = bars + 1;     // TMA  -> MMA, tail warp  (once per item)
    uint64_t* s_full = bars + 2;     // MMA  -> softmax         (once per tile)
    uint64_t* p_full = bars + 3;     // softmax(128) -> MMA     (once per tile)
    uint64_t* o_full = bars + 4;     // MMA  -> softmax         (once per tile)
    uint64_t* o_read = bars + 5;     // softmax(128) -> MMA     (once per tile): O left TMEM, S region reusable
    uint64_t* q_read = bars + 6;     // softmax(128) -> MMA, tail warp (once per item): Q/K rows consumed by the
                                     //   CUDA cores, tail-query scores published
    uint64_t* tv_done = bars + 7;    // tail warp -> MMA        (once per item): V consumed by the tail warp
    uint64_t* vt_read = bars + 8;    // softmax(128) -> MMA     (once per item): v row 256 consumed by the epilogues
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_qkv);
            tma_prefetch_desc(&tm_tail);
            mbar_init(qk_full, 1);
            mbar_init(v_full, 1);
            mbar_init(s_full, 1);
            mbar_init(p_full, 128);
            mbar_init(o_full, 1);
            mbar_init(o_read, 128);
            mbar_init(q_read, 128);
            mbar_init(tv_done, 1);
            mbar_init(vt_read, 128);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<256>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();              // everything above overlapped the previous kernel's tail

    if (warp == 0) {
        // ===================== TMA producer + MMA issuer =====================
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 256);
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, /*b_mn_major=*/1);
        const uint64_t dQ = umma_desc_k_sw128(smem_u32(sQ));
        const uint64_t dK = umma_desc_k_sw128(smem_u32(sK));
        const uint64_t dV = umma_desc_k_sw128(smem_u32(sV));

        // item = frame*16 + head; column blocks: q -> head, k -> 16+head, v -> 32+head
        auto load_qk = [&](int item) {
            const int h = item & 15, r0 = (item >> 4) * kS;   // first row of the frame in the [M] dimension
            if (elect_one()) {
                mbar_arrive_expect_tx(qk_full, kQKTx);
                tma_load_3d(sQ, &tm_qkv, qk_full, 0, r0, h);
                tma_load_3d(sQ + kTile, &tm_qkv, qk_full, 0, r0 + 128, h);
                tma_load_3d(sQT, &tm_tail, qk_full, 0, r0 + 256, h);
                tma_load_3d(sK, &tm_qkv, qk_full, 0, r0, 16 + h);
                tma_load_3d(sK + kTile, &tm_qkv, qk_full, 0, r0 + 128, 16 + h);
                tma_load_3d(sKT, &tm_tail, qk_full, 0, r0 + 256, 16 + h);
            }
            __syncwarp();
        };
        auto load_v = [&](int item) {
            const int h = item & 15, r0 = (item >> 4) * kS;
            if (elect_one()) {
                mbar_arrive_expect_tx(v_full, kVTx);
                tma_load_3d(sV, &tm_qkv, v_full, 0, r0, 32 + h);
                tma_load_3d(sV + kTile, &tm_qkv, v_full, 0, r0 + 128, 32 + h);
                tma_load_3d(sVT, &tm_tail, v_full, 0, r0 + 256, 32 + h);
            }
            __syncwarp();
        };

        int it = 0;
        uint32_t n2 = 0;   // running tile counter (s_full / p_full / o_full / o_read complete once per tile)
        // items are walked last-to-first: the QKV GEMM wrote the last frames last, so they are still in L2
        if (static_cast<int>(blockIdx.x) < n_items) {
            load_qk(n_items - 1 - blockIdx.x);
            load_v(n_items - 1 - blockIdx.x);
        }
        for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++it) {
            const int next_idx = idx + gridDim.x;
            const int next = n_items - 1 - next_idx;
            if (lane == 0) ATTN_TRACE(0, 0);
            mbar_wait(qk_full, it & 1);
            if (lane == 0) ATTN_TRACE(0, 1);
            for (int tile = 0; tile < 2; ++tile, ++n2) {
                // the S region doubles as P / O storage of the previous tile: wait until its O has been read
                if (n2 > 0) mbar_wait(o_read, (n2 - 1) & 1);
                if (lane == 0) ATTN_TRACE(0, 2 + tile * 6);
                tc_fence_after();
                // S = Q_tile K^T : 4 k-steps over the 64 head dims, N = 256 keys
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_base, dQ + static_cast<uint64_t>((tile * kTile) >> 4) + static_cast<uint64_t>(2 * k),
                                     dK + static_cast<uint64_t>(2 * k), idesc_s, k > 0);
                    umma_commit(s_full);
                }
                __syncwarp();
                if (tile == 1) {
                    // Q / K smem is free once the last S MMA has retired and the CUDA-core readers are done
                    mbar_wait(s_full, n2 & 1);
                    mbar_wait(q_read, it & 1);
                    if (next_idx < n_items) load_qk(next);
                }
                if (lane == 0) ATTN_TRACE(0, 3 + tile * 6);
                mbar_wait(p_full, n2 & 1);
                if (lane == 0) ATTN_TRACE(0, 4 + tile * 6);
                if (tile == 0) mbar_wait(v_full, it & 1);
                tc_fence_after();
                // O = P V : A = P from TMEM (16 keys = 8 columns per step), B = V rows as MN-major operand
                // (16 keys = 16 rows of 128 B = +2048 B per step)
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 16; ++kk)
                        umma_bf16_ts(tmem_base + kOCol, tmem_base + kPCol + kk * 8,
                                     dV + static_cast<uint64_t>((kk * 2048) >> 4), idesc_o, kk > 0);
                    umma_commit(o_full);
                }
                __syncwarp();
                if (lane == 0) ATTN_TRACE(0, 5 + tile * 6);
            }
            // PV of the last tile retired and the CUDA-core readers are done with V -> V smem is free
            mbar_wait(o_full, (n2 - 1) & 1);
            mbar_wait(tv_done, it & 1);
            mbar_wait(vt_read, it & 1);
            if (lane == 0) ATTN_TRACE(0, 14);
            if (next_idx < n_items) load_v(next);
        }
    } else if (warp <= 4) {
        // ===================== softmax + epilogue warps: one query row per thread =====================
        const int q = warp & 3;                       // TMEM lane quarter
        const int r = q * 32 + lane;                  // row inside the 128-row tile
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint8_t* stage = sStage + (warp - 1) * 2048;  // this warp's 16-row x 128-byte transpose buffer
        constexpr float kLog2e = 1.4426950408889634f;
        uint32_t n2 = 0;
        int it = 0;
        for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++it) {
            const int item = n_items - 1 - idx;
            const int f = item >> 4, head = item & 15;
            const bool tr = (r == 0);
            if (tr) ATTN_TRACE(1, 0);
            mbar_wait(qk_full, it & 1);
            if (tr) ATTN_TRACE(1, 1);
            // ---- CUDA-core side work, in the shadow of the first S MMA:
            //      (a) this thread's two query rows against the 257th key, (b) the 257th query against this thread's
            //      two key rows (scores of the tail query, published through smem to the tail warp)
            float s_tail[2];
            {
                float vec[64];
                load_row0(sKT, vec);
                s_tail[0] = dot_row64(sQ, r, vec);
                s_tail[1] = dot_row64(sQ + kTile, r, vec);
                float* pt = sPT + (it & 1) * 272;
                if (tr) {   // q_256 . k_256 needs both tail rows: do it before vec is overwritten
                    float kt_dot = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float qv[8];
                        unpack8(*reinterpret_cast<const uint4*>(sQT + (j << 4)), qv);
#pragma unroll
                        for (int i = 0; i < 8; ++i) kt_dot = fmaf(qv[i], vec[8 * j + i], kt_dot);
                    }
                    pt[256] = kt_dot;
                }
                load_row0(sQT, vec);
                pt[r] = dot_row64(sK, r, vec);
                pt[r + 128] = dot_row64(sK + kTile, r, vec);
            }
            mbar_arrive(q_read);
            for (int tile = 0; tile < 2; ++tile, ++n2) {
                if (tr) ATTN_TRACE(1, 2 + tile * 7);
                mbar_wait(s_full, n2 & 1);
                if (tr) ATTN_TRACE(1, 3 + tile * 7);
                tc_fence_after();
                // ---- pass 1: row max over 256 keys (TMEM) and the tail key
                float mx = s_tail[tile];
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    uint32_t v[32];
                    tmem_ld32(t_lane + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
                }
                if (tr) ATTN_TRACE(1, 4 + tile * 7);
                // ---- pass 2: p = exp(s - max) (fp32), row sum, P -> packed bf16 back into TMEM (aliases S)
                const float mxl = mx * kLog2e;
                float sum = 0.f;
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    uint32_t v[32];
                    tmem_ld32(t_lane + c * 32, v);
                    tmem_ld_wait();
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float p0 = fast_exp2(fmaf(__uint_as_float(v[2 * j]), kLog2e, -mxl));
                        const float p1 = fast_exp2(fmaf(__uint_as_float(v[2 * j + 1]), kLog2e, -mxl));
                        sum += p0 + p1;
                        pk[j] = pack_bf16(p0, p1);
                    }
                    tmem_st16(t_lane + kPCol + c * 16, pk);
                }
                const float p_tail = fast_exp2(fmaf(s_tail[tile], kLog2e, -mxl));
                sum += p_tail;
                const float inv_sum = 1.0f / sum;
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(p_full);
                if (tr) ATTN_TRACE(1, 5 + tile * 7);

                // ---- epilogue: (O + p_tail * v_256) / rowsum -> bf16
                mbar_wait(o_full, n2 & 1);
                if (tr) ATTN_TRACE(1, 6 + tile * 7);
                tc_fence_after();
                uint32_t o0[32], o1[32];
                tmem_ld32(t_lane + kOCol, o0);
                tmem_ld32(t_lane + kOCol + 32, o1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(o_read);
                if (tr) ATTN_TRACE(1, 7 + tile * 7);
                uint4 rowv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float vt[8];
                    unpack8(*reinterpret_cast<const uint4*>(sVT + (j << 4)), vt);
                    const uint32_t* o = j < 4 ? &o0[8 * j] : &o1[8 * (j - 4)];
                    float y[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[i] = fmaf(p_tail, vt[i], __uint_as_float(o[i])) * inv_sum;
                    rowv[j].x = pack_bf16(y[0], y[1]);
                    rowv[j].y = pack_bf16(y[2], y[3]);
                    rowv[j].z = pack_bf16(y[4], y[5]);
                    rowv[j].w = pack_bf16(y[6], y[7]);
                }
                // transpose through the per-warp staging buffer (16 rows at a time) so that each global store
                // instruction writes 4 complete 128-byte rows: lane L -> row L/8 + 4i, 16-byte chunk L%8
                const size_t row0 = static_cast<size_t>(f) * kS + tile * 128 + q * 32;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if ((lane >> 4) == half) {
                        uint8_t* srow = stage + (lane & 15) * 128;
#pragma unroll
                        for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(srow + ((j ^ (lane & 7)) << 4)) = rowv[j];
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = (lane >> 3) + 4 * i;
                        const uint4 w = *reinterpret_cast<const uint4*>(stage + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
                        *reinterpret_cast<uint4*>(out + (row0 + half * 16 + rr) * 1024 + head * 64 + (lane & 7) * 8) = w;
                    }
                    __syncwarp();
                }
                if (tr) ATTN_TRACE(1, 8 + tile * 7);
            }
            mbar_arrive(vt_read);
        }
    } else {
        // ===================== warp 5: softmax and P*V of the 257th query row on CUDA cores =====================
        const int g = lane >> 3;      // key group: keys 64g .. 64g+63
        const int c = lane & 7;       // 16-byte chunk of the head dim: dims 8c .. 8c+7
        int it = 0;
        for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++it) {
            const int item = n_items - 1 - idx;
            const int f = item >> 4, head = item & 15;
            float* pt = sPT + (it & 1) * 272;
            mbar_wait(q_read, it & 1);          // scores published by the softmax threads
            float sc[8];
            float mx = pt[256];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                sc[j] = pt[lane + 32 * j];
                mx = fmaxf(mx, sc[j]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float p256 = __expf(pt[256] - mx);
            __syncwarp();                        // every lane has read pt[256] / its scores before p overwrites them
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float p = __expf(sc[j] - mx);
                sum += p;
                pt[lane + 32 * j] = p;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            sum += p256;
            __syncwarp();
            mbar_wait(v_full, it & 1);
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 4
            for (int kk = 0; kk < 64; ++kk) {
                const int key = g * 64 + kk;
                const uint8_t* rowp = (key < 128 ? sV : sV + kTile) + (key & 127) * 128;
                float vv[8];
                unpack8(*reinterpret_cast<const uint4*>(rowp + ((c ^ (key & 7)) << 4)), vv);
                const float p = pt[key];
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(p, vv[i], acc[i]);
            }
            if (g == 0) {
                float vv[8];
                unpack8(*reinterpret_cast<const uint4*>(sVT + (c << 4)), vv);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(p256, vv[i], acc[i]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
                acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
            }
            if (g == 0) {
                const float inv = 1.0f / sum;
                uint4 w;
                w.x = pack_bf16(acc[0] * inv, acc[1] * inv);
                w.y = pack_bf16(acc[2] * inv, acc[3] * inv);
                w.z = pack_bf16(acc[4] * inv, acc[5] * inv);
                w.w = pack_bf16(acc[6] * inv, acc[7] * inv);
                *reinterpret_cast<uint4*>(out + (static_cast<size_t>(f) * kS + 256) * 1024 + head * 64 + c * 8) = w;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(tv_done);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem_base);
}

}  // namespace attn

// qkv bf16 column-block-major [48][n_rows][64] viewed by TMA as (d:64, row:n_rows, column block:48)
int make_qkv_hm_tmap(CUtensorMap* out, const void* qkv_hm, int n_rows, int box_rows) {
    uint64_t dims[3] = {64, static_cast<uint64_t>(n_rows), 48};
    uint64_t str[2] = {128, static_cast<uint64_t>(n_rows) * 128};
    uint32_t box[3] = {64, static_cast<uint32_t>(box_rows), 1};
    return make_tmap_bf16(out, qkv_hm, 3, dims, str, box);
}

static int launch_attention_impl(const void* qkv_hm, void* out, int n_frames, cudaStream_t s, long long* trace) {
    using namespace attn;
    const int n_items = n_frames * HVLM_VIT_HEADS;
    CUtensorMap tq, tt;
    int rc = make_qkv_hm_tmap(&tq, qkv_hm, n_frames * kS, 128);
    if (rc) return rc;
    rc = make_qkv_hm_tmap(&tt, qkv_hm, n_frames * kS, 1);
    if (rc) return rc;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(attn_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess)
            return HVLM_ERR_CUDA;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int max_ctas = 2 * num_sms();
    const int grid = n_items < max_ctas ? n_items : max_ctas;
    if (launch_pdl(attn_tcgen05_kernel, dim3(grid), dim3(kThreads), kSmem, s, tq, tt, static_cast<__nv_bfloat16*>(out), n_items,
                   trace) != cudaSuccess) {
        cudaGetLastError();
        return HVLM_ERR_CUDA;
    }
    return check_last("attention");
}

int launch_attention(const void* qkv_hm, void* out, int n_frames, cudaStream_t s) {
    return launch_attention_impl(qkv_hm, out, n_frames, s, nullptr);
}

}  // namespace hvlm

extern "C" int hvlm_vit_attention(const void* qkv_hm, void* out, int n_frames, void* stream) {
    using namespace hvlm;
    if (!qkv_hm || !out || n_frames <= 0) return HVLM_ERR_BAD_ARG;
    if (!aligned16(qkv_hm) || !aligned16(out)) return HVLM_ERR_ALIGN;
    StageTimer st(HVLM_STAGE_ATTENTION, static_cast<cudaStream_t>(stream));
    return launch_attention(qkv_hm, out, n_frames, static_cast<cudaStream_t>(stream));
}

// debug only (not part of the ABI header): trace[grid][2 roles][4 items][16 slots] clock64 timestamps
extern "C" __attribute__((visibility("default"))) int hvlm_debug_attention_trace(const void* qkv_hm, void* out,
                                                                                 int n_frames, long long* trace,
                                                                                 void* stream, int) {
    return hvlm::launch_attention_impl(qkv_hm, out, n_frames, static_cast<cudaStream_t>(stream), trace);
}
