// (v1, kept as an A/B reference: HVLM_ATTN_V1=1 selects it)
// ViT-L/14 self-attention on tcgen05 / TMEM for sm_100a:  out = softmax(q k^T) v  per (frame, head),
// 257 tokens x 64 dims, no mask, no dropout (HF CLIPAttention in eval; q already carries the 64^-1/2 scale).
//
// One persistent CTA per SM walks (frame, head) items.  Per item:
//   TMA   : Q, K, V head slices [257x64] straight out of the QKV GEMM's [M,3072] output through ONE 4-D tensor
//           map (d, token, column block, frame): Q -> 3 row tiles of 128, K and V -> [272x64]; rows >= 257 are
//           zero-filled by the TMA unit.  Q/K/P are K-major SWIZZLE_128B operands; V is consumed as an
//           MN-major B operand (no transpose anywhere).
//   MMA 1 : S[128 x 272] = Q_tile K^T     tcgen05.mma M=128, N=256 (+ N=16 for keys 256..271), K=16 x 4
//           fp32 accumulator in TMEM columns [0,272)
//   softmax (4 warps, one row per thread): two passes over the TMEM row (max, then exp/sum in fp32);
//           un-normalised P is rounded to bf16 and written to shared memory in the swizzled K-major layout
//   MMA 2 : O[128 x 64] = P V             17 k-steps of 16 keys, accumulator in TMEM columns [320,384)
//   epilogue: O / rowsum -> bf16 -> out[(frame*257+tok), head*64 .. +63]  (A operand of out_proj)
// The S-MMA of tile i+1 is issued right behind the PV-MMA of tile i, and the next item's Q/K (V^T) loads
// are issued as soon as the last S (PV) MMA of the current item has retired, so TMA latency hides behind
// the softmax / epilogue.
#include "hvlm_internal.cuh"
#include "hvlm_ptx.cuh"

namespace hvlm {
namespace attn_v1 {

constexpr int kAttnV1Threads = 160;          // warp 0: TMA + MMA issue + TMEM alloc; warps 1..4: softmax/epilogue
constexpr int kS = HVLM_VIT_TOKENS;        // 257
constexpr int kSK = 272;                   // keys padded to a multiple of 16
constexpr int kQTile = 128 * 64 * 2;       // 16384
constexpr int kQBytes = 3 * kQTile;        // 49152
constexpr int kKBytes = kSK * 64 * 2;      // 34816
constexpr int kVBytes = kSK * 64 * 2;      // 34816
constexpr int kPBlock = 128 * 64 * 2;      // 16384
constexpr int kPBytes = 5 * kPBlock;       // 81920
constexpr int kAttnSmem = kQBytes + kKBytes + kVBytes + kPBytes + 1024 + 128;
constexpr int kOCol = 320;                 // TMEM column of the O accumulator
constexpr uint32_t kQKTx = kQBytes + 2 * kQTile + 16 * 128;   // Q (3 boxes) + K (2 boxes of 128 rows + 16 rows)
constexpr uint32_t kVTx = 2 * kQTile + 16 * 128;              // V (2 boxes of 128 rows + 16 rows)

__global__ void __launch_bounds__(kAttnV1Threads, 1)
attn_v1_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_tail,
                    __nv_bfloat16* __restrict__ out, int n_items) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kQBytes;
    uint8_t* sV = sK + kKBytes;
    uint8_t* sP = sV + kVBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes);
    uint64_t* qk_full = bars + 0;
    uint64_t* v_full = bars + 1;
    uint64_t* s_full = bars + 2;
    uint64_t* p_full = bars + 3;
    uint64_t* o_full = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

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
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            constexpr uint32_t idesc_s256 = umma_idesc_bf16(128, 256);
            constexpr uint32_t idesc_s16 = umma_idesc_bf16(128, 16);
            constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, /*b_mn_major=*/1);
            const uint64_t dQ = umma_desc_k_sw128(smem_u32(sQ));
            const uint64_t dK = umma_desc_k_sw128(smem_u32(sK));
            const uint64_t dV = umma_desc_k_sw128(smem_u32(sV));
            const uint64_t dP = umma_desc_k_sw128(smem_u32(sP));

            // item = frame*16 + head; column blocks of the [M,3072] QKV matrix: q -> head, k -> 16+head, v -> 32+head
            auto load_qk = [&](int item) {
                const int f = item >> 4, h = item & 15;
                mbar_arrive_expect_tx(qk_full, kQKTx);
                tma_load_4d(sQ, &tm_qkv, qk_full, 0, 0, h, f);
                tma_load_4d(sQ + kQTile, &tm_qkv, qk_full, 0, 128, h, f);
                tma_load_4d(sQ + 2 * kQTile, &tm_qkv, qk_full, 0, 256, h, f);
                tma_load_4d(sK, &tm_qkv, qk_full, 0, 0, 16 + h, f);
                tma_load_4d(sK + kQTile, &tm_qkv, qk_full, 0, 128, 16 + h, f);
                tma_load_4d(sK + 2 * kQTile, &tm_tail, qk_full, 0, 256, 16 + h, f);
            };
            auto load_v = [&](int item) {
                const int f = item >> 4, h = item & 15;
                mbar_arrive_expect_tx(v_full, kVTx);
                tma_load_4d(sV, &tm_qkv, v_full, 0, 0, 32 + h, f);
                tma_load_4d(sV + kQTile, &tm_qkv, v_full, 0, 128, 32 + h, f);
                tma_load_4d(sV + 2 * kQTile, &tm_tail, v_full, 0, 256, 32 + h, f);
            };
            auto issue_s = [&](int tile) {
                // S = Q_tile K^T : 4 k-steps over the 64 head dims
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t a = dQ + static_cast<uint64_t>((tile * kQTile) >> 4) + static_cast<uint64_t>(2 * k);
                    umma_bf16_ss(tmem_base, a, dK + static_cast<uint64_t>(2 * k), idesc_s256, k > 0);
                    umma_bf16_ss(tmem_base + 256, a, dK + static_cast<uint64_t>((256 * 128) >> 4) + static_cast<uint64_t>(2 * k),
                                 idesc_s16, k > 0);
                }
                umma_commit(s_full);
            };

            int it = 0;
            uint32_t n3 = 0;   // running tile counter (s_full / p_full / o_full complete once per tile)
            if (static_cast<int>(blockIdx.x) < n_items) {
                load_qk(blockIdx.x);
                load_v(blockIdx.x);
            }
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int next = item + gridDim.x;
                mbar_wait(qk_full, it & 1);
                tc_fence_after();
                issue_s(0);
                for (int tile = 0; tile < 3; ++tile, ++n3) {
                    if (tile == 2) {
                        // all S MMAs of this item have retired once s_full(tile 2) fires: Q/K smem is free
                        mbar_wait(s_full, n3 & 1);
                        if (next < n_items) load_qk(next);
                    }
                    mbar_wait(p_full, n3 & 1);
                    if (tile == 0) mbar_wait(v_full, it & 1);
                    tc_fence_after();
                    // O = P V : 17 k-steps of 16 keys (272 = keys padded; P and V pad rows are zero).
                    // A = P (K-major: +32 B per step inside a 64-key block); B = V rows [key, d] as an MN-major
                    // operand: 16 keys = 16 rows of 128 B = +2048 B per step.
#pragma unroll
                    for (int kk = 0; kk < 17; ++kk) {
                        const uint32_t blk = kk >> 2, sub = kk & 3;
                        umma_bf16_ss(tmem_base + kOCol, dP + static_cast<uint64_t>((blk * kPBlock) >> 4) + 2 * sub,
                                     dV + static_cast<uint64_t>((kk * 2048) >> 4), idesc_o, kk > 0);
                    }
                    umma_commit(o_full);
                    if (tile < 2) issue_s(tile + 1);
                }
                // PV of the last tile retired -> V^T smem is free
                mbar_wait(o_full, (n3 - 1) & 1);
                if (next < n_items) load_v(next);
            }
        }
    } else {
        // ===================== softmax + epilogue warps =====================
        const int q = warp & 3;                       // TMEM lane quarter
        const int r = q * 32 + lane;                  // row inside the 128-row tile
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint32_t n3 = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int f = item >> 4, head = item & 15;
            for (int tile = 0; tile < 3; ++tile, ++n3) {
                const int tok = tile * 128 + r;
                const bool warp_active = (tile * 128 + q * 32) < kS;   // warp-uniform
                mbar_wait(s_full, n3 & 1);
                tc_fence_after();
                float inv_sum = 0.f;
                if (warp_active) {
                    // ---- pass 1: row max over the 257 valid keys
                    float mx = -INFINITY;
#pragma unroll 1
                    for (int c = 0; c < 8; ++c) {
                        uint32_t v[32];
                        tmem_ld32(t_lane + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
                    }
                    {
                        uint32_t v[16];
                        tmem_ld16(t_lane + 256, v);
                        tmem_ld_wait();
                        mx = fmaxf(mx, __uint_as_float(v[0]));      // key 256; 257..271 are padding
                    }
                    // ---- pass 2: p = exp(s - max), row sum, bf16 P into swizzled smem
                    float sum = 0.f;
                    uint8_t* prow = sP + r * 128;
                    const int sw = r & 7;
#pragma unroll 1
                    for (int c = 0; c < 8; ++c) {
                        uint32_t v[32];
                        tmem_ld32(t_lane + c * 32, v);
                        tmem_ld_wait();
                        float p[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            p[j] = __expf(__uint_as_float(v[j]) - mx);
                            sum += p[j];
                        }
                        uint8_t* blk = prow + (c >> 1) * kPBlock;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int chunk = ((c & 1) * 4 + j) ^ sw;
                            uint4 w;
                            w.x = pack_bf16(p[8 * j + 0], p[8 * j + 1]);
                            w.y = pack_bf16(p[8 * j + 2], p[8 * j + 3]);
                            w.z = pack_bf16(p[8 * j + 4], p[8 * j + 5]);
                            w.w = pack_bf16(p[8 * j + 6], p[8 * j + 7]);
                            *reinterpret_cast<uint4*>(blk + chunk * 16) = w;
                        }
                    }
                    {
                        uint32_t v[16];
                        tmem_ld16(t_lane + 256, v);
                        tmem_ld_wait();
                        const float p0 = __expf(__uint_as_float(v[0]) - mx);
                        sum += p0;
                        uint8_t* blk = prow + 4 * kPBlock;
                        uint4 w0 = make_uint4(pack_bf16(p0, 0.f), 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(blk + ((0 ^ sw) * 16)) = w0;
                        *reinterpret_cast<uint4*>(blk + ((1 ^ sw) * 16)) = make_uint4(0u, 0u, 0u, 0u);
                    }
                    inv_sum = 1.0f / sum;
                }
                // make the generic-proxy smem writes visible to the tensor core (async proxy), then signal
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(p_full);

                // ---- epilogue: O / rowsum -> bf16
                mbar_wait(o_full, n3 & 1);
                tc_fence_after();
                if (warp_active) {
                    uint32_t o0[32], o1[32];
                    tmem_ld32(t_lane + kOCol, o0);
                    tmem_ld32(t_lane + kOCol + 32, o1);
                    tmem_ld_wait();
                    if (tok < kS) {
                        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(f) * kS + tok) * 1024 + head * 64);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 w;
                            w.x = pack_bf16(__uint_as_float(o0[8 * j + 0]) * inv_sum, __uint_as_float(o0[8 * j + 1]) * inv_sum);
                            w.y = pack_bf16(__uint_as_float(o0[8 * j + 2]) * inv_sum, __uint_as_float(o0[8 * j + 3]) * inv_sum);
                            w.z = pack_bf16(__uint_as_float(o0[8 * j + 4]) * inv_sum, __uint_as_float(o0[8 * j + 5]) * inv_sum);
                            w.w = pack_bf16(__uint_as_float(o0[8 * j + 6]) * inv_sum, __uint_as_float(o0[8 * j + 7]) * inv_sum);
                            dst[j] = w;
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 w;
                            w.x = pack_bf16(__uint_as_float(o1[8 * j + 0]) * inv_sum, __uint_as_float(o1[8 * j + 1]) * inv_sum);
                            w.y = pack_bf16(__uint_as_float(o1[8 * j + 2]) * inv_sum, __uint_as_float(o1[8 * j + 3]) * inv_sum);
                            w.z = pack_bf16(__uint_as_float(o1[8 * j + 4]) * inv_sum, __uint_as_float(o1[8 * j + 5]) * inv_sum);
                            w.w = pack_bf16(__uint_as_float(o1[8 * j + 6]) * inv_sum, __uint_as_float(o1[8 * j + 7]) * inv_sum);
                            dst[4 + j] = w;
                        }
                    }
                }
                tc_fence_before();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

}  // namespace attn_v1

int launch_attention_v1(const void* qkv, void* out, int n_frames, cudaStream_t s) {
    using namespace attn_v1;
    const int n_items = n_frames * HVLM_VIT_HEADS;
    CUtensorMap tq, tt;
    {
        // qkv [n_frames*257, 3072] bf16 viewed as (d:64, token:257, column block:48, frame)
        uint64_t dims[4] = {64, static_cast<uint64_t>(kS), 48, static_cast<uint64_t>(n_frames)};
        uint64_t str[3] = {3072 * 2, 128, static_cast<uint64_t>(kS) * 3072 * 2};
        uint32_t box[4] = {64, 128, 1, 1};
        uint32_t box_tail[4] = {64, 16, 1, 1};
        int rc = make_tmap_bf16(&tq, qkv, 4, dims, str, box);
        if (rc) return rc;
        rc = make_tmap_bf16(&tt, qkv, 4, dims, str, box_tail);
        if (rc) return rc;
    }
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(attn_v1_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem) != cudaSuccess)
            return HVLM_ERR_CUDA;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int grid = n_items < num_sms() ? n_items : num_sms();
    attn_v1_tcgen05_kernel<<<grid, kAttnV1Threads, kAttnSmem, s>>>(tq, tt, static_cast<__nv_bfloat16*>(out), n_items);
    return check_last("attention");
}

}  // namespace hvlm

