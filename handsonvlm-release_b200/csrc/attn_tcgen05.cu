// ViT-L/14 self-attention v2 on tcgen05 / TMEM for sm_100a:  out = softmax(q k^T) v  per (frame, head),
// 257 tokens x 64 dims, no mask, no dropout (HF CLIPAttention in eval; q already carries the 64^-1/2 scale).
//
// Design (FA4-style, sized so that TWO CTAs are resident per SM and one CTA's MMAs overlap the other's softmax):
//   * 257 = 256 + 1.  The tensor cores handle the 256 x 256 block (queries/keys 0..255) as two 128-row tiles;
//     the 257th KEY is folded in on CUDA cores by the softmax threads (one 64-MAC dot product per row, one extra
//     term in max / sum / output), and the 257th QUERY row is computed entirely on CUDA cores by a dedicated warp.
//     No padded third M-tile, no padded key columns.
//   * TMEM: 256 columns per CTA.  S[128x256] fp32 fills them; after the softmax has consumed a chunk of S the
//     un-normalised probabilities are written back IN PLACE as packed bf16 (P aliases columns 0..127) with
//     tcgen05.st and feed the second MMA as a TMEM A-operand; O[128x64] accumulates in columns 128..191.
//   * smem (104 KB): Q (2 tiles) | K[256x64] | V[256x64] | row-256 tails of q,k,v | scratch; all TMA-loaded straight
//     from the QKV GEMM's [M,3072] output through one 4-D tensor map; V is consumed as an MN-major B operand.
//   * warps: 0 = TMA + MMA issue (one thread), 1..4 = softmax + epilogue (one query row per thread, fp32),
//     5 = the 257th query row.  Next item's Q/K (V) loads are issued as soon as the current item's last S (PV) MMA
//     has retired and the CUDA-core readers have signalled.
#include "hvlm_internal.cuh"
#include "hvlm_ptx.cuh"

namespace hvlm {
int launch_attention_v1(const void* qkv, void* out, int n_frames, cudaStream_t s);

namespace attn_v2 {

constexpr int kThreads = 192;
constexpr int kS = HVLM_VIT_TOKENS;        // 257
constexpr int kTile = 128 * 64 * 2;        // 16384 : one [128 x 64] bf16 operand tile
constexpr int kTail = 16 * 64 * 2;         // 2048  : rows 256..271 (only row 256 is real, the rest zero-filled)
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + 2 * kTile;
constexpr int kOffV = kOffK + 2 * kTile;
constexpr int kOffQT = kOffV + 2 * kTile;  // q row 256
constexpr int kOffKT = kOffQT + kTail;     // k row 256
constexpr int kOffVT = kOffKT + kTail;     // v row 256
constexpr int kOffPT = kOffVT + kTail;     // float[272] probabilities of the tail query
constexpr int kOffBar = kOffPT + 272 * 4;
constexpr int kSmem = kOffBar + 128 + 1024;   // + barriers + alignment slack
constexpr uint32_t kQKTx = 4 * kTile + 2 * kTail;
constexpr uint32_t kVTx = 2 * kTile + kTail;
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
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

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
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(base + ((j ^ sw) << 4)), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(f[i], vec[8 * j + i], acc);
    }
    return acc;
}
// row 0 of a tail tile (swizzle index 0 -> chunks in natural order) -> 64 floats
__device__ __forceinline__ void load_row0(const uint8_t* tile, float (&vec)[64]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) unpack8(*reinterpret_cast<const uint4*>(tile + (j << 4)), &vec[8 * j]);
}

__global__ void __launch_bounds__(kThreads, 2)
attn_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_tail,
                    __nv_bfloat16* __restrict__ out, int n_items) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint8_t* sQ = smem + kOffQ;
    uint8_t* sK = smem + kOffK;
    uint8_t* sV = smem + kOffV;
    uint8_t* sQT = smem + kOffQT;
    uint8_t* sKT = smem + kOffKT;
    uint8_t* sVT = smem + kOffVT;
    float* sPT = reinterpret_cast<float*>(smem + kOffPT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint64_t* qk_full = bars + 0;    // TMA  -> everyone      (once per item)
    uint64_t* v_full = bars + 1;     // TMA  -> MMA, tail     (once per item)
    uint64_t* s_full = bars + 2;     // MMA  -> softmax       (once per tile)
    uint64_t* p_full = bars + 3;     // softmax(128) -> MMA   (once per tile)
    uint64_t* o_full = bars + 4;     // MMA  -> softmax       (once per tile)
    uint64_t* o_read = bars + 5;     // softmax(128) -> MMA   (once per tile): O left TMEM, S region reusable
    uint64_t* q_read = bars + 6;     // softmax(128) -> MMA   (once per item): Q rows consumed by CUDA cores
    uint64_t* tk_done = bars + 7;    // tail warp -> MMA      (once per item): K consumed by CUDA cores
    uint64_t* tv_done = bars + 8;    // tail warp -> MMA      (once per item): V consumed by CUDA cores
    uint64_t* vt_read = bars + 9;    // softmax(128) -> MMA   (once per item): v row 256 consumed by the epilogues
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

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
            mbar_init(tk_done, 1);
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

    if (warp == 0) {
        // ===================== TMA producer + MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, 256);
            constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, /*b_mn_major=*/1);
            const uint64_t dQ = umma_desc_k_sw128(smem_u32(sQ));
            const uint64_t dK = umma_desc_k_sw128(smem_u32(sK));
            const uint64_t dV = umma_desc_k_sw128(smem_u32(sV));

            // item = frame*16 + head; column blocks of the [M,3072] QKV matrix: q -> head, k -> 16+head, v -> 32+head
            auto load_qk = [&](int item) {
                const int f = item >> 4, h = item & 15;
                mbar_arrive_expect_tx(qk_full, kQKTx);
                tma_load_4d(sQ, &tm_qkv, qk_full, 0, 0, h, f);
                tma_load_4d(sQ + kTile, &tm_qkv, qk_full, 0, 128, h, f);
                tma_load_4d(sQT, &tm_tail, qk_full, 0, 256, h, f);
                tma_load_4d(sK, &tm_qkv, qk_full, 0, 0, 16 + h, f);
                tma_load_4d(sK + kTile, &tm_qkv, qk_full, 0, 128, 16 + h, f);
                tma_load_4d(sKT, &tm_tail, qk_full, 0, 256, 16 + h, f);
            };
            auto load_v = [&](int item) {
                const int f = item >> 4, h = item & 15;
                mbar_arrive_expect_tx(v_full, kVTx);
                tma_load_4d(sV, &tm_qkv, v_full, 0, 0, 32 + h, f);
                tma_load_4d(sV + kTile, &tm_qkv, v_full, 0, 128, 32 + h, f);
                tma_load_4d(sVT, &tm_tail, v_full, 0, 256, 32 + h, f);
            };

            int it = 0;
            uint32_t n2 = 0;   // running tile counter (s_full / p_full / o_full / o_read complete once per tile)
            if (static_cast<int>(blockIdx.x) < n_items) {
                load_qk(blockIdx.x);
                load_v(blockIdx.x);
            }
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int next = item + gridDim.x;
                mbar_wait(qk_full, it & 1);
                for (int tile = 0; tile < 2; ++tile, ++n2) {
                    // the S region doubles as P / O storage of the previous tile: wait until its O has been read
                    if (n2 > 0) mbar_wait(o_read, (n2 - 1) & 1);
                    tc_fence_after();
                    // S = Q_tile K^T : 4 k-steps over the 64 head dims, N = 256 keys
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_base, dQ + static_cast<uint64_t>((tile * kTile) >> 4) + static_cast<uint64_t>(2 * k),
                                     dK + static_cast<uint64_t>(2 * k), idesc_s, k > 0);
                    umma_commit(s_full);
                    if (tile == 1) {
                        // Q / K smem is free once the last S MMA has retired and the CUDA-core readers are done
                        mbar_wait(s_full, n2 & 1);
                        mbar_wait(q_read, it & 1);
                        mbar_wait(tk_done, it & 1);
                        if (next < n_items) load_qk(next);
                    }
                    mbar_wait(p_full, n2 & 1);
                    if (tile == 0) mbar_wait(v_full, it & 1);
                    tc_fence_after();
                    // O = P V : A = P from TMEM (16 keys = 8 columns per step), B = V rows as MN-major operand
                    // (16 keys = 16 rows of 128 B = +2048 B per step)
#pragma unroll
                    for (int kk = 0; kk < 16; ++kk)
                        umma_bf16_ts(tmem_base + kOCol, tmem_base + kPCol + kk * 8,
                                     dV + static_cast<uint64_t>((kk * 2048) >> 4), idesc_o, kk > 0);
                    umma_commit(o_full);
                }
                // PV of the last tile retired and the tail warp is done with V -> V smem is free
                mbar_wait(o_full, (n2 - 1) & 1);
                mbar_wait(tv_done, it & 1);
                mbar_wait(vt_read, it & 1);
                if (next < n_items) load_v(next);
            }
        }
    } else if (warp <= 4) {
        // ===================== softmax + epilogue warps: one query row per thread =====================
        const int q = warp & 3;                       // TMEM lane quarter
        const int r = q * 32 + lane;                  // row inside the 128-row tile
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        constexpr float kLog2e = 1.4426950408889634f;
        uint32_t n2 = 0;
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int f = item >> 4, head = item & 15;
            // ---- scores against the 257th key for this thread's two rows (CUDA cores)
            mbar_wait(qk_full, it & 1);
            float s_tail[2];
            {
                float kt[64];
                load_row0(sKT, kt);
                s_tail[0] = dot_row64(sQ, r, kt);
                s_tail[1] = dot_row64(sQ + kTile, r, kt);
            }
            mbar_arrive(q_read);
            for (int tile = 0; tile < 2; ++tile, ++n2) {
                mbar_wait(s_full, n2 & 1);
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

                // ---- epilogue: (O + p_tail * v_256) / rowsum -> bf16 -> out[(frame*257 + tok), head*64 ..]
                mbar_wait(o_full, n2 & 1);
                tc_fence_after();
                uint32_t o0[32], o1[32];
                tmem_ld32(t_lane + kOCol, o0);
                tmem_ld32(t_lane + kOCol + 32, o1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(o_read);
                const int tok = tile * 128 + r;
                uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(f) * kS + tok) * 1024 + head * 64);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float vt[8];
                    unpack8(*reinterpret_cast<const uint4*>(sVT + (j << 4)), vt);
                    const uint32_t* o = j < 4 ? &o0[8 * j] : &o1[8 * (j - 4)];
                    float y[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[i] = fmaf(p_tail, vt[i], __uint_as_float(o[i])) * inv_sum;
                    uint4 w;
                    w.x = pack_bf16(y[0], y[1]);
                    w.y = pack_bf16(y[2], y[3]);
                    w.z = pack_bf16(y[4], y[5]);
                    w.w = pack_bf16(y[6], y[7]);
                    dst[j] = w;
                }
            }
            mbar_arrive(vt_read);
        }
    } else {
        // ===================== warp 5: the 257th query row on CUDA cores =====================
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int f = item >> 4, head = item & 15;
            mbar_wait(qk_full, it & 1);
            float qv[64];
            load_row0(sQT, qv);
            // scores: lane handles keys lane, lane+32, ..., lane+224; key 256 is handled by every lane (same value)
            float sc[8];
            float mx;
            {
                float kt[64];
                load_row0(sKT, kt);
                float a = 0.f;
#pragma unroll
                for (int i = 0; i < 64; ++i) a = fmaf(qv[i], kt[i], a);
                mx = a;
                sPT[256] = a;   // all lanes write the same value
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int key = lane + 32 * j;
                sc[j] = dot_row64(key < 128 ? sK : sK + kTile, key & 127, qv);
                mx = fmaxf(mx, sc[j]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float p = __expf(sc[j] - mx);
                sum += p;
                sPT[lane + 32 * j] = p;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float p256 = __expf(sPT[256] - mx);
            sum += p256;
            __syncwarp();
            if (lane == 0) mbar_arrive(tk_done);
            // output: lane owns dims 2*lane, 2*lane+1
            mbar_wait(v_full, it & 1);
            float a0 = 0.f, a1 = 0.f;
            const int chunk = lane >> 2, within = (lane & 3) * 4;
#pragma unroll 4
            for (int key = 0; key < 256; ++key) {
                const uint8_t* rowp = (key < 128 ? sV : sV + kTile) + (key & 127) * 128;
                const uint32_t w = *reinterpret_cast<const uint32_t*>(rowp + ((chunk ^ (key & 7)) << 4) + within);
                const float p = sPT[key];
                a0 = fmaf(p, __uint_as_float(w << 16), a0);
                a1 = fmaf(p, __uint_as_float(w & 0xFFFF0000u), a1);
            }
            {
                const uint32_t w = *reinterpret_cast<const uint32_t*>(sVT + (chunk << 4) + within);
                a0 = fmaf(p256, __uint_as_float(w << 16), a0);
                a1 = fmaf(p256, __uint_as_float(w & 0xFFFF0000u), a1);
            }
            const float inv = 1.0f / sum;
            *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(f) * kS + 256) * 1024 + head * 64 + 2 * lane) =
                pack_bf16(a0 * inv, a1 * inv);
            __syncwarp();
            if (lane == 0) mbar_arrive(tv_done);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem_base);
}

}  // namespace attn_v2

int launch_attention(const void* qkv, void* out, int n_frames, cudaStream_t s) {
    using namespace attn_v2;
    static const bool use_v1 = []() {
        const char* e = getenv("HVLM_ATTN_V1");
        return e && e[0] == '1';
    }();
    if (use_v1) return launch_attention_v1(qkv, out, n_frames, s);
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
        if (cudaFuncSetAttribute(attn_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess)
            return HVLM_ERR_CUDA;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int max_ctas = 2 * num_sms();
    const int grid = n_items < max_ctas ? n_items : max_ctas;
    attn_tcgen05_kernel<<<grid, kThreads, kSmem, s>>>(tq, tt, static_cast<__nv_bfloat16*>(out), n_items);
    return check_last("attention");
}

}  // namespace hvlm

extern "C" int hvlm_vit_attention(const void* qkv, void* out, int n_frames, void* stream) {
    using namespace hvlm;
    if (!qkv || !out || n_frames <= 0) return HVLM_ERR_BAD_ARG;
    if (!aligned16(qkv) || !aligned16(out)) return HVLM_ERR_ALIGN;
    StageTimer st(HVLM_STAGE_ATTENTION, static_cast<cudaStream_t>(stream));
    return launch_attention(qkv, out, n_frames, static_cast<cudaStream_t>(stream));
}
