// ViT-L/14 self-attention on tcgen05 / TMEM for sm_100a:  out = softmax(q k^T) v  per (frame, head),
// 257 tokens x 64 dims, no mask, no dropout (HF CLIPAttention in eval; q already carries the 64^-1/2 scale).
//
// Input  : qkv bf16, COLUMN-BLOCK-MAJOR [48 column blocks: q0..q15 | k0..k15 | v0..v15][M = frames*257][64] as written by
//          the QKV GEMM's epilogue -- every (frame, head) operand is one contiguous 32.9 KB block.
// Output : bf16 [frame*257 + token][head*64 + d]  (the K-major A operand of out_proj).
//
// Design (sized so that TWO CTAs are resident per SM: one CTA's MMAs overlap the other's softmax):
//   * 257 = 256 + 1.  The tensor cores handle the 256 x 256 block (queries/keys 0..255) as two 128-row tiles.
//     The scores that involve token 256 also come from the tensor cores, as four N=16 MMAs per item against 16-row
//     tiles that start at row 256 (row 0 of the tile is the token, the other rows are ignored):
//        Q_tile0/1 x Ktail^T -> score of every query against the 257th key
//        K_tile0/1 x Qtail^T -> score of the 257th query against every key (lane = key)
//     They are issued for the NEXT item while the current item's last P*V runs, into TMEM columns whose S values are
//     already dead, and every softmax thread picks its four values up right before it releases the accumulator.
//     The 257th key then costs one extra term in max / sum / output; the 257th query's softmax and P*V (257 x 64 MACs)
//     run on the CUDA cores of the four softmax warps (64 keys each) while they wait for the first P*V MMA.
//   * TMEM: 256 columns per CTA.  S[128x256] fp32 fills them; the un-normalised probabilities are written back
//     IN PLACE as packed bf16 (P aliases columns 0..127, tcgen05.st) and feed the second MMA as a TMEM A-operand;
//     O[128x64] accumulates in columns 128..191; the token-256 score blocks use columns 192..255.
//   * smem (109 KB): Q (2 tiles) | K[256x64] | V[256x64] | 16-row tail tiles of q,k | row 256 of v | tail-query
//     scratch | per-warp store staging.  V is consumed as an MN-major B operand (no transpose anywhere).
//   * warps: 0 = TMA + MMA issue (warp-uniform control flow, one elected lane issues), 1..4 = softmax + epilogue
//     (one query row per thread, fp32).  Output rows are transposed through a small per-warp smem buffer so that
//     every global store instruction writes four full 128-byte lines.
//   * next item's Q/K (V) loads are issued as soon as the current item's last S (PV) MMA has retired and the
//     CUDA-core readers of that buffer have signalled.
#include <cstdlib>

#include "hvlm_internal.cuh"
#include "hvlm_ptx.cuh"

namespace hvlm {
namespace attn {

constexpr int kThreads = 160;
constexpr int kS = HVLM_VIT_TOKENS;        // 257
constexpr int kTile = 128 * 64 * 2;        // 16384 : one [128 x 64] bf16 operand tile
constexpr int kTailTile = 16 * 128;        // 2048  : rows 256..271 of q / k (row 0 = token 256; N=16 MMA operand)
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + 2 * kTile;
constexpr int kOffV = kOffK + 2 * kTile;
constexpr int kOffQT = kOffV + 2 * kTile;  // q rows 256..271
constexpr int kOffKT = kOffQT + kTailTile; // k rows 256..271
constexpr int kOffVT = kOffKT + kTailTile; // v row 256 (128 B used; 1 KB slot keeps the swizzle phase at 0)
constexpr int kOffStage = kOffVT + 1024;   // 4 warps x 1 KB output staging
constexpr int kOffPT = kOffStage + 4 * 1024;   // float[272] scores + float[272] probabilities of the tail query
constexpr int kOffPart = kOffPT + 2 * 272 * 4; // float[4][64]: per-warp partial P*V of the tail query
constexpr int kOffBar = kOffPart + 4 * 64 * 4;
constexpr int kSmem = kOffBar + 128 + 1024;    // + barriers + alignment slack
static_assert(2 * (kSmem + 1024) <= 228 * 1024, "two CTAs must fit in one SM's shared memory");
// ping-pong variant: ONE CTA per SM runs two copies ("programs") of the algorithm, each with its own shared-memory image
// and its own 256 TMEM columns; named barriers make their exp passes take turns on the MUFU pipe
constexpr int kSmemHalf = ((kOffBar + 128 + 1023) / 1024) * 1024;
constexpr int kSmemPP = 2 * kSmemHalf + 1024;
static_assert(kSmemPP <= 227 * 1024, "both programs must fit in one CTA's shared memory");
constexpr int kBarTurnA = 5;                   // program B arrives, program A waits: "B's exp pass is over"
constexpr int kBarTurnB = 6;                   // program A arrives, program B waits
constexpr bool kPingPongDefault = false;
constexpr uint32_t kQKTx = 4 * kTile + 2 * kTailTile;
constexpr uint32_t kVTx = 2 * kTile + 128;
constexpr int kPCol = 0;                   // P (bf16 pairs) : TMEM columns [0,128)
constexpr int kOCol = 128;                 // O accumulator  : TMEM columns [128,192)
constexpr int kTCol = 192;                 // token-256 score blocks, 16 columns each (column 0 used):
                                           //   +0 / +16 : queries of tile 0 / 1 against key 256
                                           //   +32 / +48: query 256 against keys of tile 0 / 1 (lane = key)

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
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& v) {   // valid after tmem_ld_wait()
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
}
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
__device__ __forceinline__ void max_chunk32(const uint32_t (&v)[32], float& m) {
    float a = m, b = m;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        a = fmaxf(a, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
        b = fmaxf(b, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
    }
    m = fmaxf(a, b);
}
// 2^x for a PAIR of arguments on the FMA / integer pipes (no MUFU): round-to-nearest split x = n + f, |f| <= 0.5, cubic for 2^f
// (relative error 4e-4, far below the bf16 rounding of P), exponent inserted with an integer add.  x <= 0 here (row maximum
// subtracted); arguments below -126 are clamped (their probabilities flush to zero either way).
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& p0, float& p1) {
    constexpr float kMagic = 12582912.0f;     // 1.5 * 2^23: x + kMagic keeps round(x) in the low mantissa bits
    x0 = fmaxf(x0, -126.0f);
    x1 = fmaxf(x1, -126.0f);
    float t0, t1, n0, n1, f0, f1;
    fadd2(t0, t1, x0, x1, kMagic, kMagic);
    fadd2(n0, n1, t0, t1, -kMagic, -kMagic);
    fadd2(f0, f1, x0, x1, -n0, -n1);
    float q0, q1;
    ffma2(q0, q1, f0, f1, 0.0555041f, 0.0555041f, 0.2402265f, 0.2402265f);
    ffma2(q0, q1, q0, q1, f0, f1, 0.6931472f, 0.6931472f);
    ffma2(q0, q1, q0, q1, f0, f1, 1.0f, 1.0f);
    p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
    p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}
#ifndef HVLM_ATTN_POLY
#define HVLM_ATTN_POLY 0      // exp pass: every HVLM_ATTN_POLY-th pair of probabilities takes exp2_poly2 instead of MUFU.EX2 (0: none)
                              // measured stand-alone, 100 frames: 61.9 us (0) / 61.6 (4: 12.5 %) / 61.9 (2: 25 %) / 69.1 (1: 50 %): the
                              // exp pass is not what bounds the kernel (its serial chain per tile is) -- off
#endif

// The softmax warps are issue-bound as much as MUFU-bound (two of them share a scheduler), so the exp pass and the
// epilogue use the packed fp32x2 FMA / ADD / MUL of hvlm_ptx.cuh: fewer instructions per element is what pays.
// p = 2^(s*log2e - mxl) for one 32-column chunk; packed bf16 pairs; returns the chunk's sum
__device__ __forceinline__ float exp_chunk32(const uint32_t (&v)[32], float log2e, float mxl, uint32_t (&pk)[16]) {
    float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
    const float nm = -mxl;
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        float x0, x1, x2, x3;
        ffma2(x0, x1, __uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), log2e, log2e, nm, nm);
        ffma2(x2, x3, __uint_as_float(v[2 * j + 2]), __uint_as_float(v[2 * j + 3]), log2e, log2e, nm, nm);
        float p0 = fast_exp2(x0), p1 = fast_exp2(x1), p2, p3;
        if (HVLM_ATTN_POLY != 0 && ((j / 2) % HVLM_ATTN_POLY) == HVLM_ATTN_POLY - 1) {
            exp2_poly2(x2, x3, p2, p3);
        } else {
            p2 = fast_exp2(x2);
            p3 = fast_exp2(x3);
        }
        fadd2(s0, s1, s0, s1, p0, p1);
        fadd2(t0, t1, t0, t1, p2, p3);
        pk[j] = pack_bf16(p0, p1);
        pk[j + 1] = pack_bf16(p2, p3);
    }
    return (s0 + s1) + (t0 + t1);
}

// debug: per-phase timestamps (trace == nullptr in production)
#define ATTN_TRACE(role, slot)                                                                          \
    do {                                                                                                \
        if (trace != nullptr && it < 4)                                                                 \
            trace[((static_cast<size_t>(vcta) * 5 + (role)) * 4 + it) * 16 + (slot)] = clock64();   \
    } while (0)

__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// PP = false: 160 threads, two CTAs per SM (the co-resident CTA's MMAs overlap this one's softmax).
// PP = true : 320 threads, one CTA per SM = two programs A / B (threads 0..159 / 160..319) that walk interleaved item
//             lists; their exp passes strictly alternate (A, B, A, B ...), so a program's MUFU-bound pass never shares the
//             pipe with the other's and always overlaps the other's TMEM / epilogue / MMA-wait phases.
template <bool PP>
__global__ void __launch_bounds__(PP ? 2 * kThreads : kThreads, PP ? 1 : 2)
attn_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_tail16,
                    const __grid_constant__ CUtensorMap tm_tail1, __nv_bfloat16* __restrict__ out, int n_items,
                    long long* __restrict__ trace, int take_turns) {
    extern __shared__ uint8_t smem_raw[];
    const int half = PP ? static_cast<int>(threadIdx.x) / kThreads : 0;      // program of this thread
    const int tid = static_cast<int>(threadIdx.x) - half * kThreads;
    // program B of CTA b walks the item list of "virtual CTA" b + gridDim: per SM the two lists then differ by at most one
    // item in total, exactly like CTAs b and b + 148 of the two-CTA kernel (2b / 2b + 1 would pair the long lists)
    const int vcta = static_cast<int>(blockIdx.x) + (PP ? half * static_cast<int>(gridDim.x) : 0);
    const int vgrid = PP ? static_cast<int>(gridDim.x) * 2 : static_cast<int>(gridDim.x);
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem0 = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint8_t* smem = smem0 + half * kSmemHalf;
    uint8_t* sQ = smem + kOffQ;
    uint8_t* sK = smem + kOffK;
    uint8_t* sV = smem + kOffV;
    uint8_t* sQT = smem + kOffQT;
    uint8_t* sKT = smem + kOffKT;
    uint8_t* sVT = smem + kOffVT;
    uint8_t* sStage = smem + kOffStage;
    float* sSc = reinterpret_cast<float*>(smem + kOffPT);          // scores of the tail query: [0..255] keys, [256] key 256
    float* sPp = sSc + 272;                                        // its un-normalised probabilities (keys 0..255)
    float* sPart = reinterpret_cast<float*>(smem + kOffPart);      // [4 warps][64 dims]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint64_t* qk_full = bars + 0;    // TMA  -> everyone        (once per item)
    uint64_t* v_full = bars + 1;     // TMA  -> MMA, softmax    (once per item)
    uint64_t* s_full = bars + 2;     // MMA  -> softmax         (once per tile)
    uint64_t* p_full = bars + 3;     // softmax(128) -> MMA     (once per tile)
    uint64_t* o_full = bars + 4;     // MMA  -> softmax         (once per tile)
    uint64_t* o_read = bars + 5;     // softmax(128) -> MMA     (once per tile): O (and, on the item's last tile, the
                                     //   next item's token-256 scores) left TMEM, S region reusable
    uint64_t* t_full = bars + 6;     // MMA  -> softmax         (once per item): token-256 score blocks written
    uint64_t* t_read = bars + 7;     // softmax(128) -> MMA     (once, first item only): those blocks were read
    uint64_t* vt_read = bars + 8;    // softmax(128) -> MMA     (once per item): V / v row 256 consumed by the CUDA cores
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem0 + kOffBar + 9 * 8);   // program A's slot serves both
    constexpr int kCols = PP ? 512 : 256;

    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int bar_sc = 1 + 2 * half, bar_part = 2 + 2 * half;     // per-program named barriers

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_qkv);
            tma_prefetch_desc(&tm_tail16);
            tma_prefetch_desc(&tm_tail1);
            mbar_init(qk_full, 1);
            mbar_init(v_full, 1);
            mbar_init(s_full, 1);
            mbar_init(p_full, 128);
            mbar_init(o_full, 1);
            mbar_init(o_read, 128);
            mbar_init(t_full, 1);
            mbar_init(t_read, 128);
            mbar_init(vt_read, 128);
            fence_mbar_init();
        }
        __syncwarp();
        if (half == 0) tmem_alloc<kCols>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot + static_cast<uint32_t>(half * 256);
    pdl_launch_dependents();
    pdl_wait();              // everything above overlapped the previous kernel's tail

    if (warp == 0) {
        // ===================== TMA producer + MMA issuer =====================
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 256);
        constexpr uint32_t idesc_t = umma_idesc_bf16(128, 16);
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, /*b_mn_major=*/1);
        const uint64_t dQ = umma_desc_k_sw128(smem_u32(sQ));
        const uint64_t dK = umma_desc_k_sw128(smem_u32(sK));
        const uint64_t dV = umma_desc_k_sw128(smem_u32(sV));
        const uint64_t dQT = umma_desc_k_sw128(smem_u32(sQT));
        const uint64_t dKT = umma_desc_k_sw128(smem_u32(sKT));

        // item = frame*16 + head; column blocks: q -> head, k -> 16+head, v -> 32+head
        auto load_qk = [&](int item) {
            const int h = item & 15, r0 = (item >> 4) * kS;   // first row of the frame in the [M] dimension
            if (elect_one()) {
                mbar_arrive_expect_tx(qk_full, kQKTx);
                tma_load_3d(sQ, &tm_qkv, qk_full, 0, r0, h);
                tma_load_3d(sQ + kTile, &tm_qkv, qk_full, 0, r0 + 128, h);
                tma_load_3d(sQT, &tm_tail16, qk_full, 0, r0 + 256, h);
                tma_load_3d(sK, &tm_qkv, qk_full, 0, r0, 16 + h);
                tma_load_3d(sK + kTile, &tm_qkv, qk_full, 0, r0 + 128, 16 + h);
                tma_load_3d(sKT, &tm_tail16, qk_full, 0, r0 + 256, 16 + h);
            }
            __syncwarp();
        };
        auto load_v = [&](int item) {
            const int h = item & 15, r0 = (item >> 4) * kS;
            if (elect_one()) {
                mbar_arrive_expect_tx(v_full, kVTx);
                tma_load_3d(sV, &tm_qkv, v_full, 0, r0, 32 + h);
                tma_load_3d(sV + kTile, &tm_qkv, v_full, 0, r0 + 128, 32 + h);
                tma_load_3d(sVT, &tm_tail1, v_full, 0, r0 + 256, 32 + h);
            }
            __syncwarp();
        };
        // the four token-256 score blocks of the item whose Q / K are in shared memory
        auto issue_tail = [&]() {
            if (elect_one()) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint64_t da = (b < 2 ? dQ : dK) + static_cast<uint64_t>(((b & 1) * kTile) >> 4);
                    const uint64_t db = b < 2 ? dKT : dQT;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_base + kTCol + b * 16, da + static_cast<uint64_t>(2 * k),
                                     db + static_cast<uint64_t>(2 * k), idesc_t, k > 0);
                }
                umma_commit(t_full);
            }
            __syncwarp();
        };

        auto issue_s = [&](int tile) {
            // S = Q_tile K^T : 4 k-steps over the 64 head dims, N = 256 keys
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tmem_base, dQ + static_cast<uint64_t>((tile * kTile) >> 4) + static_cast<uint64_t>(2 * k),
                                 dK + static_cast<uint64_t>(2 * k), idesc_s, k > 0);
                umma_commit(s_full);
            }
            __syncwarp();
        };

        int it = 0;
        uint32_t n2 = 0;   // running tile counter (s_full / p_full / o_full / o_read complete once per tile)
        // items are walked last-to-first: the QKV GEMM wrote the last frames last, so they are still in L2
        if (vcta < n_items) {
            load_qk(n_items - 1 - vcta);
            load_v(n_items - 1 - vcta);
            mbar_wait(qk_full, 0);
            tc_fence_after();
            issue_tail();
            mbar_wait(t_read, 0);    // the softmax threads hold the first item's token-256 scores in registers
            tc_fence_after();
            issue_s(0);
        }
        for (int idx = vcta; idx < n_items; idx += vgrid, ++it) {
            const int next_idx = idx + vgrid;
            const int next = n_items - 1 - next_idx;
            if (lane == 0) ATTN_TRACE(0, 0);
            for (int tile = 0; tile < 2; ++tile, ++n2) {
                if (tile == 1) {
                    // the S region doubles as P / O storage of the previous tile: wait until its O has been read
                    mbar_wait(o_read, (n2 - 1) & 1);
                    if (lane == 0) ATTN_TRACE(0, 2 + tile * 6);
                    tc_fence_after();
                    issue_s(1);
                    // Q / K smem is free once the last S MMA has retired (the token-256 blocks retired long ago, and
                    // the CUDA-core read of the two tail rows happened before this item's first p_full arrival)
                    mbar_wait(s_full, n2 & 1);
                    if (next_idx < n_items) load_qk(next);
                }
                // (tile 0: its S MMAs were issued at the end of the previous item / in the prologue)
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
            if (next_idx < n_items) {
                // S columns 192..255 are dead (every softmax thread passed p_full): the next item's token-256
                // blocks go there, behind the P*V in the tensor pipe; the softmax threads read them before they signal
                // o_read, and then the next item's first S tile can start while this item's epilogue is still running
                mbar_wait(qk_full, (it + 1) & 1);
                tc_fence_after();
                issue_tail();
                mbar_wait(o_read, (n2 - 1) & 1);
                tc_fence_after();
                issue_s(0);
            }
            // PV of the last tile retired and the CUDA-core readers are done with V -> V smem is free
            mbar_wait(o_full, (n2 - 1) & 1);
            mbar_wait(vt_read, it & 1);
            if (lane == 0) ATTN_TRACE(0, 14);
            if (next_idx < n_items) load_v(next);
        }
    } else {
        // ===================== softmax + epilogue warps: one query row per thread =====================
        const int q = (static_cast<int>(threadIdx.x) >> 5) & 3;   // TMEM lane quarter: fixed by the HARDWARE warp index
        const int r = q * 32 + lane;                  // row inside the 128-row tile
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint8_t* stage = sStage + (warp - 1) * 1024;  // this warp's 8-row x 128-byte transpose buffer
        constexpr float kLog2e = 1.4426950408889634f;
        uint32_t n2 = 0;
        int it = 0;
        uint32_t tl[4] = {0u, 0u, 0u, 0u};             // token-256 scores of the upcoming item (see kTCol)
        if (vcta < n_items) {
            mbar_wait(t_full, 0);
            tc_fence_after();
#pragma unroll
            for (int b = 0; b < 4; ++b) tmem_ld1(t_lane + kTCol + b * 16, tl[b]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(t_read);
        }
        for (int idx = vcta; idx < n_items; idx += vgrid, ++it) {
            const int item = n_items - 1 - idx;
            const int f = item >> 4, head = item & 15;
            const bool has_next = idx + vgrid < n_items;
            const bool tr = (lane == 0);
            if (tr) ATTN_TRACE(warp, 0);
            const float s_tail[2] = {__uint_as_float(tl[0]), __uint_as_float(tl[1])};
            sSc[r] = __uint_as_float(tl[2]);
            sSc[r + 128] = __uint_as_float(tl[3]);
            mbar_wait(qk_full, it & 1);
            if (tr) ATTN_TRACE(warp, 1);
            if (q == 1) {   // q_256 . k_256: two dims per lane (row 0 of a 1024-aligned tile: natural chunk order)
                const uint32_t qa = *reinterpret_cast<const uint32_t*>(sQT + lane * 4);
                const uint32_t ka = *reinterpret_cast<const uint32_t*>(sKT + lane * 4);
                float d = __uint_as_float(qa << 16) * __uint_as_float(ka << 16) +
                          __uint_as_float(qa & 0xFFFF0000u) * __uint_as_float(ka & 0xFFFF0000u);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                if (lane == 0) sSc[256] = d;
            }
            float tq_inv = 0.f, tq_p256 = 0.f;          // tail query: 1 / rowsum and its weight on key 256
            for (int tile = 0; tile < 2; ++tile, ++n2) {
                if (tr) ATTN_TRACE(warp, 2 + tile * 7);
                mbar_wait(s_full, n2 & 1);
                if (tr) ATTN_TRACE(warp, 3 + tile * 7);
                tc_fence_after();
                // ---- pass 1: row max over 256 keys (TMEM) and the tail key.  Two register buffers: the next chunk's
                //      tcgen05.ld is in flight while the current one is reduced.
                //      Three tcgen05.ld (96 columns) are in flight per round trip, independent running maxima.
                float mx;
                {
                    float m0 = s_tail[tile], m1 = m0, m2 = m0, m3 = m0;
#pragma unroll 1
                    for (int c0 = 0; c0 < 256; c0 += 96) {          // columns [0,96) [96,192) [192,256)
                        uint32_t v0[32], v1[32], v2[32];
                        tmem_ld32(t_lane + c0, v0);
                        tmem_ld32(t_lane + c0 + 32, v1);
                        if (c0 < 192) tmem_ld32(t_lane + c0 + 64, v2);
                        tmem_ld_wait();
                        max_chunk32(v0, m0);
                        max_chunk32(v1, m1);
                        if (c0 < 192) max_chunk32(v2, m2);
                    }
                    mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                }
                if (tr) ATTN_TRACE(warp, 4 + tile * 7);
                // ---- pass 2: p = exp(s - max) (fp32), row sum, P -> packed bf16 back into TMEM (aliases S; chunk c's
                //      16 P columns land in S columns that were already consumed)
                const float mxl = mx * kLog2e;
                float sum = 0.f, sum1 = 0.f;
                if (PP && take_turns) {
                    // wait for the other program's exp pass to end (A goes first: nothing to wait for on its first tile)
                    if (half == 0) {
                        if (n2 > 0) named_bar_sync(kBarTurnA, 256);
                    } else {
                        named_bar_sync(kBarTurnB, 256);
                    }
                }
                {
                    uint32_t va[32], vb[32], pk[16];
                    tmem_ld32(t_lane, va);
#pragma unroll 1
                    for (int c = 0; c < 8; c += 2) {
                        tmem_ld_wait();
                        tmem_ld32(t_lane + (c + 1) * 32, vb);
                        sum += exp_chunk32(va, kLog2e, mxl, pk);
                        tmem_st16(t_lane + kPCol + c * 16, pk);
                        tmem_ld_wait();
                        if (c + 2 < 8) tmem_ld32(t_lane + (c + 2) * 32, va);
                        sum1 += exp_chunk32(vb, kLog2e, mxl, pk);
                        tmem_st16(t_lane + kPCol + (c + 1) * 16, pk);
                    }
                }
                if (PP && take_turns) named_bar_arrive(half == 0 ? kBarTurnB : kBarTurnA, 256);   // the MUFU pipe is the other's
                const float p_tail = fast_exp2(fmaf(s_tail[tile], kLog2e, -mxl));
                sum += sum1 + p_tail;
                const float inv_sum = 1.0f / sum;
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(p_full);
                if (tr) ATTN_TRACE(warp, 5 + tile * 7);

                // ---- the 257th query row on CUDA cores, in the shadow of the P*V MMAs
                if (tile == 0) {
                    // softmax of its 257 scores (every warp redundantly) and P*V over this warp's 64 keys
                    named_bar_sync(bar_sc, 128);       // all scores are in sSc
                    mbar_wait(v_full, it & 1);         // V / v row 256 are read by the CUDA cores from here on
                    const float s256 = sSc[256];
                    float sc[8];
                    float tmx = s256;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        sc[j] = sSc[lane + 32 * j];
                        tmx = fmaxf(tmx, sc[j]);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) tmx = fmaxf(tmx, __shfl_xor_sync(0xffffffffu, tmx, o));
                    tq_p256 = __expf(s256 - tmx);
                    float tsum = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float p = __expf(sc[j] - tmx);
                        tsum += p;
                        if ((j >> 1) == q) sPp[lane + 32 * j] = p;   // keys 64q .. 64q+63 belong to this warp
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
                    tq_inv = 1.0f / (tsum + tq_p256);
                } else {
                    // P*V over this warp's 64 keys (its own probabilities: no cross-warp dependency), then combine
                    __syncwarp();
                    const int g = lane >> 3;           // 16 keys per lane group
                    const int c = lane & 7;            // 16-byte chunk of the head dim: dims 8c .. 8c+7
                    float acc[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 4
                    for (int kk = 0; kk < 16; ++kk) {
                        const int key = q * 64 + g * 16 + kk;
                        const uint8_t* rowp = (key < 128 ? sV : sV + kTile) + (key & 127) * 128;
                        float vv[8];
                        unpack8(*reinterpret_cast<const uint4*>(rowp + ((c ^ (key & 7)) << 4)), vv);
                        const float p = sPp[key];
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i] = fmaf(p, vv[i], acc[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
                        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
                    }
                    if (g == 0) {
                        float4* dst = reinterpret_cast<float4*>(sPart + q * 64 + c * 8);
                        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                        dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
                    }
                    named_bar_sync(bar_part, 128);     // the four partial sums are in sPart
                    if (q == 1) {
                        const int d = 2 * lane;
                        const uint32_t va = *reinterpret_cast<const uint32_t*>(sVT + lane * 4);
                        const float o0 = (sPart[d] + sPart[64 + d]) + (sPart[128 + d] + sPart[192 + d]) +
                                         tq_p256 * __uint_as_float(va << 16);
                        const float o1 = (sPart[d + 1] + sPart[64 + d + 1]) + (sPart[128 + d + 1] + sPart[192 + d + 1]) +
                                         tq_p256 * __uint_as_float(va & 0xFFFF0000u);
                        *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(f) * kS + 256) * 1024 + head * 64 + d) =
                            pack_bf16(o0 * tq_inv, o1 * tq_inv);
                    }
                }

                // ---- epilogue: (O + p_tail * v_256) / rowsum -> bf16
                mbar_wait(o_full, n2 & 1);
                if (tr) ATTN_TRACE(warp, 6 + tile * 7);
                tc_fence_after();
                uint32_t o0[32], o1[32];
                tmem_ld32(t_lane + kOCol, o0);
                tmem_ld32(t_lane + kOCol + 32, o1);
                if (tile == 1 && has_next) {
                    // the next item's token-256 scores were written right behind this P*V
                    mbar_wait(t_full, (it + 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int b = 0; b < 4; ++b) tmem_ld1(t_lane + kTCol + b * 16, tl[b]);
                }
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(o_read);
                if (tr) ATTN_TRACE(warp, 7 + tile * 7);
                uint4 rowv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float vt[8];
                    unpack8(*reinterpret_cast<const uint4*>(sVT + (j << 4)), vt);
                    const uint32_t* o = j < 4 ? &o0[8 * j] : &o1[8 * (j - 4)];
                    float y[8];
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        ffma2(y[i], y[i + 1], p_tail, p_tail, vt[i], vt[i + 1], __uint_as_float(o[i]), __uint_as_float(o[i + 1]));
                        fmul2(y[i], y[i + 1], y[i], y[i + 1], inv_sum, inv_sum);
                    }
                    rowv[j].x = pack_bf16(y[0], y[1]);
                    rowv[j].y = pack_bf16(y[2], y[3]);
                    rowv[j].z = pack_bf16(y[4], y[5]);
                    rowv[j].w = pack_bf16(y[6], y[7]);
                }
                // transpose through the per-warp staging buffer (8 rows at a time) so that each global store
                // instruction writes 4 complete 128-byte rows: lane L -> row L/8 + 4i, 16-byte chunk L%8
                const size_t row0 = static_cast<size_t>(f) * kS + tile * 128 + q * 32;
#pragma unroll
                for (int part = 0; part < 4; ++part) {
                    if ((lane >> 3) == part) {
                        uint8_t* srow = stage + (lane & 7) * 128;
#pragma unroll
                        for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(srow + ((j ^ (lane & 7)) << 4)) = rowv[j];
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int rr = (lane >> 3) + 4 * i;
                        const uint4 w = *reinterpret_cast<const uint4*>(stage + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
                        *reinterpret_cast<uint4*>(out + (row0 + part * 8 + rr) * 1024 + head * 64 + (lane & 7) * 8) = w;
                    }
                    __syncwarp();
                }
                if (tr) ATTN_TRACE(warp, 8 + tile * 7);
            }
            mbar_arrive(vt_read);
        }
        if (PP && take_turns) {
            // program B may own one item less than A: it still takes its turns, so that A's waits are always answered
            if (half == 1) {
                const int va = vcta - static_cast<int>(gridDim.x);     // program A's list
                const int items_a = (va < n_items) ? (n_items - va + vgrid - 1) / vgrid : 0;
                for (int extra = it; extra < items_a; ++extra) {
#pragma unroll 1
                    for (int tile = 0; tile < 2; ++tile) {
                        named_bar_sync(kBarTurnB, 256);
                        named_bar_arrive(kBarTurnA, 256);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0 && half == 0) tmem_dealloc<kCols>(*tmem_slot);
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
    CUtensorMap tq, tt16, tt1;
    int rc = make_qkv_hm_tmap(&tq, qkv_hm, n_frames * kS, 128);
    if (rc) return rc;
    rc = make_qkv_hm_tmap(&tt16, qkv_hm, n_frames * kS, 16);    // rows past the buffer are zero-filled by TMA
    if (rc) return rc;
    rc = make_qkv_hm_tmap(&tt1, qkv_hm, n_frames * kS, 1);
    if (rc) return rc;
    // HVLM_ATTN_PINGPONG=0 / 1 selects the two-CTAs-per-SM kernel / the one-CTA ping-pong kernel (default: see DESIGN.md)
    static const bool pingpong = []() {
        const char* e = getenv("HVLM_ATTN_PINGPONG");
        return e ? e[0] == '1' : kPingPongDefault;
    }();
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(attn_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess ||
            cudaFuncSetAttribute(attn_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemPP) != cudaSuccess)
            return HVLM_ERR_CUDA;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    static const int take_turns = []() {
        const char* e = getenv("HVLM_ATTN_PP_TURNS");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    cudaError_t err;
    if (pingpong) {
        const int want = (n_items + 1) / 2;
        const int grid = want < num_sms() ? want : num_sms();
        err = launch_pdl_cls(2, attn_tcgen05_kernel<true>, dim3(grid), dim3(2 * kThreads), kSmemPP, s, tq, tt16, tt1,
                             static_cast<__nv_bfloat16*>(out), n_items, trace, take_turns);
    } else {
        const int max_ctas = 2 * num_sms();
        const int grid = n_items < max_ctas ? n_items : max_ctas;
        err = launch_pdl_cls(2, attn_tcgen05_kernel<false>, dim3(grid), dim3(kThreads), kSmem, s, tq, tt16, tt1,
                             static_cast<__nv_bfloat16*>(out), n_items, trace, 0);
    }
    if (err != cudaSuccess) {
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

// debug only (not part of the ABI header): trace[grid][5 warps][4 items][16 slots] clock64 timestamps
extern "C" __attribute__((visibility("default"))) int hvlm_debug_attention_trace(const void* qkv_hm, void* out,
                                                                                 int n_frames, long long* trace,
                                                                                 void* stream, int) {
    return hvlm::launch_attention_impl(qkv_hm, out, n_frames, static_cast<cudaStream_t>(stream), trace);
}
