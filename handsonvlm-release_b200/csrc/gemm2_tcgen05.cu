// 2-CTA (cta_group::2) persistent tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T (+ fused epilogue)
//
// A CTA pair (cluster of 2 = the two SMs of a TPC) computes one 256 x 256 output tile per step:
//   * each CTA TMA-loads ITS 128 rows of A and ITS 128 rows of B (half of the N extent) per 64-wide K block
//     (32 KB per stage instead of 48 KB: the B operand is shared across the pair by the tensor-core datapath, which
//     is what lifts the shared-memory-bandwidth ceiling of the 1-CTA 128x256 kernel)
//   * the leader CTA's elected lane issues tcgen05.mma.cta_group::2 (M=256, N=256, K=16): rows 0..127 accumulate
//     in the leader's TMEM, rows 128..255 in the peer's; tcgen05.commit multicasts the "slot free" /
//     "accumulator ready" mbarrier arrivals to both CTAs
//   * both CTAs' TMA loads complete_tx on the LEADER's full barrier (cp.async.bulk.tensor...cta_group::2)
//   * every CTA runs the same 4-warp epilogue as the 1-CTA kernel on its own 128 x 256 half (TMEM -> registers ->
//     bias / quick-GELU -> swizzled smem -> TMA store / TMA reduce-add), double-buffered against the next tile's MMAs;
//     the peer's epilogue threads release the accumulator with remote mbarrier arrivals on the leader
//   * LayerNorm folded into the GEMMs around it (template FOLD; the tower's default, see vit_forward.cu): the QKV / fc1
//     epilogues apply the row's (mean, rstd) to products of the UN-normalised bf16 rows with gamma-scaled weights, and the
//     residual GEMMs in front of them (residual-LOAD epilogue, template RL) emit those bf16 rows and the rows' statistics
#include <cstdlib>

#include "gemm_epilogue.cuh"
#include "hvlm_internal.cuh"
#include "hvlm_ptx.cuh"

// timing probes (tools/fold_probe.py builds variants of the library with these; the defaults are the product)
#ifndef HVLM_P_ARRIVE
#define HVLM_P_ARRIVE 1     // 1: cta-scope remote arrive on the leader's accumulator barrier; 0: .release.cluster (MEMBAR.GPU)
#endif
#ifndef HVLM_P_EARLY
#define HVLM_P_EARLY 0      // 1: release the accumulator right after its last tcgen05.ld, before that chunk's math
                            //    (measured: no gain for fc1, QKV with the folded LayerNorm 109-122 vs 101 us -- off)
#endif
#ifndef HVLM_P_NOXB
#define HVLM_P_NOXB 0       // 1: (WRONG RESULTS, timing only) residual-load producer without the bf16 copy
#endif
#ifndef HVLM_P_RL_LONGK
#define HVLM_P_RL_LONGK 2   // residual-load buffers for K >= 2048 (2: six K stages, one unit of lookahead; 4: five stages, two)
#endif

namespace hvlm {
namespace gemm2 {

constexpr int BM = 128;          // rows per CTA (256 per pair)
constexpr int BN = 256;          // columns per pair (each CTA loads 128 of them)
constexpr int BK = 64;
constexpr int kThreads = 192;    // warp0 TMA, warp1 MMA (+TMEM alloc), warps 2..5 epilogue (+ warps 6..9: 2nd epilogue group)
// G = number of 4-warp epilogue groups (each owns half of the tile's columns when G == 2)
// RL = "residual load" epilogue: the fp32 residual tile is TMA-LOADED into the staging buffer, the accumulator is added
//      in shared memory and the sum leaves with a plain TMA store
// RL = 0: off; 4: four staging buffers, the residual load runs two units ahead (short K: the epilogue is the critical path);
//      2: two buffers, one unit ahead, and the K pipeline keeps its six stages (long K: the epilogue has slack, the K loop
//         does not -- 4 stages instead of 5 cost the K = 4096 GEMM 20 us per call, 5 instead of 6 cost 6 us)
template <int G, int RL = 0>
struct Cfg2 {
    static constexpr int kStages = (G == 2 || RL == 4) ? 5 : 6;
    static constexpr int kBufs = RL ? RL : 2;                    // staging buffers per epilogue group
    static constexpr int kSmemBytes = kStages * (BM * BK * 2 + (BN / 2) * BK * 2) + G * kBufs * (BM * 128) + 1024 + 256 + 1024;   // + barriers + folded-epilogue vectors
};
constexpr int kABytes = BM * BK * 2;            // 16 KB
constexpr int kBBytes = (BN / 2) * BK * 2;      // 16 KB (this CTA's half of B)
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kStoreBuf = BM * 128;
constexpr int kTmemCols = 512;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are accounted on the LEADER CTA's mbarrier (peer bit of the address cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once all previously issued MMAs retired) on the mbarrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    const uint16_t mask = 0x3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst) {   // whole warp, in both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kTmemCols) : "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
#if HVLM_P_ARRIVE
    // default semantics (.release at CTA scope): what this arrival orders is the thread's own tcgen05.ld (fenced by
    // tcgen05.fence::before_thread_sync), not generic-proxy memory -- a cluster-scope release compiles to MEMBAR.ALL.GPU,
    // which waits for every outstanding global access of the thread (ncu: 18 % of the peer CTA's epilogue time)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
#else
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
#endif
}

__device__ __forceinline__ int ld_acquire_gpu(const int32_t* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// LayerNorm(1024) of `nrows` consecutive fp32 rows (<= R) by one warp: rows are read with L1-bypassing loads (they
// were just produced by TMA reduce-adds from other SMs), two-pass fp32 statistics, bf16 output.
template <int R>
__device__ __forceinline__ void ln_rows(const float* __restrict__ x, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, __nv_bfloat16* __restrict__ out, int nrows,
                                        int lane) {
    float4 v[R][8];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (r < nrows) {
            const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * 1024);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[r][j] = __ldcg(xr + lane + 32 * j);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (r >= nrows) break;
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += (v[r][j].x + v[r][j].y) + (v[r][j].z + v[r][j].w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum * (1.0f / 1024.0f);
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float a = v[r][j].x - mean, b = v[r][j].y - mean, c = v[r][j].z - mean, d = v[r][j].w - mean;
            sq += (a * a + b * b) + (c * c + d * d);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = rsqrtf(sq * (1.0f / 1024.0f) + 1e-5f);
        const float4* g4 = reinterpret_cast<const float4*>(gamma);
        const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 g = __ldg(g4 + lane + 32 * j);
            const float4 b = __ldg(b4 + lane + 32 * j);
            __nv_bfloat162 lo = __floats2bfloat162_rn((v[r][j].x - mean) * rstd * g.x + b.x, (v[r][j].y - mean) * rstd * g.y + b.y);
            __nv_bfloat162 hi = __floats2bfloat162_rn((v[r][j].z - mean) * rstd * g.z + b.z, (v[r][j].w - mean) * rstd * g.w + b.w);
            uint2 w;
            w.x = *reinterpret_cast<uint32_t*>(&lo);
            w.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(out + static_cast<size_t>(r) * 1024 + (lane + 32 * j) * 4) = w;
        }
    }
}

// Tile schedule of the persistent pairs.  The first `full` tiles (a whole number of waves) are 256 x 256; when the
// leftover tiles would occupy at most half of the pairs, each of them is cut into two 256 x 128 half tiles so that the
// last wave takes half the time (404 tiles on 74 pairs: 5.5 waves instead of 6).  A half tile runs the same K loop with
// an N=128 MMA, so every output element sees the same accumulation order either way: results do not depend on the split.
struct Sched {
    int num_tiles, num_n, full, nv, split, reverse;
    __device__ __forceinline__ Sched(int tiles, int n, int n_pairs, int allow_split, int rev)
        : num_tiles(tiles), num_n(n), reverse(rev) {
        full = (tiles / n_pairs) * n_pairs;
        const int rem = tiles - full;
        split = allow_split && rem > 0 && 2 * rem <= n_pairs;
        nv = split ? full + 2 * rem : tiles;
    }
    // half = -1: whole tile; 0 / 1: left / right 128 columns
    __device__ __forceinline__ void decode(int v, int& m_blk, int& n_blk, int& half) const {
        int tile = v;
        half = -1;
        if (split && v >= full) {
            tile = full + ((v - full) >> 1);
            half = (v - full) & 1;
        }
        if (reverse) tile = num_tiles - 1 - tile;
        m_blk = tile / num_n;
        n_blk = tile - m_blk * num_n;
    }
};

// FOLD = LayerNorm folded into the GEMMs around it (EpiArgs::ln_stats / xb_out):
//   * consumer (RL == 0; QKV / fc1): A holds the un-normalised bf16 rows, B the gamma-scaled weights; the epilogue applies
//     the row's (mean, rstd), computed from the eight per-128-column partial sums (fetched one tile ahead), with the folded
//     bias / column-sum vectors of the unit staged in shared memory; its first column tiles keep the rows' running mean
//   * producer (RL != 0; out_proj / fc2): the residual-load epilogue holds the updated fp32 row values in registers, so it
//     accumulates their partial sums (centred on the running mean), and four copy warps write the bf16 twin of every staged
//     unit (the next GEMM's A operand) -- the LayerNorm kernel between the two GEMMs (4 KB read + 2 KB write per row)
//     disappears
template <int EPI, bool LN, int G, int RL, bool FOLD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads + ((LN || G == 2 || (RL != 0 && FOLD)) ? 128 : 0), 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                     const __grid_constant__ CUtensorMap tma_bh, const __grid_constant__ CUtensorMap tma_c, int M, int N,
                     int K, EpiArgs ep, int split_tail) {
    static_assert(epi_is_staged<EPI>() || EPI == EPI_PATCH, "the 2-CTA kernel only implements the smem-staged TMA-store epilogues");
    static_assert(!(LN && G == 2), "the fused-LayerNorm warps and the second epilogue group use the same warp slots");
    static_assert(!RL || (EPI == EPI_RESID_F32 && !LN && G == 1), "residual-load epilogue: fp32 residual GEMM, one group");
    static_assert(!FOLD || (!LN && (RL || EPI == EPI_QKV_HM || EPI == EPI_BIAS_BF16 || EPI == EPI_GELU_BF16)),
                  "folded LayerNorm: residual-load producer or a bf16-output consumer");
    constexpr bool kXb = RL && FOLD;
    constexpr int kLook = RL / 2;                    // residual-load lookahead in units
    constexpr int kStages = Cfg2<G, RL>::kStages;
    constexpr int kBufs = Cfg2<G, RL>::kBufs;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * kABytes;
    uint8_t* smem_c = smem + kStages * kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + G * kBufs * kStoreBuf);
    uint64_t* full_bar = bars;                       // [kStages]  TMA (both CTAs) -> MMA        (leader's copy is used)
    uint64_t* empty_bar = bars + kStages;            // [kStages]  MMA -> TMA                    (multicast to both)
    uint64_t* tfull_bar = bars + 2 * kStages;        // [2]        MMA -> epilogue               (multicast to both)
    uint64_t* tempty_bar = bars + 2 * kStages + 2;   // [2]        epilogues of BOTH CTAs -> MMA (leader's copy is used)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    uint64_t* ld_bar = bars + 2 * kStages + 5;       // [4]        RL only: residual tile landed in staging buffer i
    uint64_t* xfull_bar = ld_bar + 4;                // [4]        RL + FOLD: epilogue (128) -> copy warps: buffer i holds the sums
    uint64_t* xdone_bar = xfull_bar + 4;             // [4]        RL + FOLD: copy warps (128) -> store warp: buffer i was read
    float4* smem_bc = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [G][32] FOLD consumer: folded bias |
                                                                                           //   column sums of the current unit

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();         // 0 = leader
    const int pair = blockIdx.x >> 1;
    const int n_pairs = gridDim.x >> 1;
    const int num_m = (M + 2 * BM - 1) / (2 * BM);
    const int num_n = N / BN;
    const int num_tiles = num_m * num_n;
    const int num_kb = (K + BK - 1) / BK;
    const Sched sched(num_tiles, num_n, n_pairs, LN ? 0 : split_tail, ep.reverse);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        tma_prefetch_desc(&tma_bh);
        tma_prefetch_desc(&tma_c);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 256 * G);   // 128*G epilogue threads in each CTA of the pair
        }
        if constexpr (RL != 0)
            for (int i = 0; i < 4; ++i) {
                mbar_init(&ld_bar[i], 1);
                mbar_init(&xfull_bar[i], 128);
                mbar_init(&xdone_bar[i], 128);
            }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();     // peer barriers are initialised and both TMEM allocations are done
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();              // everything above overlapped the previous kernel's tail

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int v = pair; v < sched.nv; v += n_pairs) {
            int m_blk, n_blk, half;
            sched.decode(v, m_blk, n_blk, half);
            const int row_a = m_blk * 2 * BM + static_cast<int>(rank) * BM;
            // this CTA supplies its half of the tile's B rows: 128 of 256, or 64 of 128 for a half tile
            const int row_b = half < 0 ? n_blk * BN + static_cast<int>(rank) * (BN / 2)
                                       : n_blk * BN + half * (BN / 2) + static_cast<int>(rank) * (BN / 4);
            const CUtensorMap* tb = half < 0 ? &tma_b : &tma_bh;
            const uint32_t tx = half < 0 ? 2u * kStageBytes : 2u * (kABytes + kBBytes / 2);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                if (elect_one()) {
                    // the leader arms its barrier with the bytes of BOTH CTAs; the peer's loads complete on it too
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx);
                    tma_load_2d_2sm(smem_a + stage * kABytes, &tma_a, &full_bar[stage], kb * BK, row_a);
                    tma_load_2d_2sm(smem_b + stage * kBBytes, tb, &full_bar[stage], kb * BK, row_b);
                }
                __syncwarp();
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0) {
            constexpr uint32_t idesc_whole = umma_idesc_bf16(2 * BM, BN);
            constexpr uint32_t idesc_half = umma_idesc_bf16(2 * BM, BN / 2);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int v = pair; v < sched.nv; v += n_pairs) {
                const uint32_t idesc = (sched.split && v >= sched.full) ? idesc_half : idesc_whole;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem_a + stage * kABytes));
                    const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem_b + stage * kBBytes));
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_bf16_ss_2sm(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                                             idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_commit_2sm(&empty_bar[stage]);
                        if (kb == num_kb - 1) umma_commit_2sm(&tfull_bar[acc]);
                    }
                    __syncwarp();
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else if (warp < 2 + 4 * G) {
        // ===================== epilogue warps, both CTAs: own 128 x 256 half =====================
        // G == 2: warps 2..5 take the tile's left 128 columns, warps 6..9 the right 128 (own staging buffers, own
        // store thread, own named barriers) -- doubles the epilogue throughput for the short-K / fp32-output GEMMs
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        constexpr bool kF32 = epi_out_f32<EPI>() || EPI == EPI_PATCH;
        constexpr int kUnitCols = kF32 ? 32 : 64;
        constexpr int kUnits = BN / kUnitCols / G;          // units per group (whole tile)
        static_assert(kUnits >= 2, "half tiles need at least one unit per group");
        const int row = q * 32 + lane;
        const int sw = row & 7;
        const bool store_warp = (((warp - 2) & 3) == 0);
        uint8_t* smem_cg = smem_c + grp * kBufs * kStoreBuf;
        const int bar_a = 1 + 2 * grp, bar_b = 2 + 2 * grp;
        uint32_t ucount = 0;
        int pending_rb = -1;
        // RL: the store warp walks the same (tile, unit) sequence two units AHEAD and TMA-loads the residual tiles
        int pf_v = pair, pf_uu = 0, pf_units = 0, pf_m0 = 0, pf_ncol0 = 0;
        auto pf_tile = [&]() {
            if (pf_v < sched.nv) {
                int m_blk, n_blk, half;
                sched.decode(pf_v, m_blk, n_blk, half);
                pf_m0 = m_blk * 2 * BM + static_cast<int>(rank) * BM;
                pf_units = half < 0 ? kUnits : kUnits / 2;
                pf_ncol0 = n_blk * BN + (half < 0 ? 0 : half * (BN / 2));
            }
        };
        auto pf_issue = [&](uint32_t g) {   // load the residual tile of the lookahead unit into its buffer, advance
            if (pf_v < sched.nv) {
                if (elect_one()) {
                    const uint32_t b = g & static_cast<uint32_t>(kBufs - 1);
                    mbar_arrive_expect_tx(&ld_bar[b], kStoreBuf);
                    tma_load_2d(smem_cg + b * kStoreBuf, &tma_c, &ld_bar[b], pf_ncol0 + pf_uu * kUnitCols, pf_m0);
                }
                __syncwarp();
                if (++pf_uu == pf_units) {
                    pf_uu = 0;
                    pf_v += n_pairs;
                    pf_tile();
                }
            }
        };
        if constexpr (RL != 0) {
            if (store_warp) {
                pf_tile();
                pf_issue(0);
                if constexpr (kLook == 2) pf_issue(1);
            }
        }
        // FOLD consumer: the row's eight (sum, sum of squares) partials are fetched ONE TILE AHEAD (64 bytes per thread from
        // L2: on the critical path they cost ~0.5 us per tile, which an epilogue-bound short-K GEMM cannot hide)
        [[maybe_unused]] float4 st_next[4];
        auto stats_fetch = [&](int v) {
            if constexpr (FOLD && !RL) {
                int m_blk = 0, n_blk, half;
                if (v < sched.nv) sched.decode(v, m_blk, n_blk, half);
                const int grow = m_blk * 2 * BM + static_cast<int>(rank) * BM + row;
                const float4* sp = reinterpret_cast<const float4*>(ep.ln_stats + static_cast<size_t>(grow < M ? grow : 0) * 16);
#pragma unroll
                for (int j = 0; j < 4; ++j) st_next[j] = __ldcg(sp + j);
            }
        };
        stats_fetch(pair);
        // FOLD producer: the row's centre (EpiArgs::shift_in), fetched one tile ahead like the statistics
        [[maybe_unused]] float sh_next = 0.f;
        auto shift_fetch = [&](int v) {
            if constexpr (kXb) {
                int m_blk = 0, n_blk, half;
                if (v < sched.nv) sched.decode(v, m_blk, n_blk, half);
                const int grow = m_blk * 2 * BM + static_cast<int>(rank) * BM + row;
                sh_next = (ep.shift_in != nullptr && grow < M) ? __ldcg(ep.shift_in + grow) : 0.f;
            }
        };
        shift_fetch(pair);
        // FOLD consumer: the unit's 64 folded-bias values and 64 column sums are staged in shared memory, fetched ONE UNIT
        // AHEAD by the group's first 32 threads (16 bytes each) -- the warp-uniform loads they replace cost an L2 round trip
        // per 32-column chunk on the epilogue's critical path
        [[maybe_unused]] float4* bc_s = smem_bc + grp * 32;
        [[maybe_unused]] float4 bc_next = make_float4(0.f, 0.f, 0.f, 0.f);
        [[maybe_unused]] int bc_v = pair, bc_uu = 0, bc_units = 0, bc_col0 = 0;
        auto bc_tile = [&]() {
            if (bc_v < sched.nv) {
                int mb, nb, hf;
                sched.decode(bc_v, mb, nb, hf);
                bc_units = hf < 0 ? kUnits : kUnits / 2;
                bc_col0 = nb * BN + (hf < 0 ? 0 : hf * (BN / 2)) + grp * bc_units * kUnitCols;
            }
        };
        auto bc_fetch = [&]() {
            if constexpr (FOLD && !RL) {
                if (bc_v < sched.nv) {
                    if (row < 32)
                        bc_next = __ldg(reinterpret_cast<const float4*>((row < 16 ? ep.bias : ep.ln_c) + bc_col0 +
                                                                        bc_uu * kUnitCols + 4 * (row & 15)));
                    if (++bc_uu == bc_units) {
                        bc_uu = 0;
                        bc_v += n_pairs;
                        bc_tile();
                    }
                }
            }
        };
        if constexpr (FOLD && !RL) {
            bc_tile();
            bc_fetch();
        }
        for (int v = pair; v < sched.nv; v += n_pairs) {
            int m_blk, n_blk, half;
            sched.decode(v, m_blk, n_blk, half);
            const int m0 = m_blk * 2 * BM + static_cast<int>(rank) * BM;
            // a half tile has half the column units; its accumulator sits in the first 128 TMEM columns of the buffer
            const int units = half < 0 ? kUnits : kUnits / 2;
            const int u0 = grp * units;
            const int ncol0 = n_blk * BN + (half < 0 ? 0 : half * (BN / 2));
            [[maybe_unused]] float f_rstd = 1.f, f_nmr = 0.f;      // consumer: this thread's row statistics
            [[maybe_unused]] float f_s = 0.f, f_ss = 0.f;          // producer: running (sum, sum of squares) of 128 columns
            if constexpr (FOLD && !RL) {
                float f_mean;
                fold_row_stats(st_next, ep.ln_eps, f_rstd, f_nmr, f_mean);
                stats_fetch(v + n_pairs);
                // the row's running mean for the next producer: one writer per row (first column tile, first half, first group)
                if (ep.shift_io != nullptr && n_blk == 0 && half <= 0 && grp == 0 && m0 + row < M)
                    ep.shift_io[m0 + row] += f_mean;
            }
            [[maybe_unused]] float f_shift = 0.f;                  // producer: this thread's row centre
            if constexpr (kXb) {
                f_shift = sh_next;
                shift_fetch(v + n_pairs);
            }
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
            for (int uu = 0; uu < units; ++uu, ++ucount) {
                const int u = u0 + uu;
                uint8_t* buf = smem_cg + (ucount & static_cast<uint32_t>(kBufs - 1)) * kStoreBuf;
                uint8_t* brow = buf + row * 128;
                if constexpr (RL != 0) {
                    if (store_warp) {
                        // the store of unit ucount-kLook has finished reading its buffer == buffer (ucount+kLook) % kBufs, and
                        // (FOLD) so have the copy warps: refill it
                        if (elect_one()) bulk_wait_read<kLook - 1>();
                        __syncwarp();
                        if constexpr (kXb && !HVLM_P_NOXB) {
                            const uint32_t prev = ucount + static_cast<uint32_t>(kLook - kBufs);    // that buffer's last unit
                            if (ucount + kLook >= static_cast<uint32_t>(kBufs))
                                mbar_wait(&xdone_bar[prev & static_cast<uint32_t>(kBufs - 1)], (prev / static_cast<uint32_t>(kBufs)) & 1u);
                        }
                        pf_issue(ucount + static_cast<uint32_t>(kLook));
                    }
                    // no CTA barrier here: the load barrier below is what says "this buffer holds the residual tile"
                } else {
                    if (store_warp) {
                        if (elect_one()) bulk_wait_read<1>();
                        __syncwarp();
                    }
                    if constexpr (FOLD) {
                        // (every thread of the group passed the previous unit's bar_b after its last read of bc_s)
                        if (row < 32) bc_s[row] = bc_next;
                        bc_fetch();
                    }
                    named_bar_sync(bar_a, 128);
                }
                const int n0 = ncol0 + u * kUnitCols;
#pragma unroll
                for (int h = 0; h < kUnitCols / 32; ++h) {
                    uint32_t r[32];
                    tmem_ld32(t_row + static_cast<uint32_t>(u * kUnitCols + h * 32), r);
                    tmem_ld_wait();
                    if (HVLM_P_EARLY && uu == units - 1 && h == kUnitCols / 32 - 1) {
                        // last TMEM read of this accumulator by this group: release it to the (leader's) MMA warp now,
                        // the values are in registers
                        tc_fence_before();
                        if (rank == 0) mbar_arrive(&tempty_bar[acc]);
                        else mbar_arrive_cluster(&tempty_bar[acc], 0);
                    }
                    float v[32];
                    if constexpr (FOLD && !RL) epilogue_math_fold<EPI>(r, bc_s + h * 8, bc_s + 16 + h * 8, f_rstd, f_nmr, v);
                    else epilogue_math<EPI>(r, ep.bias, n0 + h * 32, v);
                    if constexpr (RL) {
                        mbar_wait(&ld_bar[ucount & static_cast<uint32_t>(kBufs - 1)], (ucount / static_cast<uint32_t>(kBufs)) & 1u);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4* p4 = reinterpret_cast<float4*>(brow + ((j ^ sw) << 4));
                            float4 x = *p4;
                            x.x += v[4 * j];
                            x.y += v[4 * j + 1];
                            x.z += v[4 * j + 2];
                            x.w += v[4 * j + 3];
                            *p4 = x;
                            if constexpr (FOLD) {
                                const float d0 = x.x - f_shift, d1 = x.y - f_shift, d2 = x.z - f_shift, d3 = x.w - f_shift;
                                f_s += (d0 + d1) + (d2 + d3);
                                f_ss = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, f_ss))));
                            }
                        }
                        if constexpr (FOLD) {
                            // every 4th unit (32 columns each) closes a 128-column block of the row statistics
                            if ((u & 3) == 3) {
                                if (m0 + row < M)
                                    *reinterpret_cast<float2*>(ep.stats_out + static_cast<size_t>(m0 + row) * 16 + (n0 >> 7) * 2) =
                                        make_float2(f_s, f_ss);
                                f_s = f_ss = 0.f;
                            }
                        }
                    } else if constexpr (kF32) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<float4*>(brow + ((j ^ sw) << 4)) =
                                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 w;
                            w.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
                            w.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
                            w.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
                            w.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
                            *reinterpret_cast<uint4*>(brow + (((h * 4 + j) ^ sw) << 4)) = w;
                        }
                    }
                }
                if constexpr (kXb && !HVLM_P_NOXB) mbar_arrive(&xfull_bar[ucount & static_cast<uint32_t>(kBufs - 1)]);
                if (!HVLM_P_EARLY && uu == units - 1) {
                    // last TMEM read of this accumulator by this group: release it to the (leader's) MMA warp
                    tc_fence_before();
                    if (rank == 0) mbar_arrive(&tempty_bar[acc]);
                    else mbar_arrive_cluster(&tempty_bar[acc], 0);
                }
                fence_proxy_async_smem();
                named_bar_sync(bar_b, 128);
                if (store_warp) {
                    if (elect_one()) {
                        if constexpr (RL) {
                            tma_store_2d(&tma_c, buf, n0, m0);          // residual + acc + bias, summed in shared memory
                        } else if constexpr (EPI == EPI_RESID_F32) {
                            tma_reduce_add_2d(&tma_c, buf, n0, m0);
                        } else if constexpr (EPI == EPI_PATCH) {
                            // hidden [frame][257][1024]: this CTA's 128 patch rows land behind the frame's CLS row
                            tma_store_3d(&tma_c, buf, n0, 1 + (m0 & 255), m0 >> 8);
                        } else if constexpr (EPI == EPI_QKV_HM) {
                            tma_store_3d(&tma_c, buf, 0, m0, n0 >> 6);
                        } else {
                            tma_store_2d(&tma_c, buf, n0, m0);
                        }
                        bulk_commit();
                    }
                    __syncwarp();
                }
            }
            if constexpr (LN) {
                // publish the PREVIOUS tile's row block: its reduce-adds are complete once at most this tile's
                // kUnits groups are still pending (keeps the store pipeline running)
                if (store_warp) {
                    if (elect_one()) {
                        bulk_wait<kUnits>();
                        if (pending_rb >= 0) {
                            __threadfence();
                            atomicAdd(ep.ln_count + pending_rb, 1);
                        }
                    }
                    __syncwarp();
                }
                pending_rb = (m0 < M) ? (m0 / BM) : -1;
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
        if (store_warp) {
            if (elect_one()) {
                bulk_wait<0>();
                if constexpr (LN) {
                    if (pending_rb >= 0) {
                        __threadfence();
                        atomicAdd(ep.ln_count + pending_rb, 1);
                    }
                }
            }
            __syncwarp();
        }
    } else {
        // ===================== copy warps (6..9), only when RL + FOLD: the bf16 twin of every residual unit =====================
        // They read the fp32 sums back from the staging buffer TRANSPOSED -- 8 lanes take one row's 128 bytes (conflict-free
        // under the 128-byte swizzle) and write its 64 bf16 bytes contiguously, 4 rows per instruction -- off the epilogue
        // warps' critical path.  (Measured per out_proj call, stand-alone: this read-back inside the epilogue warps +15 us;
        // 16-byte stores from one row per lane +23 us; a second TMA store needs staging memory the K pipeline cannot
        // spare: 4 stages instead of 5 cost the K = 4096 GEMM 20 us.)
        if constexpr (kXb && !HVLM_P_NOXB) {
            constexpr int kUnits = BN / 32;
            const int cw = warp - 6;
            const int cj = lane & 7;
            uint32_t ucount = 0;
            for (int v = pair; v < sched.nv; v += n_pairs) {
                int m_blk, n_blk, half;
                sched.decode(v, m_blk, n_blk, half);
                const int m0 = m_blk * 2 * BM + static_cast<int>(rank) * BM;
                const int units = half < 0 ? kUnits : kUnits / 2;
                const int ncol0 = n_blk * BN + (half < 0 ? 0 : half * (BN / 2));
                float sh[8];            // the centres of this lane's eight rows of the tile
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int grow = m0 + cw * 32 + 4 * i + (lane >> 3);
                    sh[i] = (ep.shift_in != nullptr && grow < M) ? __ldcg(ep.shift_in + grow) : 0.f;
                }
#pragma unroll 1
                for (int uu = 0; uu < units; ++uu, ++ucount) {
                    const uint32_t b = ucount & static_cast<uint32_t>(kBufs - 1);
                    const uint8_t* buf = smem_c + b * kStoreBuf;
                    mbar_wait(&xfull_bar[b], (ucount / static_cast<uint32_t>(kBufs)) & 1u);
                    __nv_bfloat16* xo = static_cast<__nv_bfloat16*>(ep.xb_out) + ncol0 + uu * 32 + 4 * cj;
                    float4 x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = cw * 32 + 4 * i + (lane >> 3);
                        x[i] = *reinterpret_cast<const float4*>(buf + rr * 128 + ((cj ^ (rr & 7)) << 4));
                    }
                    mbar_arrive(&xdone_bar[b]);       // the values are in registers: the buffer may be refilled
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = cw * 32 + 4 * i + (lane >> 3);
                        if (m0 + rr < M)
                            *reinterpret_cast<uint2*>(xo + static_cast<size_t>(m0 + rr) * N) =
                                make_uint2(pack_bf16(x[i].x - sh[i], x[i].y - sh[i]), pack_bf16(x[i].z - sh[i], x[i].w - sh[i]));
                    }
                }
            }
        }
        // ===================== LayerNorm warps (6..9), only when LN: normalise finished 128-row blocks =====================
        if constexpr (LN) {
            const int lw = warp - 6;
            const int num_rb = (M + BM - 1) / BM;
            const float* hidden = static_cast<const float*>(ep.out);
            __nv_bfloat16* y = static_cast<__nv_bfloat16*>(ep.ln_out);
            for (int i = blockIdx.x; i < num_rb; i += gridDim.x) {
                const int rb = ep.reverse ? num_rb - 1 - i : i;     // same direction as the tile traversal
                if (lane == 0) {
                    while (ld_acquire_gpu(ep.ln_count + rb) < num_n) __nanosleep(256);
                }
                __syncwarp();
                const int r0 = rb * BM + lw * 32;
#pragma unroll 1
                for (int r = 0; r < 32; r += 2) {
                    const int row = r0 + r;
                    const int nrows = min(2, M - row);
                    if (nrows > 0)
                        ln_rows<2>(hidden + static_cast<size_t>(row) * 1024, ep.ln_gamma, ep.ln_beta,
                                   y + static_cast<size_t>(row) * 1024, nrows, lane);
                }
                named_bar_sync(5, 128);
                if (lw == 0 && lane == 0) ep.ln_count[rb] = 0;     // leave the counters clean for the next launch
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();     // nobody exits (or frees TMEM) while the peer may still signal / be signalled
    if (warp == 1) tmem_dealloc_2sm(tmem_base);
}

template <int EPI, bool LN = false, int G = 1, int RL = 0, bool FOLD = false>
static int launch_two(const void* A, const void* B, int M, int N, int K, const EpiArgs& ep, cudaStream_t s) {
    CUtensorMap ta, tb, tbh, tc;
    {
        uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
        uint64_t str[1] = {static_cast<uint64_t>(K) * 2};
        uint32_t box[2] = {BK, BM};
        int rc = make_tmap_bf16(&ta, A, 2, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
        uint64_t str[1] = {static_cast<uint64_t>(K) * 2};
        uint32_t box[2] = {BK, BN / 2};
        int rc = make_tmap_bf16(&tb, B, 2, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
        uint64_t str[1] = {static_cast<uint64_t>(K) * 2};
        uint32_t box[2] = {BK, BN / 4};               // half tiles: 64 B rows per CTA
        int rc = make_tmap_bf16(&tbh, B, 2, dims, str, box);
        if (rc) return rc;
    }
    {
        if (!ep.out || !aligned16(ep.out)) return HVLM_ERR_ALIGN;
        uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)};
        int rc;
        if constexpr (EPI == EPI_PATCH) {
            if (N != 1024 || (M & 255) != 0) return HVLM_ERR_BAD_SHAPE;
            uint64_t d3[3] = {1024, HVLM_VIT_TOKENS, static_cast<uint64_t>(M >> 8)};
            uint64_t s3[2] = {1024 * 4, static_cast<uint64_t>(HVLM_VIT_TOKENS) * 1024 * 4};
            uint32_t b3[3] = {32, BM, 1};
            rc = make_tmap_f32(&tc, ep.out, 3, d3, s3, b3);
        } else if constexpr (EPI == EPI_QKV_HM) {
            if (N != 3072) return HVLM_ERR_BAD_SHAPE;
            rc = make_qkv_hm_tmap(&tc, ep.out, M, BM);
        } else if constexpr (epi_out_f32<EPI>()) {
            uint64_t str[1] = {static_cast<uint64_t>(N) * 4};
            uint32_t box[2] = {32, BM};
            rc = make_tmap_f32(&tc, ep.out, 2, dims, str, box);
        } else {
            uint64_t str[1] = {static_cast<uint64_t>(N) * 2};
            uint32_t box[2] = {64, BM};
            rc = make_tmap_bf16(&tc, ep.out, 2, dims, str, box);
        }
        if (rc) return rc;
    }
    auto kern = gemm2_tcgen05_kernel<EPI, LN, G, RL, FOLD>;
    constexpr int kSmemBytes = Cfg2<G, RL>::kSmemBytes;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess)
            return HVLM_ERR_CUDA;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    // HVLM_GEMM_TAIL_SPLIT=0 keeps whole tiles in the last wave (A/B runs)
    static const int split_tail = []() {
        const char* e = getenv("HVLM_GEMM_TAIL_SPLIT");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    const int allow_split = LN ? 0 : split_tail;
    const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * (N / BN);
    int pairs = num_sms() / 2;
    const int want = allow_split ? 2 * tiles : tiles;      // few tiles: cut all of them in two to occupy more pairs
    if (want < pairs) pairs = want;
    if (launch_pdl(kern, dim3(2 * pairs), dim3(kThreads + ((LN || G == 2 || (RL != 0 && FOLD)) ? 128 : 0)), kSmemBytes, s, ta, tb, tbh, tc, M, N, K, ep,
                   allow_split) != cudaSuccess) {
        cudaGetLastError();
        return HVLM_ERR_CUDA;
    }
    return check_last("gemm2");
}

}  // namespace gemm2

// returns HVLM_ERR_UNSUPPORTED when this (epilogue, shape) has no 2-CTA instantiation
int launch_gemm_2cta(int epi, const void* A, const void* B, int M, int N, int K, const EpiArgs& ep, cudaStream_t s) {
    using namespace gemm2;
    if ((N % BN) != 0) return HVLM_ERR_UNSUPPORTED;
    // The quick-GELU epilogue is math-bound on 4 warps (measured 156.7 us vs 146.0 us for the plain bf16 store on
    // 25700x4096x1024); a second 4-warp group brings it to 149.7 us.  The fp32 store / reduce-add epilogues are bound by
    // the store traffic instead and measured no gain, so they keep 4 warps and the 6-stage pipeline.
    // HVLM_GEMM_EPI_GROUPS=1 forces the single group for A/B runs.
    static const bool two_groups = []() {
        const char* e = getenv("HVLM_GEMM_EPI_GROUPS");
        return !(e && e[0] == '1');
    }();
    static const int resid_load_maxk = []() {
        const char* e = getenv("HVLM_RESID_LOAD_MAXK");
        return e ? atoi(e) : 0;
    }();
    // HVLM_QKV_EPI_GROUPS=2: the same second epilogue group for the QKV GEMM (A/B runs; see DESIGN.md for the outcome)
    static const bool qkv_two = []() {
        const char* e = getenv("HVLM_QKV_EPI_GROUPS");
        return e && e[0] == '2';
    }();
    if (ep.ln_stats != nullptr) {
        // LayerNorm folded into this GEMM (consumer side)
        if (K != 1024 || !ep.ln_c || !ep.bias || !aligned16(ep.ln_c) || !aligned16(ep.ln_stats)) return HVLM_ERR_BAD_ARG;
        switch (epi) {
            case EPI_QKV_HM:
                return qkv_two ? launch_two<EPI_QKV_HM, false, 2, false, true>(A, B, M, N, K, ep, s)
                               : launch_two<EPI_QKV_HM, false, 1, false, true>(A, B, M, N, K, ep, s);
            case EPI_BIAS_BF16: return launch_two<EPI_BIAS_BF16, false, 1, false, true>(A, B, M, N, K, ep, s);
            case EPI_GELU_BF16:
                return two_groups ? launch_two<EPI_GELU_BF16, false, 2, false, true>(A, B, M, N, K, ep, s)
                                  : launch_two<EPI_GELU_BF16, false, 1, false, true>(A, B, M, N, K, ep, s);
            default: return HVLM_ERR_UNSUPPORTED;
        }
    }
    if (ep.xb_out != nullptr) {
        // producer side of the fold: residual-load epilogue + bf16 copy + row statistics
        if (epi != EPI_RESID_F32 || N != 1024 || !ep.stats_out || !aligned16(ep.xb_out) || !aligned16(ep.stats_out))
            return HVLM_ERR_BAD_ARG;
        if (K >= 2048) return launch_two<EPI_RESID_F32, false, 1, HVLM_P_RL_LONGK, true>(A, B, M, N, K, ep, s);
        return launch_two<EPI_RESID_F32, false, 1, 4, true>(A, B, M, N, K, ep, s);
    }
    if (two_groups && ep.ln_out == nullptr && epi == EPI_GELU_BF16)
        return launch_two<EPI_GELU_BF16, false, 2>(A, B, M, N, K, ep, s);
    if (qkv_two && epi == EPI_QKV_HM) return launch_two<EPI_QKV_HM, false, 2>(A, B, M, N, K, ep, s);
    switch (epi) {
        case EPI_BIAS_BF16: return launch_two<EPI_BIAS_BF16>(A, B, M, N, K, ep, s);
        case EPI_BIAS_F32: return launch_two<EPI_BIAS_F32>(A, B, M, N, K, ep, s);
        case EPI_GELU_BF16: return launch_two<EPI_GELU_BF16>(A, B, M, N, K, ep, s);
        case EPI_GELU_F32: return launch_two<EPI_GELU_F32>(A, B, M, N, K, ep, s);
        case EPI_PATCH: return launch_two<EPI_PATCH>(A, B, M, N, K, ep, s);
        case EPI_RESID_F32:
            if (ep.ln_out != nullptr) {
                if (N != 1024 || !ep.ln_gamma || !ep.ln_beta || !ep.ln_count) return HVLM_ERR_BAD_ARG;
                return launch_two<EPI_RESID_F32, true>(A, B, M, N, K, ep, s);
            }
            // Short-K residual GEMMs (out_proj, K = 1024) are bound by their epilogue: the fp32 TMA reduce-add into the
            // residual stream sustains ~7 B/clk/SM, twice the time of the K loop.  Experiment (HVLM_RESID_LOAD_MAXK=1024;
            // default 0 = off): the epilogue TMA-loads the residual tile two units ahead, adds in shared memory and leaves
            // with a plain TMA store.  Correct (same tests), but measured SLOWER on B200, same box, 100 frames: out_proj
            // 74.3 vs 66.7 us per call, step 14.91 vs 14.52 ms -- the extra 32 KB of shared-memory traffic per unit (TMA
            // write + read-modify-write) competes with the tensor cores' operand reads; the L2-side reduce-add stays.
            if (K <= resid_load_maxk) return launch_two<EPI_RESID_F32, false, 1, 4>(A, B, M, N, K, ep, s);
            return launch_two<EPI_RESID_F32>(A, B, M, N, K, ep, s);
        case EPI_QKV_HM: return launch_two<EPI_QKV_HM>(A, B, M, N, K, ep, s);
        default: return HVLM_ERR_UNSUPPORTED;
    }
}

}  // namespace hvlm
