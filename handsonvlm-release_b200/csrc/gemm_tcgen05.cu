// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T (+ fused epilogue)
//
//   * A, B bf16, K-major (row-major with K contiguous) -- exactly nn.Linear's x[M,K] and weight[N,K]
//   * TMA (cp.async.bulk.tensor, SWIZZLE_128B) fills a 4-stage shared-memory ring of [128x64] A tiles and
//     [BNx64] B tiles; out-of-bounds rows/columns (M tail, K tail) are zero-filled by the TMA unit
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16), fp32 accumulators in
//     TMEM, double-buffered (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1
//   * 4 epilogue warps read the accumulator with tcgen05.ld (32 lanes x 32 columns), apply the epilogue
//     (bias / quick-GELU / fp32 residual via TMA reduce-add / patch-embed rows behind each frame's CLS row)
//     and store 128-bit vectors
//   * persistent: grid = min(tiles, #SM); tiles are walked N-fastest so concurrently running CTAs share
//     A rows and the (small, L2-resident) weight matrix
//
// Replaces every nn.Linear on the path (see include/hvlm_b200.h).
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "gemm_epilogue.cuh"
#include "hvlm_internal.cuh"
#include "hvlm_ptx.cuh"

namespace hvlm {

// ------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_encode_once;

static void load_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
        g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
}

// Descriptor cache.  A forward of the tower encodes ~4 tensor maps for each of its ~95 GEMM / attention launches, and the
// same (pointer, shape, box) tuples come back every step (weights never move, the caching allocator hands the workspace and
// the residual stream out at the same addresses): a CUtensorMap is a pure function of its arguments (it holds an address,
// not a reference), so the encoded 128 bytes are memoised.  HVLM_TMAP_CACHE=0 disables it (A/B runs).
struct TmapKey {
    uint64_t base, dims[5], str[4];
    uint32_t box[5], rank, dt, swz;
    bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
        uint64_t h = 0xcbf29ce484222325ull;
        for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ w[i]) * 0x100000001b3ull;
        return static_cast<size_t>(h ^ (h >> 29));
    }
};
static_assert(sizeof(TmapKey) % 8 == 0, "hashed as 64-bit words");
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
constexpr size_t kTmapCacheMax = 8192;      // ~2 MB of host memory; cleared wholesale when full

static int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    std::call_once(g_encode_once, load_encode);
    if (!g_encode) return HVLM_ERR_CUDA;
    static const bool use_cache = []() {
        const char* e = getenv("HVLM_TMAP_CACHE");
        return !(e && e[0] == '0');
    }();
    TmapKey key;
    if (use_cache) {
        memset(&key, 0, sizeof(key));
        key.base = reinterpret_cast<uint64_t>(base);
        key.rank = static_cast<uint32_t>(rank);
        key.dt = static_cast<uint32_t>(dt);
        key.swz = static_cast<uint32_t>(swz);
        for (int i = 0; i < rank; ++i) {
            key.dims[i] = dims[i];
            key.box[i] = box[i];
            if (i > 0) key.str[i - 1] = strides_bytes[i - 1];
        }
        std::lock_guard<std::mutex> lk(g_tmap_mu);
        auto it = g_tmap_cache.find(key);
        if (it != g_tmap_cache.end()) {
            *out = it->second;
            return HVLM_OK;
        }
    }
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t bx[5];
    cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
    }
    CUresult r = g_encode(out, dt, static_cast<cuuint32_t>(rank),
                          const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS && getenv("HVLM_DEBUG"))
        fprintf(stderr, "[hvlm] cuTensorMapEncodeTiled failed: CUresult %d (rank %d)\n", static_cast<int>(r), rank);
    if (r == CUDA_SUCCESS && use_cache) {
        std::lock_guard<std::mutex> lk(g_tmap_mu);
        if (g_tmap_cache.size() >= kTmapCacheMax) g_tmap_cache.clear();
        g_tmap_cache.emplace(key, *out);
    }
    return r == CUDA_SUCCESS ? HVLM_OK : HVLM_ERR_CUDA;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box) {
    return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
    return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box);
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int check_last(const char* what) {
    count_launch();
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && getenv("HVLM_DEBUG")) fprintf(stderr, "[hvlm] %s: %s\n", what, cudaGetErrorString(e));
    return e == cudaSuccess ? HVLM_OK : HVLM_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kGemmThreads = 192;   // warp0 TMA, warp1 MMA(+TMEM alloc), warps 2..5 epilogue

template <int BN>
struct GemmCfg {
    static constexpr int kStages = (BN == 256) ? 4 : 6;
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = 2 * BN;   // double-buffered accumulator: 256 or 512 columns
    static constexpr int kStoreBuf = BM * 128;   // one staging unit: 128 rows x 128 B (SWIZZLE_128B box)
    static constexpr int kSmemBytes = kStages * kStageBytes + 2 * kStoreBuf + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&acc)[32], int m, int n0, int M, int N,
                                               const EpiArgs& ep) {
    if (m >= M) return;
    float v[32];
    if constexpr (EPI != EPI_PATCH) {
        if (ep.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(ep.bias + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 b = __ldg(b4 + j);
                v[4 * j + 0] = __uint_as_float(acc[4 * j + 0]) + b.x;
                v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + b.y;
                v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + b.z;
                v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + b.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
        }
    }

    if constexpr (EPI == EPI_GELU_BF16 || EPI == EPI_GELU_F32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
    }
    if constexpr (EPI == EPI_RESID_F32) {
        const float4* r4 = reinterpret_cast<const float4*>(ep.resid + static_cast<size_t>(m) * N + n0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 r = r4[j];
            v[4 * j + 0] += r.x;
            v[4 * j + 1] += r.y;
            v[4 * j + 2] += r.z;
            v[4 * j + 3] += r.w;
        }
    }

    if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_GELU_BF16) {
        uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(ep.out) + static_cast<size_t>(m) * N + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 w;
            w.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
            w.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
            w.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
            w.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
            o[j] = w;
        }
    } else if constexpr (EPI == EPI_BIAS_F32 || EPI == EPI_RESID_F32 || EPI == EPI_GELU_F32) {
        float4* o = reinterpret_cast<float4*>(static_cast<float*>(ep.out) + static_cast<size_t>(m) * N + n0);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else if constexpr (EPI == EPI_PATCH) {
        const int f = m >> 8;
        const int p = m & 255;
        float4* o = reinterpret_cast<float4*>(static_cast<float*>(ep.out) +
                                              (static_cast<size_t>(f) * HVLM_VIT_TOKENS + 1 + p) * N + n0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            o[j] = make_float4(__uint_as_float(acc[4 * j + 0]), __uint_as_float(acc[4 * j + 1]),
                               __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
    }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_c, int M, int N, int K, EpiArgs ep) {
    using Cfg = GemmCfg<BN>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * Cfg::kABytes;
    uint8_t* smem_c = smem + kStages * Cfg::kStageBytes;   // 2 staging units for the TMA-store epilogue
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + 2 * Cfg::kStoreBuf);
    uint64_t* full_bar = bars;                       // [kStages]  TMA -> MMA
    uint64_t* empty_bar = bars + kStages;            // [kStages]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * kStages;        // [2]        MMA -> epilogue
    uint64_t* tempty_bar = bars + 2 * kStages + 2;   // [2]        epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_m = (M + BM - 1) / BM;
    const int num_n = N / BN;
    const int num_tiles = num_m * num_n;
    const int num_kb = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        if constexpr (epi_is_staged<EPI>()) tma_prefetch_desc(&tma_c);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 128);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();              // everything above overlapped the previous kernel's tail

    if (warp == 0) {
        // ===================== TMA producer =====================
        // The whole warp walks the loop (warp-uniform control flow and operands); one elected lane issues.
        int stage = 0;
        uint32_t phase = 0;
        for (int t_ = blockIdx.x; t_ < num_tiles; t_ += gridDim.x) {
            const int tile = ep.reverse ? num_tiles - 1 - t_ : t_;
            const int m_blk = tile / num_n;
            const int n_blk = tile - m_blk * num_n;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    tma_load_2d(smem_a + stage * Cfg::kABytes, &tma_a, &full_bar[stage], kb * BK, m_blk * BM);
                    tma_load_2d(smem_b + stage * Cfg::kBBytes, &tma_b, &full_bar[stage], kb * BK, n_blk * BN);
                }
                __syncwarp();
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // Whole warp in the loop, tcgen05.mma / tcgen05.commit issued by one elected lane (keeps the descriptors
        // in uniform registers and avoids a per-instruction divergence loop).
        constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
                const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // +32 bytes (encoded >>4 => +2) per 16-element K step inside the 128-byte swizzle row
                        umma_bf16_ss(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                                     idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);   // frees the smem slot once these MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);   // accumulator complete -> epilogue
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
    } else {
        // ===================== epilogue warps (2..5) =====================
        const int q = warp & 3;   // TMEM lane quarter this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        if constexpr (epi_is_staged<EPI>()) {
            // TMEM -> registers -> (bias / GELU) -> swizzled smem staging unit -> TMA store (or TMA reduce-add
            // into the fp32 residual stream).  Global traffic is fully coalesced and asynchronous; the M tail
            // is clipped by the TMA unit.
            constexpr bool kF32 = epi_out_f32<EPI>();
            constexpr int kUnitCols = kF32 ? 32 : 64;            // 128 bytes per row either way
            constexpr int kUnits = BN / kUnitCols;
            const int row = q * 32 + lane;
            const int sw = row & 7;
            const bool store_warp = (warp == 2);   // one elected lane of this warp issues the TMA stores
            uint32_t ucount = 0;
            for (int t_ = blockIdx.x; t_ < num_tiles; t_ += gridDim.x) {
                const int tile = ep.reverse ? num_tiles - 1 - t_ : t_;
                const int m_blk = tile / num_n;
                const int n_blk = tile - m_blk * num_n;
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
                for (int u = 0; u < kUnits; ++u, ++ucount) {
                    uint8_t* buf = smem_c + (ucount & 1u) * Cfg::kStoreBuf;
                    uint8_t* brow = buf + row * 128;
                    if (store_warp) {
                        if (elect_one()) bulk_wait_read<1>();     // the store that last used this buffer is done
                        __syncwarp();
                    }
                    named_bar_sync(1, 128);
                    const int n0 = n_blk * BN + u * kUnitCols;
#pragma unroll
                    for (int h = 0; h < kUnitCols / 32; ++h) {
                        uint32_t r[32];
                        tmem_ld32(t_row + static_cast<uint32_t>(u * kUnitCols + h * 32), r);
                        tmem_ld_wait();
                        float v[32];
                        epilogue_math<EPI>(r, ep.bias, n0 + h * 32, v);
                        if constexpr (kF32) {
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
                    if (u == kUnits - 1) {
                        // last TMEM read of this accumulator: hand it back to the MMA warp
                        tc_fence_before();
                        mbar_arrive(&tempty_bar[acc]);
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(2, 128);
                    if (store_warp) {
                      if (elect_one()) {
                        if constexpr (EPI == EPI_RESID_F32) {
                            tma_reduce_add_2d(&tma_c, buf, n0, m_blk * BM);
                        } else if constexpr (EPI == EPI_QKV_HM) {
                            // column-block-major [48][M][64]: the 64-column unit is one column block
                            tma_store_3d(&tma_c, buf, 0, m_blk * BM, n0 >> 6);
                        } else {
                            tma_store_2d(&tma_c, buf, n0, m_blk * BM);
                        }
                        bulk_commit();
                      }
                      __syncwarp();
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
            if (store_warp) {
                if (elect_one()) bulk_wait<0>();
                __syncwarp();
            }
        } else {
            for (int t_ = blockIdx.x; t_ < num_tiles; t_ += gridDim.x) {
                const int tile = ep.reverse ? num_tiles - 1 - t_ : t_;
                const int m_blk = tile / num_n;
                const int n_blk = tile - m_blk * num_n;
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                const int m = m_blk * BM + q * 32 + lane;
                const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(t_row + static_cast<uint32_t>(c * 32), r);
                    tmem_ld_wait();
                    epilogue_chunk<EPI>(r, m, n_blk * BN + c * 32, M, N, ep);
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar[acc]);
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------------------
template <int BN, int EPI>
static int launch_one(const void* A, const void* B, int M, int N, int K, const EpiArgs& ep, cudaStream_t s) {
    using Cfg = GemmCfg<BN>;
    CUtensorMap ta, tb;
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
        uint32_t box[2] = {BK, BN};
        int rc = make_tmap_bf16(&tb, B, 2, dims, str, box);
        if (rc) return rc;
    }
    CUtensorMap tc = tb;   // unused by the direct-store epilogues
    if constexpr (epi_is_staged<EPI>()) {
        if (!ep.out || !aligned16(ep.out)) return HVLM_ERR_ALIGN;
        uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)};
        int rc;
        if constexpr (EPI == EPI_QKV_HM) {
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
    auto kern = gemm_tcgen05_kernel<BN, EPI>;
    static bool attr_set[64] = {false};   // per instantiation, per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
            return HVLM_ERR_CUDA;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int tiles = ((M + BM - 1) / BM) * (N / BN);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    if (launch_pdl(kern, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, s, ta, tb, tc, M, N, K, ep) != cudaSuccess) {
        cudaGetLastError();
        return HVLM_ERR_CUDA;
    }
    return check_last("gemm");
}

int launch_gemm_2cta(int epi, const void* A, const void* B, int M, int N, int K, const EpiArgs& ep, cudaStream_t s);

int launch_gemm(int epi, const void* A, const void* B, int M, int N, int K, const EpiArgs& ep, cudaStream_t s) {
    if (!A || !B || M <= 0 || N <= 0 || K <= 0) return HVLM_ERR_BAD_ARG;
    if ((N % 128) != 0 || (K % 8) != 0) return HVLM_ERR_BAD_SHAPE;
    if (!aligned16(A) || !aligned16(B)) return HVLM_ERR_ALIGN;
    // preferred path: CTA pairs (cta_group::2, 256x256 tiles); HVLM_GEMM_1CTA=1 selects the single-CTA kernel (A/B runs)
    static const bool force_1cta = []() {
        const char* e = getenv("HVLM_GEMM_1CTA");
        return e && e[0] == '1';
    }();
    if (!force_1cta) {
        const int rc2 = launch_gemm_2cta(epi, A, B, M, N, K, ep, s);
        if (rc2 != HVLM_ERR_UNSUPPORTED) return rc2;
    }
    if (ep.ln_stats != nullptr || ep.xb_out != nullptr) return HVLM_ERR_UNSUPPORTED;   // the folded LayerNorm lives in the 2-CTA kernel
    const bool wide = (N % 256) == 0;
#define HVLM_GEMM_CASE(E)                                                         \
    case E:                                                                       \
        return wide ? launch_one<256, E>(A, B, M, N, K, ep, s) : launch_one<128, E>(A, B, M, N, K, ep, s);
    switch (epi) {
        HVLM_GEMM_CASE(EPI_BIAS_BF16)
        HVLM_GEMM_CASE(EPI_BIAS_F32)
        HVLM_GEMM_CASE(EPI_GELU_BF16)
        HVLM_GEMM_CASE(EPI_RESID_F32)
        HVLM_GEMM_CASE(EPI_GELU_F32)
        case EPI_PATCH:
            return launch_one<256, EPI_PATCH>(A, B, M, N, K, ep, s);
        case EPI_QKV_HM:
            return launch_one<256, EPI_QKV_HM>(A, B, M, N, K, ep, s);
        default:
            return HVLM_ERR_BAD_ARG;
    }
#undef HVLM_GEMM_CASE
}

}  // namespace hvlm

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int hvlm_gemm_bf16(const void* A, const void* B, const float* bias, const float* resid, void* out, int M,
                              int N, int K, int epilogue, int out_dtype, void* stream) {
    using namespace hvlm;
    if (!out) return HVLM_ERR_BAD_ARG;
    if (!aligned16(out) || (bias && !aligned16(bias)) || (resid && !aligned16(resid))) return HVLM_ERR_ALIGN;
    if (out_dtype != HVLM_BF16 && out_dtype != HVLM_F32) return HVLM_ERR_BAD_DTYPE;
    EpiArgs ep;
    ep.bias = bias;
    ep.resid = resid;
    ep.out = out;
    int epi;
    switch (epilogue) {
        case HVLM_EPI_BIAS:
            epi = out_dtype == HVLM_BF16 ? EPI_BIAS_BF16 : EPI_BIAS_F32;
            break;
        case HVLM_EPI_BIAS_QUICKGELU:
            epi = out_dtype == HVLM_BF16 ? EPI_GELU_BF16 : EPI_GELU_F32;
            break;
        case HVLM_EPI_BIAS_RESIDUAL:
            if (!resid) return HVLM_ERR_BAD_ARG;
            if (out_dtype != HVLM_F32) return HVLM_ERR_BAD_DTYPE;
            // the kernel accumulates into the fp32 residual stream in place (TMA reduce-add)
            if (resid != out &&
                cudaMemcpyAsync(out, resid, static_cast<size_t>(M) * N * 4, cudaMemcpyDeviceToDevice,
                                static_cast<cudaStream_t>(stream)) != cudaSuccess)
                return HVLM_ERR_CUDA;
            epi = EPI_RESID_F32;
            break;
        default:
            return HVLM_ERR_BAD_ARG;
    }
    StageTimer st(HVLM_STAGE_GEMM, static_cast<cudaStream_t>(stream));
    return launch_gemm(epi, A, B, M, N, K, ep, static_cast<cudaStream_t>(stream));
}

extern "C" int hvlm_gemm_ln_fold_bf16(const void* xb, const float* stats, const void* w_f, const float* c, const float* b_f,
                                      void* out, int M, int N, int epilogue, int qkv_hm, float eps, float* shift_io,
                                      void* stream) {
    using namespace hvlm;
    if (!xb || !stats || !w_f || !c || !b_f || !out || M <= 0 || N <= 0) return HVLM_ERR_BAD_ARG;
    if (!aligned16(out) || !aligned16(stats) || !aligned16(c) || !aligned16(b_f)) return HVLM_ERR_ALIGN;
    if ((N % 256) != 0 || (qkv_hm && N != 3072)) return HVLM_ERR_BAD_SHAPE;
    int epi;
    if (epilogue == HVLM_EPI_BIAS) epi = qkv_hm ? EPI_QKV_HM : EPI_BIAS_BF16;
    else if (epilogue == HVLM_EPI_BIAS_QUICKGELU && !qkv_hm) epi = EPI_GELU_BF16;
    else return HVLM_ERR_BAD_ARG;
    EpiArgs ep;
    ep.bias = b_f;
    ep.ln_c = c;
    ep.ln_stats = stats;
    ep.ln_eps = eps;
    ep.shift_io = shift_io;
    ep.out = out;
    StageTimer st(HVLM_STAGE_GEMM, static_cast<cudaStream_t>(stream));
    return launch_gemm(epi, xb, w_f, M, N, 1024, ep, static_cast<cudaStream_t>(stream));
}

extern "C" int hvlm_gemm_resid_stats(const void* A, const void* B, const float* bias, float* hidden, const float* shift,
                                     void* xb_out, float* stats_out, int M, int K, void* stream) {
    using namespace hvlm;
    if (!hidden || !xb_out || !stats_out) return HVLM_ERR_BAD_ARG;
    if (!aligned16(hidden) || !aligned16(xb_out) || !aligned16(stats_out) || (bias && !aligned16(bias))) return HVLM_ERR_ALIGN;
    EpiArgs ep;
    ep.bias = bias;
    ep.resid = hidden;
    ep.out = hidden;
    ep.xb_out = xb_out;
    ep.stats_out = stats_out;
    ep.shift_in = shift;
    StageTimer st(HVLM_STAGE_GEMM, static_cast<cudaStream_t>(stream));
    return launch_gemm(EPI_RESID_F32, A, B, M, 1024, K, ep, static_cast<cudaStream_t>(stream));
}
