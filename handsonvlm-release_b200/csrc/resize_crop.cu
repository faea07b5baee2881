// CLIPImageProcessor resize (shortest edge -> 224, PIL BICUBIC) + centre crop on raw uint8 frames.
//
// Replaces the CPU side of hoi_forecast/dataset/video_utils.py:28-53 (`processor.preprocess(image)` per frame;
// transformers==4.31.0 CLIPImageProcessor -> PIL.Image.resize(BICUBIC) -> centre crop).  The rescale + normalise steps that
// follow in the processor are fused into the patch extraction of the tower (hvlm_vit_l14_fwd_u8), so a decoded clip goes
// uint8 [N,H,W,3] -> uint8 [N,224,224,3] -> ViT without ever becoming a float image.
//
// Pillow's 8-bit resampler (third party, src/libImaging/Resample.c of Pillow 12.x, the version installed next to the
// reference's pinned transformers) restated -- published algorithm:
//   precompute_coeffs : per output index xx: center = (xx + 0.5) * scale, support = 2 * max(scale, 1),
//                       xmin = int(center - support + 0.5) clipped to 0, xmax = int(center + support + 0.5) clipped to
//                       in_size, taps w(x) = bicubic((x + xmin - center + 0.5) / max(scale, 1)), a = -0.5, normalised to 1
//   normalize_coeffs_8bpc : k = int(w * 2^22 +- 0.5)      (PRECISION_BITS = 32 - 8 - 2)
//   ImagingResampleHorizontal_8bpc, then ...Vertical_8bpc : acc = 2^21 + sum pix * k ; out = clip8(acc >> 22)
//                       -- the intermediate image is uint8: rounding and clipping happen after EACH pass.
// Only the centre-crop window is computed.  Results are bit-identical to PIL (tests/test_oracle_golden.py pins the host
// tables + a numpy restatement to PIL itself; tests/test_gpu_round2.py pins the kernel).
//
// The arithmetic is 2 passes x ~7 exact 32-bit integer multiply-adds per output byte (22-bit coefficients: neither fp32
// nor the tensor cores can carry it), i.e. the kernel is bound by the integer pipe, not by HBM (32 MB per 100-frame clip).
// CTA = (block of kRows output rows, frame):
//   phase 0  coefficient tables and the source window (only the rows / columns the crop reads) -> shared memory with
//            16-byte global loads
//   phase 1  horizontal pass, shared -> shared (uint8 [rows][out_w*3])
//   phase 2  vertical pass from shared memory, four output bytes per thread, 32-bit coalesced stores
// Frames whose window does not fit in shared memory (very large sources) take the staging-free kernel below.
#include <cmath>
#include <vector>

#include "hvlm_internal.cuh"

namespace hvlm {
namespace resize {

constexpr int kRows = 8;
constexpr int kThreads = 256;
constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr size_t kSmemLimit = 100 * 1024;      // two CTAs per SM

static double bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

__device__ __forceinline__ uint32_t clip8(int acc) {
    const int v = acc >> kPrecisionBits;     // arithmetic shift == Pillow's table index
    return static_cast<uint32_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// table layout (int32): xb [2*out_w] | xc [out_w*xk] | yb [2*out_h] | yc [out_h*yk]
struct Tab {
    const int32_t *xb, *xc, *yb, *yc;
};
__host__ __device__ inline Tab carve(const int32_t* t, const hvlm_resize_plan& p) {
    Tab r;
    r.xb = t;
    r.xc = r.xb + 2 * p.out_w;
    r.yb = r.xc + p.out_w * p.xk;
    r.yc = r.yb + 2 * p.out_h;
    return r;
}

// KMAX: compile-time bound on the horizontal taps (xk <= KMAX): a thread keeps the coefficients of ITS output column in
// registers and walks the rows, reading the 3*xk source bytes of a row as aligned 32-bit words
template <int KMAX>
__global__ void __launch_bounds__(kThreads)
resize_crop_u8_kernel(const uint8_t* __restrict__ src, size_t src_bytes, const __grid_constant__ hvlm_resize_plan p,
                      const int32_t* __restrict__ table, uint8_t* __restrict__ dst, int src_pitch) {
    extern __shared__ __align__(16) uint8_t smem[];
    const Tab t = carve(table, p);
    const int out_w = p.out_w, xk = p.xk, yk = p.yk, pitch = out_w * 3;
    int32_t* xb_s = reinterpret_cast<int32_t*>(smem);                 // [2*out_w]
    int32_t* xc_s = xb_s + 2 * out_w;                                 // [out_w*xk]
    int32_t* yb_s = xc_s + out_w * xk;                                // [2*kRows]
    int32_t* yc_s = yb_s + 2 * kRows;                                 // [kRows*yk]
    int32_t* off_s = yc_s + kRows * yk;                               // [rows_cap] byte offset of x_lo inside each staged row
    const int n_ints = (2 * out_w + out_w * xk + 2 * kRows + kRows * yk + p.rows_cap + 3) / 4 * 4;    // 16-byte multiple
    uint8_t* src_s = smem + static_cast<size_t>(n_ints) * 4;                                            // [rows_cap][src_pitch]
    uint8_t* temp = src_s + static_cast<size_t>(p.rows_cap) * src_pitch;                                  // [rows_cap][pitch]

    const int n = blockIdx.y;
    const int yy0 = blockIdx.x * kRows;
    const int yy1 = min(yy0 + kRows, p.out_h);
    const int y_first = t.yb[2 * yy0];
    const int rows = min(t.yb[2 * (yy1 - 1)] + t.yb[2 * (yy1 - 1) + 1] - y_first, p.rows_cap);
    const int W = p.in_w;

    // ---- phase 0: tables + source window -> shared memory
    for (int i = threadIdx.x; i < 2 * out_w + out_w * xk; i += kThreads) xb_s[i] = t.xb[i];     // xb | xc are contiguous
    for (int i = threadIdx.x; i < 2 * (yy1 - yy0); i += kThreads) yb_s[i] = t.yb[2 * yy0 + i];
    for (int i = threadIdx.x; i < (yy1 - yy0) * yk; i += kThreads) yc_s[i] = t.yc[yy0 * yk + i];
    const uintptr_t base = reinterpret_cast<uintptr_t>(src);
    const uintptr_t end = base + src_bytes;
    const int vec_per_row = src_pitch >> 4;
    for (int i = threadIdx.x; i < rows * vec_per_row; i += kThreads) {
        const int r = i / vec_per_row, v = i - r * vec_per_row;
        const uintptr_t a0 = base + ((static_cast<size_t>(n) * p.in_h + (y_first + r)) * W + p.x_lo) * 3;
        const uintptr_t al = a0 & ~static_cast<uintptr_t>(15);
        if (v == 0) off_s[r] = static_cast<int>(a0 - al);
        const uintptr_t a = al + static_cast<uintptr_t>(v) * 16;
        uint4 w = make_uint4(0, 0, 0, 0);
        if (a >= base && a + 16 <= end) {
            w = __ldg(reinterpret_cast<const uint4*>(a));
        } else {   // first / last vector of the buffer: byte-wise, never touching memory outside [src, src + src_bytes)
            uint8_t b[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) b[j] = (a + j >= base && a + j < end) ? *reinterpret_cast<const uint8_t*>(a + j) : 0;
            w = *reinterpret_cast<const uint4*>(b);
        }
        *reinterpret_cast<uint4*>(src_s + static_cast<size_t>(r) * src_pitch + v * 16) = w;
    }
    __syncthreads();

    // ---- phase 1: horizontal pass.  Thread = output column (its taps live in registers), loop over the staged rows.
    //      The 3*xk source bytes of a (row, column) are contiguous: read them as aligned words, realign with funnel shifts,
    //      pick the bytes with PRMT -- 4x fewer shared-memory transactions than byte loads.
    for (int xx = threadIdx.x; xx < out_w; xx += kThreads) {
        const int xmin = xb_s[2 * xx], xmax = xb_s[2 * xx + 1];
        int kreg[KMAX];
#pragma unroll
        for (int x = 0; x < KMAX; ++x) kreg[x] = (x < xmax) ? xc_s[xx * xk + x] : 0;     // zero taps contribute nothing
        constexpr int kWords = (3 * KMAX + 3) / 4;            // aligned words that cover 3*KMAX bytes
        const int col_byte = (xmin - p.x_lo) * 3;
        for (int r = 0; r < rows; ++r) {
            const int b0 = off_s[r] + col_byte;               // first byte inside the staged row
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(src_s + static_cast<size_t>(r) * src_pitch + (b0 & ~3));
            const uint32_t sh = static_cast<uint32_t>(b0 & 3) * 8u;
            uint32_t w[kWords + 1];
#pragma unroll
            for (int j = 0; j <= kWords; ++j) w[j] = wp[j];   // staged rows are padded: reads stay inside the row buffer
            uint32_t a[kWords];
#pragma unroll
            for (int j = 0; j < kWords; ++j) a[j] = __funnelshift_r(w[j], w[j + 1], sh);
            int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
#pragma unroll
            for (int x = 0; x < KMAX; ++x) {
                const int kv = kreg[x];
                s0 += static_cast<int>((a[(3 * x) >> 2] >> (((3 * x) & 3) * 8)) & 0xffu) * kv;
                s1 += static_cast<int>((a[(3 * x + 1) >> 2] >> (((3 * x + 1) & 3) * 8)) & 0xffu) * kv;
                s2 += static_cast<int>((a[(3 * x + 2) >> 2] >> (((3 * x + 2) & 3) * 8)) & 0xffu) * kv;
            }
            uint8_t* o = temp + r * pitch + xx * 3;
            o[0] = static_cast<uint8_t>(clip8(s0));
            o[1] = static_cast<uint8_t>(clip8(s1));
            o[2] = static_cast<uint8_t>(clip8(s2));
        }
    }
    __syncthreads();

    // ---- phase 2: vertical pass; four consecutive output bytes per thread (pitch % 4 == 0 is checked by the host side)
    uint8_t* out = dst + static_cast<size_t>(n) * p.out_h * pitch;
    const int q4 = pitch >> 2;
    for (int i = threadIdx.x; i < (yy1 - yy0) * q4; i += kThreads) {
        const int ry = i / q4, c4 = (i - ry * q4) * 4;
        const int ymin = yb_s[2 * ry] - y_first, ymax = yb_s[2 * ry + 1];
        const int32_t* k = yc_s + ry * yk;
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0, s3 = s0;
        for (int y = 0; y < ymax; ++y) {
            if (ymin + y >= rows) break;
            const uint32_t w = *reinterpret_cast<const uint32_t*>(temp + (ymin + y) * pitch + c4);
            const int kv = k[y];
            s0 += static_cast<int>(w & 0xffu) * kv;
            s1 += static_cast<int>((w >> 8) & 0xffu) * kv;
            s2 += static_cast<int>((w >> 16) & 0xffu) * kv;
            s3 += static_cast<int>(w >> 24) * kv;
        }
        *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(yy0 + ry) * pitch + c4) =
            clip8(s0) | (clip8(s1) << 8) | (clip8(s2) << 16) | (clip8(s3) << 24);
    }
}

// staging-free variant for windows that do not fit in shared memory: the horizontal pass reads the source through L1
__global__ void __launch_bounds__(kThreads)
resize_crop_u8_big_kernel(const uint8_t* __restrict__ src, const __grid_constant__ hvlm_resize_plan p,
                          const int32_t* __restrict__ table, uint8_t* __restrict__ dst) {
    extern __shared__ __align__(16) uint8_t temp[];   // [rows_cap][out_w * 3]
    const Tab t = carve(table, p);
    const int out_w = p.out_w, xk = p.xk, yk = p.yk, pitch = out_w * 3, W = p.in_w;
    const int n = blockIdx.y;
    const int yy0 = blockIdx.x * kRows;
    const int yy1 = min(yy0 + kRows, p.out_h);
    const int y_first = t.yb[2 * yy0];
    const int rows = min(t.yb[2 * (yy1 - 1)] + t.yb[2 * (yy1 - 1) + 1] - y_first, p.rows_cap);
    const uint8_t* img = src + static_cast<size_t>(n) * p.in_h * W * 3;
    for (int i = threadIdx.x; i < rows * out_w; i += kThreads) {
        const int r = i / out_w, xx = i - r * out_w;
        const int xmin = t.xb[2 * xx], xmax = t.xb[2 * xx + 1];
        const int32_t* k = t.xc + xx * xk;
        const uint8_t* q = img + (static_cast<size_t>(y_first + r) * W + xmin) * 3;
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
        for (int x = 0; x < xmax; ++x) {
            const int kv = __ldg(k + x);
            s0 += static_cast<int>(__ldg(q + 3 * x)) * kv;
            s1 += static_cast<int>(__ldg(q + 3 * x + 1)) * kv;
            s2 += static_cast<int>(__ldg(q + 3 * x + 2)) * kv;
        }
        uint8_t* o = temp + r * pitch + xx * 3;
        o[0] = static_cast<uint8_t>(clip8(s0));
        o[1] = static_cast<uint8_t>(clip8(s1));
        o[2] = static_cast<uint8_t>(clip8(s2));
    }
    __syncthreads();
    uint8_t* out = dst + static_cast<size_t>(n) * p.out_h * pitch;
    for (int i = threadIdx.x; i < (yy1 - yy0) * pitch; i += kThreads) {
        const int ry = i / pitch, col = i - ry * pitch;
        const int yy = yy0 + ry;
        const int ymin = t.yb[2 * yy] - y_first, ymax = t.yb[2 * yy + 1];
        const int32_t* k = t.yc + yy * yk;
        int s = 1 << (kPrecisionBits - 1);
        for (int y = 0; y < ymax; ++y)
            if (ymin + y < rows) s += static_cast<int>(temp[(ymin + y) * pitch + col]) * __ldg(k + y);
        out[static_cast<size_t>(yy) * pitch + col] = static_cast<uint8_t>(clip8(s));
    }
}

// expand2square (hoi_forecast/dataset/video_utils.py:13-25, the `image_aspect_ratio == 'pad'` branch of load_image :30-31):
// the frame is pasted in the middle of a square canvas of the background colour.  Thread = 4 output bytes.
__global__ void __launch_bounds__(256)
pad_square_u8_kernel(const uint8_t* __restrict__ src, int H, int W, uint8_t* __restrict__ dst, int S, int x0, int y0,
                     uint32_t bg /* r | g << 8 | b << 16 */) {
    const int n = blockIdx.z, y = blockIdx.y;
    const size_t row_bytes = static_cast<size_t>(S) * 3;
    uint8_t* out = dst + (static_cast<size_t>(n) * S + y) * row_bytes;
    const int ys = y - y0;
    const uint8_t* in = (ys >= 0 && ys < H) ? src + (static_cast<size_t>(n) * H + ys) * W * 3 : nullptr;
    const int lo = x0 * 3, hi = (x0 + W) * 3;
    for (int b4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4; b4 < static_cast<int>(row_bytes); b4 += gridDim.x * blockDim.x * 4) {
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = b4 + j;
            uint32_t v = (bg >> (8 * (b % 3))) & 0xffu;
            if (in != nullptr && b >= lo && b < hi) v = in[b - lo];
            w |= v << (8 * j);
        }
        if (b4 + 4 <= static_cast<int>(row_bytes)) {
            *reinterpret_cast<uint32_t*>(out + b4) = w;      // S*3 bytes per row; rows are 4-byte aligned when S % 4 == 0
        } else {
            for (int j = 0; b4 + j < static_cast<int>(row_bytes); ++j) out[b4 + j] = static_cast<uint8_t>(w >> (8 * j));
        }
    }
}

static inline int kmax_of(const hvlm_resize_plan& p) { return p.xk <= 8 ? 8 : (p.xk <= 12 ? 12 : 24); }
// staged row = 15 bytes of alignment slack + the window + the aligned-word over-read of the last column (3*KMAX + 8 bytes)
static inline int src_pitch_of(const hvlm_resize_plan& p) {
    return ((15 + p.x_cols * 3 + 3 * kmax_of(p) + 8 + 15) / 16) * 16;
}
static inline size_t smem_small(const hvlm_resize_plan& p) {
    size_t ints = 2 * p.out_w + static_cast<size_t>(p.out_w) * p.xk + 2 * kRows + static_cast<size_t>(kRows) * p.yk + p.rows_cap;
    ints = (ints + 3) / 4 * 4;
    return ints * 4 + static_cast<size_t>(p.rows_cap) * src_pitch_of(p) + static_cast<size_t>(p.rows_cap) * p.out_w * 3;
}

}  // namespace resize
}  // namespace hvlm

extern "C" int hvlm_resize_table_host(int in_size, int out_size, int crop0, int crop_n, int32_t* bounds_host,
                                      int32_t* coef_host) {
    using namespace hvlm::resize;
    if (in_size <= 0 || out_size <= 0 || crop0 < 0 || crop_n <= 0 || crop0 + crop_n > out_size) return HVLM_ERR_BAD_ARG;
    // Pillow: in0 / in1 are floats (the resize box), scale is computed in double from their float difference
    const float in0 = 0.0f, in1 = static_cast<float>(in_size);
    double filterscale, scale;
    filterscale = scale = static_cast<double>(in1 - in0) / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
    if (!coef_host) return ksize;
    if (!bounds_host) return HVLM_ERR_BAD_ARG;
    std::vector<double> k(static_cast<size_t>(ksize));
    for (int i = 0; i < crop_n; ++i) {
        const int xx = crop0 + i;
        const double center = in0 + (xx + 0.5) * scale;
        const double ss = 1.0 / filterscale;
        double ww = 0.0;
        int xmin = static_cast<int>(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = static_cast<int>(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            const double w = bicubic((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x)
            if (ww != 0.0) k[x] /= ww;
        for (int x = 0; x < ksize; ++x) {
            const double v = x < xmax ? k[x] : 0.0;
            coef_host[static_cast<size_t>(i) * ksize + x] =
                v < 0 ? static_cast<int>(-0.5 + v * (1 << kPrecisionBits)) : static_cast<int>(0.5 + v * (1 << kPrecisionBits));
        }
        bounds_host[2 * i] = xmin;
        bounds_host[2 * i + 1] = xmax;
    }
    return ksize;
}

extern "C" int hvlm_resize_plan_host(int in_h, int in_w, int shortest_edge, int crop, hvlm_resize_plan* plan) {
    if (!plan || in_h <= 0 || in_w <= 0 || shortest_edge <= 0 || crop <= 0 || crop > shortest_edge) return HVLM_ERR_BAD_ARG;
    hvlm_resize_plan p{};
    p.in_h = in_h;
    p.in_w = in_w;
    // transformers get_resize_output_image_size(size=int, default_to_square=False): short side -> size,
    // long side -> int(size * long / short)
    const int short_side = in_w <= in_h ? in_w : in_h, long_side = in_w <= in_h ? in_h : in_w;
    const int new_long = static_cast<int>(static_cast<double>(shortest_edge) * long_side / short_side);
    p.new_h = in_w <= in_h ? new_long : shortest_edge;
    p.new_w = in_w <= in_h ? shortest_edge : new_long;
    p.top = (p.new_h - crop) / 2;       // transformers center_crop
    p.left = (p.new_w - crop) / 2;
    p.out_h = p.out_w = crop;
    p.xk = hvlm_resize_table_host(in_w, p.new_w, p.left, crop, nullptr, nullptr);
    p.yk = hvlm_resize_table_host(in_h, p.new_h, p.top, crop, nullptr, nullptr);
    if (p.xk <= 0 || p.yk <= 0) return HVLM_ERR_BAD_ARG;
    p.table_ints = 2 * crop + crop * p.xk + 2 * crop + crop * p.yk;
    // spans: computed from the actual tables
    std::vector<int32_t> tab(static_cast<size_t>(p.table_ints));
    *plan = p;
    int rc = hvlm_resize_tables_host(plan, tab.data());
    if (rc) return rc;
    const hvlm::resize::Tab t = hvlm::resize::carve(tab.data(), p);
    p.x_lo = t.xb[0];
    p.x_cols = t.xb[2 * (crop - 1)] + t.xb[2 * (crop - 1) + 1] - p.x_lo;
    int cap = 0;
    for (int yy0 = 0; yy0 < crop; yy0 += hvlm::resize::kRows) {
        const int yy1 = (yy0 + hvlm::resize::kRows < crop ? yy0 + hvlm::resize::kRows : crop) - 1;
        const int rows = t.yb[2 * yy1] + t.yb[2 * yy1 + 1] - t.yb[2 * yy0];
        if (rows > cap) cap = rows;
    }
    p.rows_cap = cap;
    *plan = p;
    return HVLM_OK;
}

extern "C" int hvlm_resize_tables_host(const hvlm_resize_plan* plan, int32_t* table_host) {
    if (!plan || !table_host) return HVLM_ERR_BAD_ARG;
    const hvlm_resize_plan& p = *plan;
    int32_t* xb = table_host;
    int32_t* xc = xb + 2 * p.out_w;
    int32_t* yb = xc + p.out_w * p.xk;
    int32_t* yc = yb + 2 * p.out_h;
    if (hvlm_resize_table_host(p.in_w, p.new_w, p.left, p.out_w, xb, xc) != p.xk) return HVLM_ERR_BAD_ARG;
    if (hvlm_resize_table_host(p.in_h, p.new_h, p.top, p.out_h, yb, yc) != p.yk) return HVLM_ERR_BAD_ARG;
    return HVLM_OK;
}

extern "C" int hvlm_resize_crop_u8(const uint8_t* src, int N, const hvlm_resize_plan* plan_host, const int32_t* table,
                                   uint8_t* dst, void* stream) {
    using namespace hvlm;
    using namespace hvlm::resize;
    if (!src || !dst || !plan_host || !table || N <= 0) return HVLM_ERR_BAD_ARG;
    const hvlm_resize_plan p = *plan_host;
    if (p.in_h <= 0 || p.in_w <= 0 || p.out_h <= 0 || p.out_w <= 0 || p.xk <= 0 || p.yk <= 0 || p.rows_cap <= 0 ||
        p.x_lo < 0 || p.x_cols <= 0 || p.x_lo + p.x_cols > p.in_w)
        return HVLM_ERR_BAD_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const dim3 grid((p.out_h + kRows - 1) / kRows, N);
    StageTimer st(HVLM_STAGE_IM2COL, s);
    const size_t small = smem_small(p);
    if (small <= kSmemLimit && p.xk <= 24 && (p.out_w * 3) % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 3u) == 0) {
        const int kmax = kmax_of(p);
        auto kern = kmax == 8 ? resize_crop_u8_kernel<8> : (kmax == 12 ? resize_crop_u8_kernel<12> : resize_crop_u8_kernel<24>);
        if (small > 48 * 1024 &&
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemLimit)) != cudaSuccess)
            return HVLM_ERR_CUDA;
        kern<<<grid, kThreads, small, s>>>(src, static_cast<size_t>(N) * p.in_h * p.in_w * 3, p, table, dst, src_pitch_of(p));
        return check_last("resize_crop_u8");
    }
    const size_t big = static_cast<size_t>(p.rows_cap) * p.out_w * 3;
    if (big > 200 * 1024) return HVLM_ERR_BAD_SHAPE;
    if (big > 48 * 1024 && cudaFuncSetAttribute(resize_crop_u8_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                static_cast<int>(big)) != cudaSuccess)
        return HVLM_ERR_CUDA;
    resize_crop_u8_big_kernel<<<grid, kThreads, big, s>>>(src, p, table, dst);
    return check_last("resize_crop_u8_big");
}

extern "C" int hvlm_pad_square_u8(const uint8_t* src, int N, int H, int W, uint8_t* dst, int bg_r, int bg_g, int bg_b,
                                  void* stream) {
    using namespace hvlm;
    if (!src || !dst || N <= 0 || H <= 0 || W <= 0) return HVLM_ERR_BAD_ARG;
    if ((bg_r | bg_g | bg_b) < 0 || bg_r > 255 || bg_g > 255 || bg_b > 255) return HVLM_ERR_BAD_ARG;
    const int S = H > W ? H : W;
    if ((S % 4) != 0 || (reinterpret_cast<uintptr_t>(dst) & 3u) != 0) return HVLM_ERR_ALIGN;   // 32-bit row stores
    if (S > 65535 || N > 65535) return HVLM_ERR_BAD_SHAPE;
    const int x0 = W >= H ? 0 : (H - W) / 2, y0 = W > H ? (W - H) / 2 : 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    StageTimer st(HVLM_STAGE_IM2COL, s);
    const int gx = (S * 3 / 4 + 255) / 256;
    resize::pad_square_u8_kernel<<<dim3(gx, S, N), 256, 0, s>>>(src, H, W, dst, S, x0, y0,
                                                               static_cast<uint32_t>(bg_r) | (static_cast<uint32_t>(bg_g) << 8) |
                                                                   (static_cast<uint32_t>(bg_b) << 16));
    return check_last("pad_square_u8");
}
