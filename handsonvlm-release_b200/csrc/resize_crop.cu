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
// tables + a numpy restatement to PIL itself; tests/test_gpu_parity.py pins the kernel).
//
// CTA = (block of kRows output rows, frame).  Phase 1: horizontal pass of the source rows this block needs into shared
// memory (uint8 [rows][out_w*3]); phase 2: vertical pass from shared memory, coalesced byte-plane stores.
#include <cmath>
#include <vector>

#include "hvlm_internal.cuh"

namespace hvlm {
namespace resize {

constexpr int kRows = 8;
constexpr int kThreads = 256;
constexpr int kPrecisionBits = 32 - 8 - 2;

static double bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

__device__ __forceinline__ uint8_t clip8(int acc) {
    const int v = acc >> kPrecisionBits;     // arithmetic shift == Pillow's table index
    return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__global__ void __launch_bounds__(kThreads)
resize_crop_u8_kernel(const uint8_t* __restrict__ src, int H, int W, uint8_t* __restrict__ dst, int out_h, int out_w,
                      const int32_t* __restrict__ xb, const int32_t* __restrict__ xc, int xk,
                      const int32_t* __restrict__ yb, const int32_t* __restrict__ yc, int yk, int max_rows) {
    extern __shared__ uint8_t temp[];   // [max_rows][out_w * 3]
    const int n = blockIdx.y;
    const int yy0 = blockIdx.x * kRows;
    const int yy1 = min(yy0 + kRows, out_h);
    const int y_first = yb[2 * yy0];
    int y_end = yb[2 * (yy1 - 1)] + yb[2 * (yy1 - 1) + 1];
    if (y_end - y_first > max_rows) y_end = y_first + max_rows;   // cannot happen for tables of hvlm_resize_table_host
    const int rows = y_end - y_first;
    const int pitch = out_w * 3;
    const uint8_t* img = src + static_cast<size_t>(n) * H * W * 3;

    // ---- phase 1: horizontal pass (one thread per temp pixel, 3 channels)
    for (int i = threadIdx.x; i < rows * out_w; i += kThreads) {
        const int r = i / out_w, xx = i - r * out_w;
        const int xmin = xb[2 * xx], xmax = xb[2 * xx + 1];
        const int32_t* k = xc + xx * xk;
        const uint8_t* p = img + (static_cast<size_t>(y_first + r) * W + xmin) * 3;
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
        for (int x = 0; x < xmax; ++x) {
            const int kv = __ldg(k + x);
            s0 += static_cast<int>(__ldg(p + 3 * x)) * kv;
            s1 += static_cast<int>(__ldg(p + 3 * x + 1)) * kv;
            s2 += static_cast<int>(__ldg(p + 3 * x + 2)) * kv;
        }
        uint8_t* t = temp + r * pitch + xx * 3;
        t[0] = clip8(s0);
        t[1] = clip8(s1);
        t[2] = clip8(s2);
    }
    __syncthreads();

    // ---- phase 2: vertical pass (one thread per output byte: adjacent threads write adjacent bytes)
    uint8_t* out = dst + static_cast<size_t>(n) * out_h * pitch;
    for (int i = threadIdx.x; i < (yy1 - yy0) * pitch; i += kThreads) {
        const int ry = i / pitch, col = i - ry * pitch;
        const int yy = yy0 + ry;
        const int ymin = yb[2 * yy] - y_first, ymax = yb[2 * yy + 1];
        const int32_t* k = yc + yy * yk;
        int s = 1 << (kPrecisionBits - 1);
        for (int y = 0; y < ymax; ++y)
            if (ymin + y < rows) s += static_cast<int>(temp[(ymin + y) * pitch + col]) * __ldg(k + y);
        out[static_cast<size_t>(yy) * pitch + col] = clip8(s);
    }
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

extern "C" int hvlm_resize_crop_u8(const uint8_t* src, int N, int H, int W, uint8_t* dst, int out_h, int out_w,
                                   const int32_t* xbounds, const int32_t* xcoef, int xk, const int32_t* ybounds,
                                   const int32_t* ycoef, int yk, void* stream) {
    using namespace hvlm;
    using namespace hvlm::resize;
    if (!src || !dst || !xbounds || !xcoef || !ybounds || !ycoef) return HVLM_ERR_BAD_ARG;
    if (N <= 0 || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0 || xk <= 0 || yk <= 0) return HVLM_ERR_BAD_ARG;
    // rows of the intermediate image one block of kRows output rows can need: consecutive output rows start at most
    // ceil(scale) + 1 source rows apart and the last one spans yk taps (scale <= H / out_h because out_h is a crop)
    const int step = (H + out_h - 1) / out_h + 1;
    const int max_rows = (kRows - 1) * step + yk;
    const size_t smem = static_cast<size_t>(max_rows) * out_w * 3;
    if (smem > 200 * 1024) return HVLM_ERR_BAD_SHAPE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(resize_crop_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) !=
            cudaSuccess)
        return HVLM_ERR_CUDA;
    StageTimer st(HVLM_STAGE_IM2COL, s);
    resize_crop_u8_kernel<<<dim3((out_h + kRows - 1) / kRows, N), kThreads, smem, s>>>(src, H, W, dst, out_h, out_w, xbounds,
                                                                                      xcoef, xk, ybounds, ycoef, yk, max_rows);
    return check_last("resize_crop_u8");
}
