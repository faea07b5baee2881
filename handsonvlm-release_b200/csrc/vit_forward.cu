// CLIP ViT-L/14 tower forward: host-side orchestration of the sm_100a kernels.
//
// Replaces CLIPVisionTower.forward (llava/model/multimodal_encoder/clip_encoder.py:39-51), i.e. HF
// CLIPVisionModel(..., output_hidden_states=True).hidden_states[select_layer].  Only the layers that are needed
// run (23 of 24 for select_layer=-2); no hidden-state list is kept; the fp32 residual stream is the single
// [n,257,1024] output buffer and every GEMM epilogue reads/writes it in place.
//
// Per layer (5 launches; default):  QKV GEMM (LN1 folded in; column-block-major q|k|v [48][M][64] via 3-D TMA stores) -> attention
//                          -> out_proj GEMM (+residual; emits bf16 rows + row statistics) -> fc1 GEMM (LN2 folded in, +quick-GELU)
//                          -> fc2 GEMM (+residual; emits bf16 rows + row statistics)
// HVLM_LN_FOLD=0 / hvlm_vit_set_ln_fold(0) (7 launches):  LN1 -> QKV GEMM -> attention -> out_proj GEMM (+residual) -> LN2 -> fc1 GEMM -> fc2 GEMM
#include <cstdlib>

#include "hvlm_internal.cuh"

namespace hvlm {
int launch_layernorm(const float* x, const float* g, const float* b, void* out, int rows, int out_dtype, float eps,
                     cudaStream_t s, int reverse, const float* add_rows = nullptr, int add_period = 1, void* xb_out = nullptr,
                     float* stats_out = nullptr, float* shift_out = nullptr);
int launch_im2col(const void* pixels, int pix_dtype, int n_frames, void* A, const float* cls, const float* pos,
                  float* x0, cudaStream_t s);
int launch_im2col_u8(const uint8_t* frames, const float* mean, const float* stdv, int n_frames, void* A, const float* cls,
                     const float* pos, float* x0, cudaStream_t s);
int launch_attention(const void* qkv, void* out, int n_frames, cudaStream_t s);

static inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

struct VitWorkspace {
    uint64_t a_patch, y, qkv, attn, f1, ln_count, stats, shift, total;
};

static VitWorkspace vit_workspace(int n_frames) {
    const uint64_t F = static_cast<uint64_t>(n_frames);
    const uint64_t M = F * HVLM_VIT_TOKENS;
    VitWorkspace w;
    uint64_t off = 0;
    auto take = [&](uint64_t bytes) {
        uint64_t o = off;
        off = align_up(off + bytes, 1024);
        return o;
    };
    w.a_patch = take(F * 256 * HVLM_VIT_PATCH_KPAD * 2);
    w.y = take(M * 1024 * 2);
    w.qkv = take(M * 3072 * 2);
    w.attn = take(M * 1024 * 2);
    w.f1 = take(M * 4096 * 2);
    w.ln_count = take(((M + 127) / 128) * 4);
    w.stats = take(M * 16 * 4);      // folded LayerNorm: (sum, sum of squares) per row and 128-column block
    w.shift = take(M * 4);           // folded LayerNorm: the row's running mean (the centre of its bf16 copy / statistics)
    w.total = off;
    return w;
}

}  // namespace hvlm

extern "C" int hvlm_vit_l14_layout(int n_layers, hvlm_vit_layout* L) {
    using namespace hvlm;
    if (!L || n_layers < 0 || n_layers > HVLM_VIT_MAX_LAYERS) return HVLM_ERR_BAD_ARG;
    uint64_t off = 0;
    auto take = [&](uint64_t bytes) {
        uint64_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    *L = hvlm_vit_layout{};
    L->patch_w = take(1024ull * HVLM_VIT_PATCH_KPAD * 2);
    L->cls = take(1024 * 4);
    L->pos = take(257ull * 1024 * 4);
    L->pre_ln_g = take(1024 * 4);
    L->pre_ln_b = take(1024 * 4);
    for (int l = 0; l < n_layers; ++l) {
        auto& y = L->layer[l];
        y.ln1_g = take(1024 * 4);
        y.ln1_b = take(1024 * 4);
        y.w_qkv = take(3072ull * 1024 * 2);
        y.b_qkv = take(3072 * 4);
        y.w_o = take(1024ull * 1024 * 2);
        y.b_o = take(1024 * 4);
        y.ln2_g = take(1024 * 4);
        y.ln2_b = take(1024 * 4);
        y.w_fc1 = take(4096ull * 1024 * 2);
        y.b_fc1 = take(4096 * 4);
        y.w_fc2 = take(1024ull * 4096 * 2);
        y.b_fc2 = take(1024 * 4);
        auto& f = L->fold[l];       // ABI 3: the folded-LayerNorm operands of the layer (per layer, so that a blob packed
        f.w_qkv_f = take(3072ull * 1024 * 2);   // for n layers still runs any prefix of them)
        f.c_qkv = take(3072 * 4);
        f.b_qkv_f = take(3072 * 4);
        f.w_fc1_f = take(4096ull * 1024 * 2);
        f.c_fc1 = take(4096 * 4);
        f.b_fc1_f = take(4096 * 4);
    }
    L->total_bytes = off;
    L->n_layers = n_layers;
    return HVLM_OK;
}

namespace hvlm {
static int g_ln_fold = -1;    // -1: not decided yet (environment); 0 off, 1 both LayerNorms, 2 LN1 only, 3 LN2 only
static int ln_fold_setting() {
    if (g_ln_fold < 0) {
        const char* e = getenv("HVLM_LN_FOLD");
        const char* one = getenv("HVLM_GEMM_1CTA");     // the fold lives in the 2-CTA kernel only
        int v = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 1;
        g_ln_fold = (one && one[0] == '1') ? 0 : v;
    }
    return g_ln_fold;
}
}  // namespace hvlm

extern "C" int hvlm_vit_set_ln_fold(int on) {
    const int prev = hvlm::ln_fold_setting();
    if (on >= 0) {
        const char* one = getenv("HVLM_GEMM_1CTA");
        hvlm::g_ln_fold = (one && one[0] == '1') ? 0 : (on <= 3 ? on : 1);
    }
    return prev;
}

extern "C" size_t hvlm_vit_l14_workspace_bytes(int n_frames) {
    if (n_frames <= 0) return 0;
    return static_cast<size_t>(hvlm::vit_workspace(n_frames).total);
}

extern "C" int hvlm_vit_qkv_gemm(const void* A, const void* w_qkv, const float* b_qkv, void* qkv_hm, int n_frames,
                                 void* stream) {
    using namespace hvlm;
    if (!A || !w_qkv || !qkv_hm || n_frames <= 0) return HVLM_ERR_BAD_ARG;
    if (!aligned16(qkv_hm) || (b_qkv && !aligned16(b_qkv))) return HVLM_ERR_ALIGN;
    EpiArgs ep;
    ep.bias = b_qkv;
    ep.out = qkv_hm;
    StageTimer st(HVLM_STAGE_QKV_GEMM, static_cast<cudaStream_t>(stream));
    return launch_gemm(EPI_QKV_HM, A, w_qkv, n_frames * HVLM_VIT_TOKENS, 3072, 1024, ep, static_cast<cudaStream_t>(stream));
}

static int vit_l14_fwd_impl(const void* weight_blob, int n_layers_run, const void* pixels, int pix_dtype,
                            const float* u8_mean, const float* u8_std, int n_frames, float* hidden, void* workspace,
                            size_t workspace_bytes, void* stream, bool open_last_mlp = false) {
    using namespace hvlm;
    if (!weight_blob || !pixels || !hidden || !workspace) return HVLM_ERR_BAD_ARG;
    if (n_frames <= 0 || n_layers_run < 0 || n_layers_run > HVLM_VIT_MAX_LAYERS) return HVLM_ERR_BAD_ARG;
    if (!aligned16(weight_blob) || !aligned16(pixels) || !aligned16(hidden) ||
        (reinterpret_cast<uintptr_t>(workspace) & 1023u))
        return HVLM_ERR_ALIGN;
    const VitWorkspace ws = vit_workspace(n_frames);
    if (workspace_bytes < ws.total) return HVLM_ERR_WORKSPACE;
    const PdlScope pdl(n_frames <= 32);      // small batches: overlap each kernel's prologue with its predecessor's tail
    hvlm_vit_layout L;
    int rc = hvlm_vit_l14_layout(n_layers_run, &L);
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint8_t* wb = static_cast<const uint8_t*>(weight_blob);
    uint8_t* w8 = static_cast<uint8_t*>(workspace);
    auto f32 = [&](uint64_t off) { return reinterpret_cast<const float*>(wb + off); };
    const int M = n_frames * HVLM_VIT_TOKENS;

    // Experimental (HVLM_LN_FUSION=1): LayerNorm fused into the residual GEMMs (extra warps normalise each 128-row block
    // as soon as its column tiles are reduced).  Correct, but measured SLOWER on B200 (15.75 vs 15.20 ms per 100 frames):
    // the in-kernel LN warps are L2-latency-bound (2 rows in flight per warp) while the stand-alone LN kernel already
    // streams at ~5 TB/s, so the separate kernels stay the default.
    static const bool fuse_ln = []() {
        const char* e = getenv("HVLM_LN_FUSION");
        return e && e[0] == '1';
    }();
    // which of the two LayerNorms of a layer are folded: LN1 (fc2 of the previous layer / pre_layrnorm produce, QKV consumes),
    // LN2 (out_proj produces, fc1 consumes)
    const int fold_set = (!fuse_ln && n_layers_run > 0) ? ln_fold_setting() : 0;
    const bool fold1 = fold_set == 1 || fold_set == 2, fold2 = fold_set == 1 || fold_set == 3;
    const bool fold = fold1 || fold2;
    // tile traversal directions of the four layer GEMMs (bit 0 QKV, 1 out_proj, 2 fc1, 3 fc2: 1 = last-to-first); A/B switch
    static const int rev_mask = []() {
        const char* e = getenv("HVLM_GEMM_REV_MASK");
        return e ? atoi(e) : 8;
    }();
    float* stats = reinterpret_cast<float*>(w8 + ws.stats);
    float* shift = reinterpret_cast<float*>(w8 + ws.shift);
    int32_t* ln_count = reinterpret_cast<int32_t*>(w8 + ws.ln_count);
    if (fuse_ln && cudaMemsetAsync(ln_count, 0, static_cast<size_t>((M + 127) / 128) * 4, s) != cudaSuccess) return HVLM_ERR_CUDA;

    // embeddings: im2col (+ CLS rows) -> patch GEMM (+ position embedding) -> pre_layrnorm (in place)
    {
        StageTimer st(HVLM_STAGE_IM2COL, s);
        if (u8_mean)
            rc = launch_im2col_u8(static_cast<const uint8_t*>(pixels), u8_mean, u8_std, n_frames, w8 + ws.a_patch, f32(L.cls),
                                  f32(L.pos), hidden, s);
        else
            rc = launch_im2col(pixels, pix_dtype, n_frames, w8 + ws.a_patch, f32(L.cls), f32(L.pos), hidden, s);
    }
    if (rc) return rc;
    {
        EpiArgs ep;
        ep.out = hidden;      // bare convolution output; the position embedding is added by the pre-LayerNorm kernel
        StageTimer st(HVLM_STAGE_PATCH_GEMM, s);
        rc = launch_gemm(EPI_PATCH, w8 + ws.a_patch, wb + L.patch_w, n_frames * 256, 1024, HVLM_VIT_PATCH_KPAD, ep, s);
        if (rc) return rc;
    }
    {
        StageTimer st(HVLM_STAGE_LAYERNORM, s);
        // + position embedding (tokens 1..256; the CLS row got pos[0] from im2col), then pre_layrnorm, in place
        // (folded LayerNorms: it also writes the bf16 rows + row statistics the first QKV GEMM consumes)
        rc = launch_layernorm(hidden, f32(L.pre_ln_g), f32(L.pre_ln_b), hidden, M, HVLM_F32, 1e-5f, s, 1, f32(L.pos),
                              HVLM_VIT_TOKENS, fold ? w8 + ws.y : nullptr, fold ? stats : nullptr, fold ? shift : nullptr);   // (shift: any fold)
    }
    if (rc) return rc;

    for (int l = 0; l < n_layers_run; ++l) {
        const auto& y = L.layer[l];
        const auto& yf = L.fold[l];
        if (!fold1 && (!fuse_ln || l == 0)) {   // with fusion, LN1 of layer l > 0 was produced by fc2 of layer l-1
            StageTimer st(HVLM_STAGE_LAYERNORM, s);
            rc = launch_layernorm(hidden, f32(y.ln1_g), f32(y.ln1_b), w8 + ws.y, M, HVLM_BF16, 1e-5f, s, 0);
            if (rc) return rc;
        }
        {
            EpiArgs ep;
            ep.bias = f32(fold1 ? yf.b_qkv_f : y.b_qkv);
            ep.out = w8 + ws.qkv;
            ep.reverse = rev_mask & 1;
            if (fold1) {      // LN1 folded in: A = bf16 residual rows, B = gamma-scaled weights
                ep.ln_c = f32(yf.c_qkv);
                ep.ln_stats = stats;
                ep.shift_io = shift;
            }
            StageTimer st(HVLM_STAGE_QKV_GEMM, s);
            rc = launch_gemm(EPI_QKV_HM, w8 + ws.y, wb + (fold1 ? yf.w_qkv_f : y.w_qkv), M, 3072, 1024, ep, s);
            if (rc) return rc;
        }
        {
            StageTimer st(HVLM_STAGE_ATTENTION, s);
            rc = launch_attention(w8 + ws.qkv, w8 + ws.attn, n_frames, s);
        }
        if (rc) return rc;
        {
            EpiArgs ep;
            ep.bias = f32(y.b_o);
            ep.resid = hidden;
            ep.out = hidden;
            ep.reverse = (rev_mask >> 1) & 1;
            if (fuse_ln) {   // LN2 fused: y = LN2(hidden) is produced as row blocks complete
                ep.ln_gamma = f32(y.ln2_g);
                ep.ln_beta = f32(y.ln2_b);
                ep.ln_out = w8 + ws.y;
                ep.ln_count = ln_count;
            }
            if (fold2) {      // feeds the folded LN2 of fc1
                ep.xb_out = w8 + ws.y;
                ep.stats_out = stats;
                ep.shift_in = shift;
            }
            StageTimer st(HVLM_STAGE_OUTPROJ_GEMM, s);
            rc = launch_gemm(EPI_RESID_F32, w8 + ws.attn, wb + y.w_o, M, 1024, 1024, ep, s);
            if (rc) return rc;
        }
        if (!fuse_ln && !fold2) {
            StageTimer st(HVLM_STAGE_LAYERNORM, s);
            rc = launch_layernorm(hidden, f32(y.ln2_g), f32(y.ln2_b), w8 + ws.y, M, HVLM_BF16, 1e-5f, s, 1);
            if (rc) return rc;
        }
        {
            EpiArgs ep;
            ep.bias = f32(fold2 ? yf.b_fc1_f : y.b_fc1);
            ep.out = w8 + ws.f1;
            ep.reverse = (rev_mask >> 2) & 1;
            if (fold2) {
                ep.ln_c = f32(yf.c_fc1);
                ep.ln_stats = stats;
                ep.shift_io = shift;
            }
            StageTimer st(HVLM_STAGE_FC1_GEMM, s);
            rc = launch_gemm(EPI_GELU_BF16, w8 + ws.y, wb + (fold2 ? yf.w_fc1_f : y.w_fc1), M, 4096, 1024, ep, s);
            if (rc) return rc;
        }
        if (open_last_mlp && l + 1 == n_layers_run) break;   // the caller applies fc2 after pooling (it is linear)
        {
            EpiArgs ep;
            ep.bias = f32(y.b_fc2);
            ep.resid = hidden;
            ep.out = hidden;
            ep.reverse = (rev_mask >> 3) & 1;   // default 1: fc1 wrote f1 first-to-last, its tail is what L2 still holds
            if (fuse_ln && l + 1 < n_layers_run) {   // LN1 of the next layer fused into this GEMM
                ep.ln_gamma = f32(L.layer[l + 1].ln1_g);
                ep.ln_beta = f32(L.layer[l + 1].ln1_b);
                ep.ln_out = w8 + ws.y;
                ep.ln_count = ln_count;
            }
            if (fold1 && l + 1 < n_layers_run) {      // feeds the folded LN1 of the next layer (nobody reads them after the last)
                ep.xb_out = w8 + ws.y;
                ep.stats_out = stats;
                ep.shift_in = shift;
            }
            StageTimer st(HVLM_STAGE_FC2_GEMM, s);
            rc = launch_gemm(EPI_RESID_F32, w8 + ws.f1, wb + y.w_fc2, M, 1024, 4096, ep, s);
            if (rc) return rc;
        }
    }
    return HVLM_OK;
}

extern "C" int hvlm_vit_l14_fwd(const void* weight_blob, int n_layers_run, const void* pixels, int pix_dtype,
                                int n_frames, float* hidden, void* workspace, size_t workspace_bytes, void* stream) {
    return vit_l14_fwd_impl(weight_blob, n_layers_run, pixels, pix_dtype, nullptr, nullptr, n_frames, hidden, workspace,
                            workspace_bytes, stream);
}

extern "C" int hvlm_vit_l14_fwd_u8(const void* weight_blob, int n_layers_run, const uint8_t* frames_nhwc,
                                   const float* mean_host, const float* std_host, int n_frames, float* hidden,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    if (!mean_host || !std_host) return HVLM_ERR_BAD_ARG;
    for (int c = 0; c < 3; ++c)
        if (!(std_host[c] > 0.f)) return HVLM_ERR_BAD_ARG;
    return vit_l14_fwd_impl(weight_blob, n_layers_run, frames_nhwc, HVLM_F32, mean_host, std_host, n_frames, hidden,
                            workspace, workspace_bytes, stream);
}

// Same as hvlm_vit_l14_fwd / _fwd_u8 but the LAST layer's second MLP matmul is left to the caller: on return `hidden` is the
// residual stream after the last attention block and workspace + *f1_offset holds gelu(fc1(LN2(hidden))) as bf16
// [n_frames*257, 4096].  Token pooling is a fixed linear map over tokens, so the caller can pool both tensors first and
// apply fc2 (+ bias + residual) to the 356 pooled rows per clip instead of the 25 700 token rows:
//   pool(hidden + f1 W2^T + b2) = pool(hidden) + pool(f1) W2^T + b2.
extern "C" int hvlm_vit_l14_fwd_open_mlp(const void* weight_blob, int n_layers_run, const void* pixels, int pix_dtype,
                                         const float* mean_host, const float* std_host, int n_frames, float* hidden,
                                         void* workspace, size_t workspace_bytes, uint64_t* f1_offset, void* stream) {
    if (!f1_offset || n_layers_run < 1) return HVLM_ERR_BAD_ARG;
    const bool u8 = pix_dtype < 0;
    if (u8) {
        if (!mean_host || !std_host) return HVLM_ERR_BAD_ARG;
        for (int c = 0; c < 3; ++c)
            if (!(std_host[c] > 0.f)) return HVLM_ERR_BAD_ARG;
    }
    *f1_offset = hvlm::vit_workspace(n_frames > 0 ? n_frames : 1).f1;
    return vit_l14_fwd_impl(weight_blob, n_layers_run, pixels, u8 ? HVLM_F32 : pix_dtype, u8 ? mean_host : nullptr,
                            u8 ? std_host : nullptr, n_frames, hidden, workspace, workspace_bytes, stream, true);
}
