/*
 * hvlm_b200.h -- C ABI of the B200-native (sm_100a) HandsOnVLM visual-token path.
 *
 * The reference (Kami-code/HandsOnVLM-release) has NO native / FFI boundary: the path is five Python
 * methods.  This header is the boundary a maintainer would bind instead; every entry
 * point names the reference code it replaces (paths relative to the reference checkout).  The Python
 * drop-ins in handsonvlm-release_b200/ call exactly these symbols through ctypes, wrapped as
 * torch.library custom ops (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - plain pointers + sizes; every pointer is a DEVICE pointer unless the name ends in _host
 *   - the caller owns all buffers (inputs, outputs, workspace); nothing is retained after return
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises the host
 *   - return value: 0 (HVLM_OK) or a negative hvlm_status; never throws, never aborts
 *   - there is NO CPU fallback: without an sm_100 device every compute entry returns HVLM_ERR_CUDA
 *   - "bf16 GEMM regime": bf16 operands, fp32 accumulation in TMEM (tcgen05.mma kind::f16)
 */
#ifndef HVLM_B200_H
#define HVLM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HVLM_ABI_VERSION 3

#if defined(__GNUC__)
#define HVLM_API __attribute__((visibility("default")))
#else
#define HVLM_API
#endif

typedef enum {
    HVLM_OK = 0,
    HVLM_ERR_BAD_ARG = -1,      /* null pointer / negative size / unknown enum */
    HVLM_ERR_BAD_SHAPE = -2,    /* shape the kernels do not support (e.g. tokens/frame != 256) */
    HVLM_ERR_BAD_DTYPE = -3,
    HVLM_ERR_ALIGN = -4,        /* pointer not 16-byte aligned */
    HVLM_ERR_CUDA = -5,         /* launch / driver failure, or no sm_100 device */
    HVLM_ERR_WORKSPACE = -6,    /* workspace too small */
    HVLM_ERR_UNSUPPORTED = -7
} hvlm_status;

typedef enum { HVLM_F32 = 0, HVLM_BF16 = 1, HVLM_F16 = 2 } hvlm_dtype;

/* video_arch / video_compress_mode  (lita/model/lita_arch.py:41-75, hoi_forecast/model/visual_to_tokens.py:230-272) */
typedef enum {
    HVLM_POOL_TEMPORAL_SPATIAL_POOL = 0, /* [t fast means] ++ [4 frames x 8x8 2x2-avg]  -> t+256 tokens */
    HVLM_POOL_SPATIAL_POOL = 1,          /* slow tokens only                             -> 256 tokens   */
    HVLM_POOL_TEMPORAL = 2,              /* mean over the 256 tokens of each frame       -> t tokens     */
    HVLM_POOL_SPATIAL = 3,               /* mean over frames                             -> 256 tokens   */
    HVLM_POOL_TEMPORAL_SPATIAL = 4       /* temporal ++ spatial                          -> t+256 tokens */
} hvlm_pool_mode;

/* which prepare_inputs_labels_for_multimodal is being replaced */
typedef enum {
    HVLM_SPLICE_LLAVA = 0,      /* llava/model/llava_arch.py:110-234  (mask NOT position-spliced)          */
    HVLM_SPLICE_HANDSONVLM = 1  /* handsonvlm/model/language_model/handsonvlm.py:212-451 (mask spliced,    */
                                /* hand positional embeddings added at <hand_traj> rows of the tail)       */
} hvlm_splice_variant;

/* GEMM epilogues (C = A * B^T, A [M,K] bf16 row-major, B [N,K] bf16 row-major) */
typedef enum {
    HVLM_EPI_BIAS = 0,          /* out = acc + bias[n]                                   */
    HVLM_EPI_BIAS_QUICKGELU = 1,/* out = g(acc + bias[n]), g(x) = x * sigmoid(1.702 x)   */
    HVLM_EPI_BIAS_RESIDUAL = 2  /* out = resid[m,n] + acc + bias[n]  (out may alias resid) */
} hvlm_epilogue;

HVLM_API const char* hvlm_strerror(int status);
HVLM_API int hvlm_abi_version(void);
/* 0 if `device` is an sm_100 GPU this library can run on, else HVLM_ERR_CUDA. */
HVLM_API int hvlm_device_check(int device);

/* ------------------------------------------------------------------------------------------------
 * tcgen05 GEMM  -- replaces every nn.Linear on the path:
 *   mm_projector                           llava/model/llava_arch.py:33,92 ; visual_to_tokens.py:279
 *   CLIP q/k/v/out_proj, fc1, fc2          transformers CLIPEncoderLayer (clip_encoder.py:48)
 *   projector wgrad (dW = dY^T X)          autograd of the above (training-shaped variant)
 * bias may be NULL (treated as 0).  N % 128 == 0, K % 8 == 0 (K tail < 64 is zero-filled by TMA).
 * out_dtype: HVLM_BF16 or HVLM_F32.  resid (fp32 [M,N]) only for HVLM_EPI_BIAS_RESIDUAL.
 * ---------------------------------------------------------------------------------------------- */
HVLM_API int hvlm_gemm_bf16(const void* A, const void* B, const float* bias, const float* resid, void* out, int M, int N,
                   int K, int epilogue, int out_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CLIP ViT-L/14 tower -- replaces CLIPVisionTower.forward
 *   llava/model/multimodal_encoder/clip_encoder.py:39-51 (+ feature_select :29-37)
 * Fixed architecture: hidden 1024, 16 heads x 64, MLP 4096, 224x224 / patch 14 -> 257 tokens,
 * quick-GELU, LayerNorm eps 1e-5.  bf16 GEMM regime with an fp32 residual stream, fp32 LayerNorm
 * statistics and fp32 softmax.
 *
 * Weights live in ONE device blob whose layout is defined here (hvlm_vit_layout); the host side fills
 * it once from the HF-named state dict (q_proj weight/bias pre-scaled by 64^-1/2, exact in bf16).
 * ---------------------------------------------------------------------------------------------- */
#define HVLM_VIT_HIDDEN 1024
#define HVLM_VIT_HEADS 16
#define HVLM_VIT_HEAD_DIM 64
#define HVLM_VIT_MLP 4096
#define HVLM_VIT_PATCH 14
#define HVLM_VIT_IMAGE 224
#define HVLM_VIT_TOKENS 257      /* CLS + 256 patches */
#define HVLM_VIT_PATCHES 256
#define HVLM_VIT_PATCH_K 588     /* 3*14*14 */
#define HVLM_VIT_PATCH_KPAD 640  /* padded to a multiple of 64 with zeros */
#define HVLM_VIT_MAX_LAYERS 24

typedef struct {
    /* byte offsets into the blob */
    uint64_t patch_w;   /* bf16 [1024, 640]   conv weight reshaped [E, c*196+i*14+j], zero padded */
    uint64_t cls;       /* f32  [1024]        class_embedding                                   */
    uint64_t pos;       /* f32  [257, 1024]   position_embedding.weight                          */
    uint64_t pre_ln_g;  /* f32  [1024]        pre_layrnorm.weight   (sic, HF spelling)           */
    uint64_t pre_ln_b;  /* f32  [1024] */
    struct {
        uint64_t ln1_g, ln1_b; /* f32 [1024]                                                      */
        uint64_t w_qkv;        /* bf16 [3072,1024] rows: q (pre-scaled 1/8) | k | v                */
        uint64_t b_qkv;        /* f32 [3072]                                                      */
        uint64_t w_o, b_o;     /* bf16 [1024,1024], f32 [1024]                                    */
        uint64_t ln2_g, ln2_b; /* f32 [1024]                                                      */
        uint64_t w_fc1, b_fc1; /* bf16 [4096,1024], f32 [4096]                                    */
        uint64_t w_fc2, b_fc2; /* bf16 [1024,4096], f32 [1024]                                    */
    } layer[HVLM_VIT_MAX_LAYERS];
    uint64_t total_bytes;
    int32_t n_layers;
    int32_t _pad;
    /* ABI 3: LayerNorm folded into the GEMM that consumes it (see hvlm_gemm_ln_fold_bf16).  For LN(x) W^T + b with
     * LN(x) = (x - mean) * rstd * gamma + beta the host also packs, per layer and for both LayerNorms:
     *   w_*_f  bf16 [N,1024]  round(gamma[k] * W[n,k])                (from the fp32 master weights: one rounding)
     *   c_*    f32  [N]       sum_k float(w_*_f[n,k])                 (of the ROUNDED folded weights)
     *   b_*_f  f32  [N]       b[n] + sum_k beta[k] * W[n,k]
     * (qkv: W, b are the q|k|v rows with q pre-scaled like w_qkv / b_qkv.) */
    struct {
        uint64_t w_qkv_f, c_qkv, b_qkv_f;
        uint64_t w_fc1_f, c_fc1, b_fc1_f;
    } fold[HVLM_VIT_MAX_LAYERS];
} hvlm_vit_layout;

HVLM_API int hvlm_vit_l14_layout(int n_layers, hvlm_vit_layout* out_host);
/* workspace needed for `n_frames` frames (bytes). */
HVLM_API size_t hvlm_vit_l14_workspace_bytes(int n_frames);
/*
 * pixels  [n_frames,3,224,224] NCHW, pix_dtype f32|bf16|f16
 * hidden  f32 [n_frames,257,1024]: residual stream after `n_layers_run` encoder layers == HF
 *         hidden_states[n_layers_run]; select_layer=-2 of the 24-layer tower => n_layers_run = 23
 *         (layer 24 and post_layernorm are never computed).  feature_select 'patch' = rows 1..256.
 */
HVLM_API int hvlm_vit_l14_fwd(const void* weight_blob, int n_layers_run, const void* pixels, int pix_dtype, int n_frames,
                     float* hidden, void* workspace, size_t workspace_bytes, void* stream);
/* Same tower fed with raw uint8 frames [n_frames,224,224,3] (NHWC, as decoded); the CLIPImageProcessor rescale +
 * normalise ((u8/255 - mean[c]) / std[c], hoi_forecast/dataset/video_utils.py + transformers CLIPImageProcessor) is fused
 * into the patch extraction, so a clip crosses PCIe as 15 MB instead of 60 MB fp32.  mean/std: 3 floats each, HOST. */
HVLM_API int hvlm_vit_l14_fwd_u8(const void* weight_blob, int n_layers_run, const uint8_t* frames_nhwc,
                                 const float* mean_host, const float* std_host, int n_frames, float* hidden,
                                 void* workspace, size_t workspace_bytes, void* stream);
/* The same forward with the LAST layer's second MLP matmul left to the caller (pix_dtype < 0: uint8 NHWC frames with
 * mean_host / std_host, else an hvlm_dtype and the two pointers are ignored).  On return `hidden` is the residual stream
 * after the last attention block and workspace + *f1_offset holds gelu(fc1(LN2(hidden))) as bf16 [n_frames*257, 4096].
 * The LITA / HandsOnVLM token pooling (lita_arch.py:54-70, visual_to_tokens.py:252-271) is a fixed linear map over tokens,
 * so  pool(hidden + f1 W2^T + b2) = pool(hidden) + pool(f1) W2^T + b2 : fc2 runs on the pooled rows only. */
HVLM_API int hvlm_vit_l14_fwd_open_mlp(const void* weight_blob, int n_layers_run, const void* pixels, int pix_dtype,
                              const float* mean_host, const float* std_host, int n_frames, float* hidden,
                              void* workspace, size_t workspace_bytes, uint64_t* f1_offset, void* stream);
/* hidden f32 [n,257,1024] -> feats [n,256,1024] (drop CLS) cast to out_dtype (clip_encoder.py:31-32,49). */
HVLM_API int hvlm_feature_select(const float* hidden, void* feats, int n_frames, int out_dtype, int keep_cls, void* stream);

/* The tower runs with its 46 per-layer LayerNorm launches folded into the GEMMs around them (default = 1; the environment
 * variable HVLM_LN_FOLD=0 or this call with on = 0 restores the stand-alone LayerNorm kernels; 2 / 3 fold only LayerNorm 1
 * (fc2 -> QKV) / only LayerNorm 2 (out_proj -> fc1), for A/B runs; on < 0 only queries).
 * Process-wide switch, returns the previous setting.  Both settings produce the same tower within bf16-operand rounding
 * (same operand precision: the GEMM reads bf16(x) and bf16(gamma*W) instead of bf16(LN(x)) and bf16(W)). */
HVLM_API int hvlm_vit_set_ln_fold(int on);

/* building blocks of the tower, exported for per-stage parity tests */
HVLM_API int hvlm_layernorm_1024(const float* x, const float* gamma, const float* beta, void* out, int rows, int out_dtype,
                        float eps, void* stream);
/* The same LayerNorm with f32 output `out` [rows,1024] (may alias x) that also emits what a folded GEMM consumes next.
 * A LayerNorm does not see a constant added to its row, so everything the fold hands over is CENTRED on a per-row value
 * ("shift", ~ the row mean): the bf16 rounding then acts on the centred row, like LayerNorm-then-round does, and the
 * one-pass variance has nothing to cancel.
 *   shift_out f32 [rows]        = row mean of `out`
 *   xb_out    bf16 [rows,1024]  = bf16(out - shift)
 *   stats_out f32 [rows][8][2]  = (sum, sum of squares) of (out - shift) per row (block 0 carries the whole row, 1..7 zero)
 * This is the tower's pre_layrnorm. */
HVLM_API int hvlm_layernorm_1024_stats(const float* x, const float* gamma, const float* beta, float* out, void* xb_out,
                                       float* stats_out, float* shift_out, int rows, float eps, void* stream);
/* LayerNorm folded into its consumer GEMM (replaces LayerNorm + Linear of HF CLIPEncoderLayer, modeling_clip.py, as run by
 * clip_encoder.py:39-51):   out[i,n] = act( rstd_i * (sum_k xb[i,k] w_f[n,k] - mean_i * c[n]) + b_f[n] )
 *   xb     bf16 [M,1024]  the un-normalised rows (minus any per-row constant: see hvlm_layernorm_1024_stats);
 *   stats  f32 [M][8][2]  (sum, sum of squares) of those same rows per 128-column block
 *   shift_io f32 [M] or NULL: the running row mean kept for the NEXT producer -- shift_io[i] += mean of row i of xb
 *   w_f, c, b_f           as described at hvlm_vit_layout.fold;  N % 256 == 0, K = 1024
 *   epilogue              HVLM_EPI_BIAS or HVLM_EPI_BIAS_QUICKGELU, bf16 output [M,N];  qkv_hm != 0 (N = 3072, bias
 *                         epilogue): column-block-major [48][M][64] output as hvlm_vit_qkv_gemm */
HVLM_API int hvlm_gemm_ln_fold_bf16(const void* xb, const float* stats, const void* w_f, const float* c, const float* b_f,
                                    void* out, int M, int N, int epilogue, int qkv_hm, float eps, float* shift_io,
                                    void* stream);
/* Residual GEMM that feeds a folded LayerNorm:  hidden[M,1024] (f32, in place) += A[M,K] B[1024,K]^T + bias, and the same
 * epilogue writes xb_out = bf16(hidden - shift[i]) and stats_out over (hidden - shift[i]) (layout above; shift f32 [M] or
 * NULL = 0).  hidden equals what hvlm_gemm_bf16 with HVLM_EPI_BIAS_RESIDUAL computes, bit for bit. */
HVLM_API int hvlm_gemm_resid_stats(const void* A, const void* B, const float* bias, float* hidden, const float* shift,
                                   void* xb_out, float* stats_out, int M, int K, void* stream);
/* y bf16 [M = n_frames*257, 1024] (LN1 output) -> qkv bf16 COLUMN-BLOCK-MAJOR [48][M][64]: column blocks
 * q0..q15 | k0..k15 | v0..v15 of (y W_qkv^T + b); q carries the 64^-1/2 scale (folded into the packed weights).
 * Every (frame, head) operand is a contiguous [257][64] block. */
HVLM_API int hvlm_vit_qkv_gemm(const void* y, const void* w_qkv, const float* b_qkv, void* qkv_hm, int n_frames,
                               void* stream);
/* qkv_hm (layout above) -> out bf16 [n_frames*257, 1024] = concat_heads(softmax(q k^T) v). */
HVLM_API int hvlm_vit_attention(const void* qkv_hm, void* out, int n_frames, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LITA slow-fast token pooling -- replaces
 *   LitaMetaForCausalLM.videos_to_tokens pooling branch     lita/model/lita_arch.py:41-73
 *   VisualToTokenHelper.compress_tokens                     hoi_forecast/model/visual_to_tokens.py:230-272
 * tok  [B, t, *, C]: element (b,f,s,c) at ((b*t+f)*frame_stride + s)*C + c  (frame_stride >= 256 lets the
 *      kernel read rows 1..256 of the tower's [*,257,1024] hidden state in place: pass hidden + 1024).
 * out  [B, n_out, C], n_out per hvlm_pool_mode.  Each input element is read exactly once.
 * fp32 accumulation; in/out dtype f32 or bf16.  C % 8 == 0.  t >= 1.
 * ---------------------------------------------------------------------------------------------- */
HVLM_API int hvlm_pool_out_tokens(int t, int mode);
HVLM_API int hvlm_pool_slowfast_fwd(const void* tok, int in_dtype, int64_t frame_stride, void* out, int out_dtype, int B,
                           int t, int C, int mode, void* stream);
/* Same, with a frame indirection: logical frame (b,f) is read from row block frame_map[b*t+f] of `tok` (frame
 * de-duplication: EPIC clips are 10 distinct frames tiled x10, handsonvlm/dataset/epic_dataset.py:90-95; single images are
 * tiled x100, hybrid_dataset.py:141-142 -- the tower then only encodes the distinct frames).  Modes 0,1,2 only. */
HVLM_API int hvlm_pool_slowfast_fwd_mapped(const void* tok, int in_dtype, int64_t frame_stride, const int32_t* frame_map,
                                           void* out, int out_dtype, int B, int t, int C, int mode, void* stream);
/* d_tok [B,t,256,C] (dense) = autograd of the above wrt tok; dout [B,n_out,C]. */
HVLM_API int hvlm_pool_slowfast_bwd(const void* dout, int dout_dtype, void* dtok, int dtok_dtype, int B, int t, int C,
                           int mode, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Token splice -- replaces prepare_inputs_labels_for_multimodal
 *   llava/model/llava_arch.py:110-234            (HVLM_SPLICE_LLAVA)
 *   handsonvlm/.../handsonvlm.py:212-451         (HVLM_SPLICE_HANDSONVLM)
 * Step 1  hvlm_splice_count : per-sample number of IMAGE_TOKEN_INDEX (-200) tokens (device int32 [B]).
 * Step 2  hvlm_splice_plan  : index plan.  src_index [B,L] int32: >=0 text position p (row = ids[b,p]);
 *                             -(1+g) visual row g of visual[n_img*Nv, D]; INT32_MIN = padding.
 *                             hand_code [B,L] int8: -1 or k = which hand positional embedding is added.
 *                             lens [B] int32 = spliced length of each sample.  status: device int32, error bits
 *                             OR-ed in (HVLM_PLAN_*).  Image-slot bookkeeping follows cur_image_idx
 *                             (a sample without image token still consumes a slot).
 * Step 3  hvlm_splice_fwd   : index-driven row gather into embeds [B,L,D] (+ labels [B,L] i64, mask [B,L] u8),
 *                             fusing the embedding lookup (embed_tokens) and the sinusoidal hand embedding
 *                             (process_traj_positional_embedding, handsonvlm.py:310-338).
 * L is chosen by the caller: T-1+Nv when every sample has exactly one image token (the collator's
 * contract, hybrid_dataset.py:155-158), else max(lens) after reading `lens` back.
 * last_visual_end (optional device int64 scalar): the reference's `self.last_visual_token_index` side effect
 *                             (handsonvlm.py:288) = position of the last image token of the last sample that has one,
 *                             relative to the ids left after the previous image token, + its number of visual rows.
 *                             Left untouched when no sample has an image token.
 * Any number of image tokens per sample (round 1 capped it at 64).
 * ---------------------------------------------------------------------------------------------- */
/* OR-able into `variant` of hvlm_splice_plan / hvlm_splice_fwd: the `tune_mm_mlp_adapter && mm_use_im_start_end` branch of
 * llava_arch.py:146-161,172-173 -- the embeddings are the same rows, but the token right after an image token (<im_end>)
 * takes the label of the image-token position (`cur_labels[image_token_start:image_token_start+1]`, :159); in the
 * HandsOnVLM splice (handsonvlm.py:263-286) the position-spliced attention mask follows the same rule (:283) and the
 * caller passes hand_mode 0 (that branch adds no hand embeddings, :343-344). */
#define HVLM_SPLICE_FLAG_IM_START_END 0x100
#define HVLM_IGNORE_INDEX (-100)
#define HVLM_IMAGE_TOKEN_INDEX (-200)
#define HVLM_HAND_TRAJ_TOKEN_ID 32100

#define HVLM_PLAN_ERR_LEN_OVERFLOW 1    /* a spliced sample is longer than L                      */
#define HVLM_PLAN_ERR_IMG_OVERFLOW 2    /* more image slots consumed than n_img                   */
#define HVLM_PLAN_ERR_HAND_COUNT 4      /* training: >4 hand tokens; eval: count != n_hand points */
#define HVLM_PLAN_ERR_BAD_ID 8          /* token id outside [0,vocab) (and not -200)              */
#define HVLM_PLAN_NOT_UNIFORM 16        /* lens differ between samples (ragged output)            */

HVLM_API int hvlm_splice_count(const int64_t* ids, int B, int T, int32_t* counts, void* stream);
/* Step 1 with everything the HOST side needs to size the output and to raise the reference's errors, from the ids alone:
 * info int32 [4][B] = image-token count (usable as `counts` of hvlm_splice_plan) | position of the last image token (-1:
 * none) | number of <hand_traj> tokens after it | 1 if some id is outside [0, vocab) and is not the image token.
 * Launched before the visual pipeline of the same call and copied out asynchronously, it is on the host long before the
 * splice needs it: the reference's per-sample host syncs (llava_arch.py:127,137; handsonvlm.py:234,247) without a stall. */
HVLM_API int hvlm_splice_info(const int64_t* ids, int B, int T, int vocab, int32_t* info, void* stream);
HVLM_API int hvlm_splice_plan(const int64_t* ids, const int32_t* counts /*from hvlm_splice_count*/, int B, int T, int Nv,
                     int n_img, int L, int vocab, int variant,
                     int hand_mode /*0 none, 1 training (4 points, cnt/4 scaling), 2 eval (n points)*/,
                     int n_hand_points, int32_t* src_index, int8_t* hand_code, int32_t* lens,
                     float* hand_scale /*[B]*/, int32_t* status,
                     int64_t* last_visual_end /*NULL ok: device scalar, see below*/, void* stream);
/* Same plan for visual token blocks of DIFFERENT lengths per image slot -- the list path of images_to_tokens
 * (llava_arch.py:95-106: each sample's image group becomes one flat [n_i*256, D] block and its single image token expands to
 * all of it).  slot_offsets int32 [n_slots+1] (device): rows of slot g are visual[slot_offsets[g] .. slot_offsets[g+1]) of
 * the row-concatenated visual tensor that hvlm_splice_fwd / _bwd receive. */
HVLM_API int hvlm_splice_plan_ragged(const int64_t* ids, const int32_t* counts, const int32_t* slot_offsets, int B, int T,
                            int n_slots, int L, int vocab, int variant, int hand_mode, int n_hand_points,
                            int32_t* src_index, int8_t* hand_code, int32_t* lens, float* hand_scale, int32_t* status,
                            int64_t* last_visual_end /*NULL ok*/, void* stream);
HVLM_API int hvlm_splice_fwd(const int32_t* src_index, const int8_t* hand_code, const int32_t* lens, const float* hand_scale,
                    const int64_t* ids, const int64_t* labels /*NULL ok*/, const uint8_t* mask /*NULL ok*/,
                    const void* embed_table, const void* visual, const uint8_t* visual_mask /*NULL = all true*/,
                    const float* future_hands /*[B,2,n,2] f32 or NULL*/, int n_hand_points, int B, int T, int L,
                    int Nv, int D, int dtype, int variant, void* out_embeds, int64_t* out_labels,
                    uint8_t* out_mask, void* stream);
/* backward of the copy: d_visual [n_img*Nv, D] f32 (written, each row at most once) and scatter-ADD of the text
 * rows into d_embed_table [vocab, D] f32 (atomics; caller zero-initialises).  Either output may be NULL. */
HVLM_API int hvlm_splice_bwd(const void* d_embeds, int dtype, const int32_t* src_index, const int64_t* ids, int B, int T,
                    int L, int n_visual_rows, int D, float* d_visual, float* d_embed_table, void* stream);

/* ------------------------------------------------------------------------------------------------
 * <hand_traj> hidden-state gather -- replaces the inline loop of HandsOnVLMForCausalLM.forward
 *   handsonvlm/.../handsonvlm.py:146-187  and the generation-time gather :609-622
 * rows[b,k] = k-th position i with labels[b,i+1] == hand_id (i.e. the state that predicts the token);
 * out[b,h,k,j] = hidden[b, rows[b,k], 2j+h]; samples without hand tokens give zeros and valid[b]=0.
 * counts[b] = number of such positions (must be 0 or 4 in the reference; the caller decides how to react).
 * ---------------------------------------------------------------------------------------------- */
HVLM_API int hvlm_hand_gather_fwd(const void* hidden, int dtype, const int64_t* labels, int64_t hand_id, int B, int L, int D,
                         void* out, uint8_t* valid, int32_t* rows /*[B,4]*/, int32_t* counts /*[B]*/, void* stream);
/* d_hidden [B,L,D] f32 += scatter of dout [B,2,4,D/2]; caller zero-initialises d_hidden. */
HVLM_API int hvlm_hand_gather_bwd(const void* dout, int dtype, const int32_t* rows, int B, int L, int D, float* d_hidden,
                         void* stream);
/* generation step: hidden_last [B,D] -> out [B,2,1,D/2] (even/odd de-interleave). */
HVLM_API int hvlm_hand_gather_step(const void* hidden_last, int dtype, int B, int D, void* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * trajectory head after the gather, generation side -- replaces, in one launch, the <hand_traj> branch of the sampling
 * loop  handsonvlm/.../handsonvlm.py:609-622  ->  TrajDecoder.inference  handsonvlm/.../traj_decoder.py:39-47
 *   ->  TrajCVAE.inference  hoi_forecast/architecture/traj_decoder.py:75-91  ->  VAE.inference  decoder_modules.py:56-60:
 *   out[r,:] = W2 . ELU(W1 . cat(z[r], cond[r]) + b1) + b2        W1 [H, L+Dc], W2 [2, H]   (dec_MLP.0 / dec_MLP.2)
 * cond source: interleaved = 0: cond [R, Dc] rows with stride ld_cond (elements) -- the reshaped pred_hand_embeddings;
 *              interleaved = 1: cond = last hidden rows [R/2, >= 2*Dc], row r = (b, hand) reads hidden[b, 2j + hand]
 *              (the gather of handsonvlm.py:613-616 fused in; R must be even).
 * z [R, L] is the caller-drawn, already scaled noise (the reference draws z_scale * randn itself, traj_decoder.py:87).
 * All tensors share `dtype`; out [R, 2] is fp32.  workspace: >= hvlm_traj_decode_workspace_bytes(R, H) bytes whose first
 * 4096 bytes (arrival counters) are zeroed ONCE by the caller before the first use -- every call leaves them zeroed,
 * whatever its R and H, so one buffer sized for the largest call can be reused; one workspace per stream.
 * Accumulation order is fixed: results are bit-reproducible run to run.
 * ---------------------------------------------------------------------------------------------- */
HVLM_API size_t hvlm_traj_decode_workspace_bytes(int R, int H);
HVLM_API int hvlm_traj_decode(const void* cond, int64_t ld_cond, int interleaved, const void* z, const void* W1, const void* b1,
                     const void* W2, const void* b2, int dtype, int R, int Dc, int L, int H, float* out, void* workspace,
                     size_t workspace_bytes, void* stream);

/* out[r, j] = act(b[j] + W[j, :] . x[r, :]) for a handful of rows (R = 2B in the generation loop): the layers of the MLP
 * trajectory decoder, TrajMLP.inference  hoi_forecast/architecture/traj_decoder.py:94-104,139-147  (Linear-ReLU-Linear-
 * ReLU-Linear).  x [R, K] with row stride ld_x, W [H, K], b [H] or NULL, out [R, H]; all in `dtype`, fp32 accumulation.
 * act: 0 none, 1 ReLU, 2 ELU. */
HVLM_API int hvlm_skinny_linear(const void* x, int64_t ld_x, const void* W, const void* b, int act, int dtype, int R, int K,
                       int H, void* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Frame de-duplication in front of the tower.  Real clips repeat frames: EPIC clips are 10 distinct
 * frames tiled x10 (handsonvlm/dataset/epic_dataset.py:90-95, handsonvlm/evaluation/handsonvlm_inference.py:205), single
 * images are tiled x100 (handsonvlm/dataset/hybrid_dataset.py:141-142); the reference encodes every copy.
 * frames   [n_frames] rows of frame_bytes bytes (any dtype / layout; frame_bytes % 16 == 0, base 16-byte aligned)
 * frame_map int32 [n_frames]: index of frame i's representative in the compacted list (-> hvlm_pool_slowfast_fwd_mapped)
 * rep       int32 [n_frames]: rep[u] = first frame with unique index u, u < *n_unique; entries past *n_unique are 0, so a
 *           gather of a fixed capacity (hvlm_gather_rows with n_out = capacity) is always in bounds
 * n_unique  int32 device scalar
 * capacity  0, or the number of distinct frames the caller has sized its buffers for WITHOUT reading n_unique back (a
 *           per-dataset contract, e.g. 10 per EPIC clip): frame_map is then clamped to capacity-1 so that every later
 *           access stays in bounds, and *n_unique > capacity tells the caller (asynchronously) that the contract broke
 * Two frames are duplicates iff their bytes are equal: a 64-bit checksum groups candidates (one streaming pass), a
 * byte-wise comparison against the first frame of the group confirms them (second pass over the duplicates only); a
 * checksum collision just leaves the frame un-deduplicated.  Deterministic.  No host synchronisation.
 * workspace: hvlm_frame_dedup_workspace_bytes(n_frames) bytes, 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
HVLM_API size_t hvlm_frame_dedup_workspace_bytes(int n_frames);
HVLM_API int hvlm_frame_dedup(const void* frames, size_t frame_bytes, int n_frames, int capacity, int32_t* frame_map,
                              int32_t* rep, int32_t* n_unique, void* workspace, size_t workspace_bytes, void* stream);
/* dst[i, :] = src[idx[i], :] for i < n_out; rows of row_bytes bytes (% 16 == 0); idx int32 (device), values in [0, n_src). */
HVLM_API int hvlm_gather_rows(const void* src, size_t row_bytes, int n_src, const int32_t* idx, int n_out, void* dst,
                              void* stream);

/* ------------------------------------------------------------------------------------------------
 * CLIPImageProcessor resize + centre crop on raw decoded frames -- replaces the CPU side of
 *   hoi_forecast/dataset/video_utils.py:28-53 (processor.preprocess per frame; transformers==4.31.0
 *   CLIPImageProcessor: resize shortest edge -> 224 with PIL BICUBIC, centre crop 224x224; the rescale + normalise that
 *   follow are fused into hvlm_vit_l14_fwd_u8).
 * The arithmetic is Pillow's 8-bit ImagingResample restated: separable filter, horizontal pass then vertical pass, 22-bit
 * fixed-point coefficients, rounding + clipping to uint8 after EACH pass -- results are bit-identical to
 * PIL.Image.resize(..., BICUBIC) followed by the crop.
 * Host side (pure host code, no CUDA): hvlm_resize_plan_host derives the geometry the way transformers does
 *   (get_resize_output_image_size with an int size / center_crop), hvlm_resize_tables_host fills the fixed-point
 *   coefficient table (plan->table_ints int32: xb [2*out_w] | xc [out_w*xk] | yb [2*out_h] | yc [out_h*yk], b = (first
 *   source index, taps) per output index of the crop window); the caller copies the table to the device once per geometry.
 *   hvlm_resize_table_host is the single-axis building block (Pillow precompute_coeffs + normalize_coeffs_8bpc for the
 *   output window [crop0, crop0 + crop_n) of an in_size -> out_size pass; returns ksize, or with coef_host == NULL only
 *   queries it).
 * hvlm_resize_crop_u8: src uint8 [N, in_h, in_w, 3] -> dst uint8 [N, out_h, out_w, 3] (dst 4-byte aligned).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t in_h, in_w;     /* decoded frame size                                                  */
    int32_t new_h, new_w;   /* size after the shortest-edge resize                                 */
    int32_t top, left;      /* centre-crop offsets inside the resized image                        */
    int32_t out_h, out_w;   /* crop size (224 x 224 for CLIP ViT-L/14)                             */
    int32_t xk, yk;         /* filter taps per output column / row                                 */
    int32_t x_lo, x_cols;   /* source columns [x_lo, x_lo + x_cols) the crop window reads          */
    int32_t rows_cap;       /* most source rows any block of 8 output rows reads                   */
    int32_t table_ints;     /* int32 entries of the coefficient table                              */
} hvlm_resize_plan;
HVLM_API int hvlm_resize_plan_host(int in_h, int in_w, int shortest_edge, int crop, hvlm_resize_plan* plan_host);
HVLM_API int hvlm_resize_tables_host(const hvlm_resize_plan* plan_host, int32_t* table_host);
HVLM_API int hvlm_resize_table_host(int in_size, int out_size, int crop0, int crop_n, int32_t* bounds_host,
                                    int32_t* coef_host);
HVLM_API int hvlm_resize_crop_u8(const uint8_t* src, int N, const hvlm_resize_plan* plan_host, const int32_t* table,
                                 uint8_t* dst, void* stream);
/* expand2square (hoi_forecast/dataset/video_utils.py:13-25), the `image_aspect_ratio == 'pad'` branch of load_image
 * (:30-31): src uint8 [N,H,W,3] pasted in the middle of dst uint8 [N,S,S,3], S = max(H,W) (S % 4 == 0), filled with the
 * background colour (the reference passes tuple(int(255 * m) for m in processor.image_mean) = (122, 116, 104)). */
HVLM_API int hvlm_pad_square_u8(const uint8_t* src, int N, int H, int W, uint8_t* dst, int bg_r, int bg_g, int bg_b,
                                void* stream);

/* ------------------------------------------------------------------------------------------------
 * helpers for the training-shaped variant
 * ---------------------------------------------------------------------------------------------- */
/* out bf16 [C, R_pad] = in[R, C]^T (zero padded to R_pad, a multiple of 8): makes dY / X K-major for wgrad. */
HVLM_API int hvlm_transpose_to_bf16(const void* in, int in_dtype, void* out, int R, int C, int R_pad, void* stream);
/* db[n] = sum_m dY[m,n]  (fp32) */
HVLM_API int hvlm_colsum(const void* dy, int dtype, float* db, int M, int N, void* stream);

/* ------------------------------------------------------------------------------------------------
 * measurement hooks (bench.py): number of kernels this library has launched in the process, and optional
 * per-stage device timing with CUDA events recorded on the launching stream around every launch.
 * ---------------------------------------------------------------------------------------------- */
typedef enum {
    HVLM_STAGE_IM2COL = 0,
    HVLM_STAGE_PATCH_GEMM = 1,
    HVLM_STAGE_LAYERNORM = 2,
    HVLM_STAGE_QKV_GEMM = 3,
    HVLM_STAGE_ATTENTION = 4,
    HVLM_STAGE_OUTPROJ_GEMM = 5,
    HVLM_STAGE_FC1_GEMM = 6,
    HVLM_STAGE_FC2_GEMM = 7,
    HVLM_STAGE_POOL = 8,
    HVLM_STAGE_GEMM = 9,        /* hvlm_gemm_bf16: projector fwd / wgrad */
    HVLM_STAGE_SPLICE = 10,
    HVLM_STAGE_GATHER = 11,
    HVLM_STAGE_OTHER = 12,
    HVLM_STAGE_COUNT = 13
} hvlm_stage;

HVLM_API uint64_t hvlm_launch_count(void);
HVLM_API int hvlm_profile_enable(int on);
/* synchronises the recorded events, sums device ms and launches per stage, clears the record list */
HVLM_API int hvlm_profile_collect(float* ms_by_stage_host /*[HVLM_STAGE_COUNT]*/, int32_t* launches_by_stage_host);

#ifdef __cplusplus
}
#endif
#endif /* HVLM_B200_H */
