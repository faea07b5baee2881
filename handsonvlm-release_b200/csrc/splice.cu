// Visual-token splice: index plan + index-driven row gather (forward) and gather / scatter-add (backward).
//
// Replaces prepare_inputs_labels_for_multimodal
//   llava/model/llava_arch.py:110-234                                  (HVLM_SPLICE_LLAVA)
//   handsonvlm/model/language_model/handsonvlm.py:212-451              (HVLM_SPLICE_HANDSONVLM)
// The reference walks the batch in Python: per sample ~10 tiny kernels (embed_tokens x2, cat x3, full, where)
// and 3 host syncs.  Here: count -> plan -> gather, three launches per batch, no host sync.  The gather fuses
// the embed_tokens lookup, the visual rows, labels / mask construction, padding, and the sinusoidal hand
// positional embedding (process_traj_positional_embedding, handsonvlm.py:310-338).
#include <limits.h>

#include "hvlm_internal.cuh"
#include "hvlm_scan.cuh"
#include "hvlm_vec.cuh"

namespace hvlm {

constexpr int kPad = INT_MIN;

__global__ void splice_count_kernel(const int64_t* __restrict__ ids, int T, int32_t* __restrict__ counts) {
    __shared__ int wsum[kPlanThreads / 32];
    const int b = blockIdx.x;
    int c = 0;
    for (int p = threadIdx.x; p < T; p += blockDim.x) c += (ids[static_cast<int64_t>(b) * T + p] == HVLM_IMAGE_TOKEN_INDEX);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < kPlanThreads / 32; ++w) s += wsum[w];
        counts[b] = s;
    }
}

// Everything the HOST needs to size the output and to raise the reference's errors, from the ids alone -- so it can be
// launched (and copied out asynchronously) BEFORE the visual pipeline of the same call and read without stalling once that
// pipeline has been enqueued.  info is [4][B]: image-token count | position of the last image token (-1: none) | number of
// <hand_traj> tokens after it | 1 if some id is outside [0, vocab) and not the image token.
__global__ void __launch_bounds__(kPlanThreads)
splice_info_kernel(const int64_t* __restrict__ ids, int B, int T, int vocab, int32_t* __restrict__ info) {
    __shared__ int s_cnt, s_last, s_bad, s_hand;
    const int b = blockIdx.x;
    const int64_t* row = ids + static_cast<int64_t>(b) * T;
    if (threadIdx.x == 0) {
        s_cnt = 0;
        s_last = -1;
        s_bad = 0;
        s_hand = 0;
    }
    __syncthreads();
    int cnt = 0, last = -1, bad = 0;
    for (int p = threadIdx.x; p < T; p += blockDim.x) {
        const int64_t tok = row[p];
        if (tok == HVLM_IMAGE_TOKEN_INDEX) {
            ++cnt;
            last = p;
        } else if (tok < 0 || tok >= vocab) {
            bad = 1;
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    last = __reduce_max_sync(0xffffffffu, last);
    bad = __reduce_or_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_cnt, cnt);
        atomicMax(&s_last, last);
        atomicOr(&s_bad, bad);
    }
    __syncthreads();
    const int last_pos = s_last;
    int hand = 0;
    for (int p = last_pos + 1 + threadIdx.x; p < T; p += blockDim.x) hand += (row[p] == HVLM_HAND_TRAJ_TOKEN_ID);
    hand = __reduce_add_sync(0xffffffffu, hand);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_hand, hand);
    __syncthreads();
    if (threadIdx.x == 0) {
        info[b] = s_cnt;
        info[B + b] = last_pos;
        info[2 * B + b] = s_hand;
        info[3 * B + b] = s_bad;
    }
}

__global__ void __launch_bounds__(kPlanThreads)
splice_plan_kernel(const int64_t* __restrict__ ids, const int32_t* __restrict__ counts,
                   const int32_t* __restrict__ slot_offsets /* NULL: every slot has Nv rows */, int B, int T, int Nv,
                   int n_img, int L, int vocab, int variant, int hand_mode, int n_hand_points,
                   int32_t* __restrict__ src_index, int8_t* __restrict__ hand_code, int32_t* __restrict__ lens,
                   float* __restrict__ hand_scale, int32_t* __restrict__ status,
                   int64_t* __restrict__ last_visual_end) {
    __shared__ int scan_smem[9];
    __shared__ int chunk_pos[kPlanThreads];     // positions of the image tokens found in the current chunk of ids
    const int b = blockIdx.x;
    const int64_t* row = ids + static_cast<int64_t>(b) * T;
    int32_t* dst = src_index + static_cast<int64_t>(b) * L;
    int8_t* hc = hand_code + static_cast<int64_t>(b) * L;
    int err = 0;

    // image-slot base: a sample with k image tokens consumes max(k,1) slots (cur_image_idx bookkeeping)
    int slot0 = 0;
    for (int i = 0; i < b; ++i) slot0 += max(counts[i], 1);
    const int k_img = counts[b];
    // rows of visual slot g: uniform Nv, or slot_offsets[g+1] - slot_offsets[g] when the caller hands in per-sample token
    // blocks of different lengths (the list path of images_to_tokens, llava_arch.py:95-106); slots past n_img have none
    auto nv_of = [&](int g) { return slot_offsets ? (g < n_img ? slot_offsets[g + 1] - slot_offsets[g] : 0) : Nv; };
    // rows ADDED by the first m image tokens of the sample whose first slot is s0 (each token is replaced by its block):
    // closed forms, so any number of image tokens per sample is fine (slot_offsets is itself a prefix sum)
    auto cum_rows = [&](int s0, int m) {
        if (!slot_offsets) return m * (Nv - 1);
        const int g0 = min(s0, n_img), g1 = min(s0 + m, n_img);
        return slot_offsets[g1] - slot_offsets[g0] - m;
    };
    const int len = T + cum_rows(slot0, k_img);
    const int len0 = T + cum_rows(0, counts[0]);
    if (len > L) err |= HVLM_PLAN_ERR_LEN_OVERFLOW;
    // a sample without image token advances cur_image_idx but never indexes the features (llava_arch.py:127-135,
    // handsonvlm.py:234-245), so only samples that READ slots can overflow
    if (k_img > 0 && slot0 + k_img > n_img) err |= HVLM_PLAN_ERR_IMG_OVERFLOW;
    if (len != len0) err |= HVLM_PLAN_NOT_UNIFORM;

    // defaults: padding + no hand code
    for (int r = threadIdx.x; r < L; r += blockDim.x) {
        if (r >= len) dst[r] = kPad;
        hc[r] = -1;
    }

    // text rows, image-token positions, visual rows -- one 256-token chunk of the ids at a time
    int img_seen = 0;
    int last_pos = -1, prev_pos = -1;            // positions of the last two image tokens seen so far (block-uniform)
    for (int p0 = 0; p0 < T; p0 += kPlanThreads) {
        const int p = p0 + threadIdx.x;
        const int64_t tok = p < T ? row[p] : 0;
        const int is_img = (p < T) && (tok == HVLM_IMAGE_TOKEN_INDEX);
        int total;
        const int local = block_excl_scan(is_img, &total, scan_smem);
        const int before = img_seen + local;
        if (p < T) {
            if (is_img) {
                chunk_pos[local] = p;
            } else {
                const int opos = p + cum_rows(slot0, before);
                const bool bad = tok < 0 || tok >= vocab;   // never index the table out of bounds
                if (bad) err |= HVLM_PLAN_ERR_BAD_ID;
                if (opos >= 0 && opos < L) dst[opos] = bad ? kPad : p;
            }
        }
        __syncthreads();
        for (int jj = 0; jj < total; ++jj) {
            const int j = img_seen + jj;
            const int o0 = chunk_pos[jj] + cum_rows(slot0, j);
            const int g = slot0 + j;
            const bool slot_ok = g < n_img;      // never plan a read past the visual tensor (error flagged above)
            const int g0 = slot_offsets ? (slot_ok ? slot_offsets[g] : 0) : g * Nv;
            const int nv = nv_of(g);
            for (int r = threadIdx.x; r < nv; r += blockDim.x)
                if (o0 + r >= 0 && o0 + r < L) dst[o0 + r] = slot_ok ? -(1 + g0 + r) : kPad;
        }
        if (total >= 2) prev_pos = chunk_pos[total - 2];
        else if (total == 1) prev_pos = last_pos;
        if (total >= 1) last_pos = chunk_pos[total - 1];
        img_seen += total;
        __syncthreads();                         // chunk_pos is rewritten by the next chunk
    }

    // hand positional embedding codes: only the tail segment (text after the LAST image token) of samples
    // that have an image token (handsonvlm.py:342-396)
    float scale = 0.f;
    if (variant == HVLM_SPLICE_HANDSONVLM && hand_mode != 0 && k_img > 0) {
        const int tail0 = last_pos + 1;                       // first tail position
        const int shift = cum_rows(slot0, k_img);             // output row = p + shift
        int seen = 0;
        for (int p0 = tail0; p0 < T; p0 += kPlanThreads) {
            const int p = p0 + threadIdx.x;
            const int is_hand = (p < T) && (row[p] == HVLM_HAND_TRAJ_TOKEN_ID);
            int total;
            const int ord = seen + block_excl_scan(is_hand, &total, scan_smem);
            const int limit = hand_mode == 1 ? 4 : n_hand_points;
            if (is_hand && ord < limit && p + shift >= 0 && p + shift < L) hc[p + shift] = static_cast<int8_t>(ord);
            seen += total;
        }
        __syncthreads();
        if (hand_mode == 1) {
            if (seen > 4) err |= HVLM_PLAN_ERR_HAND_COUNT;
            scale = static_cast<float>(seen) / 4.0f;
            // zero.scatter(0, idx, emb) with idx padded by 0: row 0 of the tail receives emb[3] (last writer)
            if (seen > 0 && seen < 4 && tail0 < T && threadIdx.x == 0 && tail0 + shift >= 0 && tail0 + shift < L)
                hc[tail0 + shift] = 3;
        } else {
            if (tail0 < T && seen != n_hand_points) err |= HVLM_PLAN_ERR_HAND_COUNT;
            scale = 1.0f;
        }
    }
    if (threadIdx.x == 0) {
        lens[b] = len;
        if (hand_scale) hand_scale[b] = scale;
        // handsonvlm.py:288 side effect: `self.last_visual_token_index = image_token_start + n_visual_tokens`, overwritten
        // for every image token of every sample, where image_token_start indexes the ids that are LEFT after the previous
        // image token was cut off.  The value that survives is the one of the last image token of the last sample that
        // has one: only that sample's CTA writes.
        if (last_visual_end && k_img > 0) {
            bool last = true;
            for (int i = b + 1; i < B; ++i) last = last && counts[i] == 0;
            if (last) {
                const int rel = last_pos - (k_img > 1 ? prev_pos + 1 : 0);
                *last_visual_end = static_cast<int64_t>(rel) + nv_of(slot0 + k_img - 1);
            }
        }
    }
    if (err) atomicOr(status, err);
}

// sinusoidal hand embedding value for output channel `ch` (handsonvlm.py:310-338):
//   out[k, 2c+h] = enc(hand h, point k)[c],  enc = cat[sin(x f), cos(y f), sin(x f), cos(y f)],
//   f_i = 10000^(-2i/(D/4)), i in [0, D/8)
__device__ __forceinline__ float hand_embed_value(const float* __restrict__ fh /*[2,n,2] of sample*/, int n, int k,
                                                  int ch, int D) {
    const int h = ch & 1;
    const int c = ch >> 1;
    const int eighth = D >> 3;
    const int seg = c / eighth;
    const int i = c - seg * eighth;
    // 10000^(-2i/(D/4)) = 2^(-log2(10000) * 2i / (D/4)); arguments of sin/cos are coordinates in [0,1] times freq <= 1,
    // where the fast intrinsics are accurate to ~4e-7 absolute (tolerance on these rows: 1e-5)
    const float freq = exp2f(-13.287712379549449f * static_cast<float>(2 * i) / static_cast<float>(D >> 2));
    const float* pt = fh + (h * n + k) * 2;
    const float a = ((seg & 1) ? pt[1] : pt[0]) * freq;
    if (fabsf(a) > 4.0f) return (seg & 1) ? cosf(a) : sinf(a);   // un-normalised coordinates: precise path
    return (seg & 1) ? __cosf(a) : __sinf(a);
}

template <typename T>
__global__ void __launch_bounds__(128)
splice_fwd_kernel(const int32_t* __restrict__ src_index, const int8_t* __restrict__ hand_code,
                  const int32_t* __restrict__ lens, const float* __restrict__ hand_scale,
                  const int64_t* __restrict__ ids, const int64_t* __restrict__ labels,
                  const uint8_t* __restrict__ mask, const T* __restrict__ table, const T* __restrict__ visual,
                  const uint8_t* __restrict__ visual_mask, const float* __restrict__ future_hands, int n_hand, int Tlen,
                  int L, int Nv, int D, int variant, T* __restrict__ out, int64_t* __restrict__ out_labels,
                  uint8_t* __restrict__ out_mask) {
    constexpr int V = Vec16<T>::N;
    const int r = blockIdx.x, b = blockIdx.y;
    const int64_t orow = static_cast<int64_t>(b) * L + r;
    const int code = src_index[orow];
    const int hk = hand_code ? hand_code[orow] : -1;
    const T* src = nullptr;
    int64_t lab = HVLM_IGNORE_INDEX;
    uint8_t mk = 0;
    if (code >= 0) {
        const int64_t ipos = static_cast<int64_t>(b) * Tlen + code;
        src = table + ids[ipos] * D;
        // <im_end> right after an image token keeps the image-token position's label (and, in the HandsOnVLM splice,
        // its attention-mask entry) in the im_start_end branch: llava_arch.py:159, handsonvlm.py:278,283
        const bool after_img = (variant & HVLM_SPLICE_FLAG_IM_START_END) && code > 0 &&
                               ids[ipos - 1] == HVLM_IMAGE_TOKEN_INDEX;
        if (labels) lab = labels[after_img ? ipos - 1 : ipos];
        if (mask) mk = mask[after_img ? ipos - 1 : ipos];
    } else if (code != kPad) {
        const int g = -(code + 1);
        src = visual + static_cast<int64_t>(g) * D;
        mk = visual_mask ? visual_mask[g] : 1;
    }
    if (threadIdx.x == 0) {
        if (out_labels) out_labels[orow] = lab;
        if (out_mask) {
            if ((variant & 0xff) == HVLM_SPLICE_LLAVA && mask) {
                // llava_arch.py:215-232: True x (len - T) prepended, original mask, False right-pad
                const int len = lens[b];
                const int lead = len - Tlen;
                mk = r < lead ? 1 : (r < len ? mask[static_cast<int64_t>(b) * Tlen + (r - lead)] : 0);
            }
            out_mask[orow] = mk;
        }
    }
    T* dst = out + orow * D;
    const float hs = (hk >= 0 && hand_scale) ? hand_scale[b] : 0.f;
    const float* fh = (hk >= 0 && future_hands) ? future_hands + static_cast<int64_t>(b) * 2 * n_hand * 2 : nullptr;
    for (int c = threadIdx.x * V; c < D; c += blockDim.x * V) {
        uint4 raw = src ? *reinterpret_cast<const uint4*>(src + c) : make_uint4(0, 0, 0, 0);
        if (fh) {
            float f[V];
            unpack16<T>(raw, f);
#pragma unroll
            for (int i = 0; i < V; ++i) f[i] += hs * hand_embed_value(fh, n_hand, hk, c + i, D);
            raw = pack16<T>(f);
        }
        *reinterpret_cast<uint4*>(dst + c) = raw;
    }
}

template <typename T>
__global__ void __launch_bounds__(128)
splice_bwd_kernel(const T* __restrict__ d_embeds, const int32_t* __restrict__ src_index,
                  const int64_t* __restrict__ ids, int Tlen, int L, int n_visual_rows, int D,
                  float* __restrict__ d_visual, float* __restrict__ d_table) {
    constexpr int V = Vec16<T>::N;
    const int r = blockIdx.x, b = blockIdx.y;
    const int64_t orow = static_cast<int64_t>(b) * L + r;
    const int code = src_index[orow];
    if (code == kPad) return;
    const T* src = d_embeds + orow * D;
    if (code >= 0) {
        if (!d_table) return;
        float* dst = d_table + ids[static_cast<int64_t>(b) * Tlen + code] * D;
        for (int c = threadIdx.x * V; c < D; c += blockDim.x * V) {
            float f[V];
            unpack16<T>(*reinterpret_cast<const uint4*>(src + c), f);
#pragma unroll
            for (int i = 0; i < V; ++i) atomicAdd(dst + c + i, f[i]);
        }
    } else {
        const int g = -(code + 1);
        if (!d_visual || g >= n_visual_rows) return;
        float* dst = d_visual + static_cast<int64_t>(g) * D;
        for (int c = threadIdx.x * V; c < D; c += blockDim.x * V) {
            float f[V];
            unpack16<T>(*reinterpret_cast<const uint4*>(src + c), f);
#pragma unroll
            for (int i = 0; i < V; ++i) dst[c + i] = f[i];
        }
    }
}

}  // namespace hvlm

extern "C" int hvlm_splice_count(const int64_t* ids, int B, int T, int32_t* counts, void* stream) {
    using namespace hvlm;
    if (!ids || !counts || B <= 0 || T <= 0) return HVLM_ERR_BAD_ARG;
    StageTimer st(HVLM_STAGE_SPLICE, static_cast<cudaStream_t>(stream));
    splice_count_kernel<<<B, kPlanThreads, 0, static_cast<cudaStream_t>(stream)>>>(ids, T, counts);
    return check_last("splice_count");
}

extern "C" int hvlm_splice_info(const int64_t* ids, int B, int T, int vocab, int32_t* info, void* stream) {
    using namespace hvlm;
    if (!ids || !info || B <= 0 || T <= 0 || vocab <= 0) return HVLM_ERR_BAD_ARG;
    StageTimer st(HVLM_STAGE_SPLICE, static_cast<cudaStream_t>(stream));
    splice_info_kernel<<<B, kPlanThreads, 0, static_cast<cudaStream_t>(stream)>>>(ids, B, T, vocab, info);
    return check_last("splice_info");
}

static int splice_plan_impl(const int64_t* ids, const int32_t* counts, const int32_t* slot_offsets, int B, int T, int Nv,
                            int n_img, int L, int vocab, int variant, int hand_mode, int n_hand_points,
                            int32_t* src_index, int8_t* hand_code, int32_t* lens, float* hand_scale, int32_t* status,
                            int64_t* last_visual_end, void* stream) {
    using namespace hvlm;
    if (!ids || !counts || !src_index || !hand_code || !lens || !status) return HVLM_ERR_BAD_ARG;
    if (B <= 0 || T <= 0 || Nv <= 0 || n_img <= 0 || L <= 0 || vocab <= 0) return HVLM_ERR_BAD_ARG;
    if ((variant & 0xff) != HVLM_SPLICE_LLAVA && (variant & 0xff) != HVLM_SPLICE_HANDSONVLM) return HVLM_ERR_BAD_ARG;
    variant &= 0xff;        // the plan does not depend on the flags
    if (hand_mode < 0 || hand_mode > 2 || n_hand_points < 0 || n_hand_points > 127) return HVLM_ERR_BAD_ARG;
    StageTimer st(HVLM_STAGE_SPLICE, static_cast<cudaStream_t>(stream));
    splice_plan_kernel<<<B, kPlanThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        ids, counts, slot_offsets, B, T, Nv, n_img, L, vocab, variant, hand_mode, n_hand_points, src_index, hand_code,
        lens, hand_scale, status, last_visual_end);
    return check_last("splice_plan");
}

extern "C" int hvlm_splice_plan(const int64_t* ids, const int32_t* counts, int B, int T, int Nv, int n_img, int L,
                                int vocab, int variant, int hand_mode, int n_hand_points, int32_t* src_index,
                                int8_t* hand_code, int32_t* lens, float* hand_scale, int32_t* status,
                                int64_t* last_visual_end, void* stream) {
    return splice_plan_impl(ids, counts, nullptr, B, T, Nv, n_img, L, vocab, variant, hand_mode, n_hand_points, src_index,
                            hand_code, lens, hand_scale, status, last_visual_end, stream);
}

extern "C" int hvlm_splice_plan_ragged(const int64_t* ids, const int32_t* counts, const int32_t* slot_offsets, int B,
                                       int T, int n_slots, int L, int vocab, int variant, int hand_mode,
                                       int n_hand_points, int32_t* src_index, int8_t* hand_code, int32_t* lens,
                                       float* hand_scale, int32_t* status, int64_t* last_visual_end, void* stream) {
    if (!slot_offsets) return HVLM_ERR_BAD_ARG;
    return splice_plan_impl(ids, counts, slot_offsets, B, T, 1, n_slots, L, vocab, variant, hand_mode, n_hand_points,
                            src_index, hand_code, lens, hand_scale, status, last_visual_end, stream);
}

extern "C" int hvlm_splice_fwd(const int32_t* src_index, const int8_t* hand_code, const int32_t* lens,
                               const float* hand_scale, const int64_t* ids, const int64_t* labels,
                               const uint8_t* mask, const void* embed_table, const void* visual,
                               const uint8_t* visual_mask, const float* future_hands, int n_hand_points, int B, int T,
                               int L, int Nv, int D, int dtype, int variant, void* out_embeds, int64_t* out_labels,
                               uint8_t* out_mask, void* stream) {
    using namespace hvlm;
    if (!src_index || !lens || !ids || !embed_table || !visual || !out_embeds) return HVLM_ERR_BAD_ARG;
    if (B <= 0 || T <= 0 || L <= 0 || Nv <= 0 || D <= 0) return HVLM_ERR_BAD_ARG;
    if (future_hands && (!hand_code || !hand_scale || n_hand_points <= 0)) return HVLM_ERR_BAD_ARG;
    if (!aligned16(embed_table) || !aligned16(visual) || !aligned16(out_embeds)) return HVLM_ERR_ALIGN;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    StageTimer st(HVLM_STAGE_SPLICE, s);
    HVLM_DISPATCH_DTYPE(dtype, TT, {
        if (D % Vec16<TT>::N != 0 || (future_hands && (D % 8) != 0)) return HVLM_ERR_BAD_SHAPE;
        splice_fwd_kernel<TT><<<dim3(L, B), 128, 0, s>>>(
            src_index, hand_code, lens, hand_scale, ids, labels, mask, static_cast<const TT*>(embed_table),
            static_cast<const TT*>(visual), visual_mask, future_hands, n_hand_points, T, L, Nv, D, variant,
            static_cast<TT*>(out_embeds), out_labels, out_mask);
    });
    return check_last("splice_fwd");
}

extern "C" int hvlm_splice_bwd(const void* d_embeds, int dtype, const int32_t* src_index, const int64_t* ids, int B,
                               int T, int L, int n_visual_rows, int D, float* d_visual, float* d_embed_table,
                               void* stream) {
    using namespace hvlm;
    if (!d_embeds || !src_index || !ids || B <= 0 || T <= 0 || L <= 0 || D <= 0) return HVLM_ERR_BAD_ARG;
    if (!aligned16(d_embeds)) return HVLM_ERR_ALIGN;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HVLM_DISPATCH_DTYPE(dtype, TT, {
        if (D % Vec16<TT>::N != 0) return HVLM_ERR_BAD_SHAPE;
        splice_bwd_kernel<TT><<<dim3(L, B), 128, 0, s>>>(static_cast<const TT*>(d_embeds), src_index, ids, T, L,
                                                        n_visual_rows, D, d_visual, d_embed_table);
    });
    return check_last("splice_bwd");
}
