"""Drop-in mixins / helpers that keep the reference's call signatures for the visual-token path.

  reference                                                             here
  -----------------------------------------------------------------    -------------------------------
  LlavaMetaForCausalLM          llava/model/llava_arch.py:73-234        LlavaMetaForCausalLM
  LitaMetaForCausalLM           lita/model/lita_arch.py:17-85           LitaMetaForCausalLM
  HandsOnVLMMetaForCausalLM     handsonvlm/model/handsonvlm_arch.py:8   HandsOnVLMMetaForCausalLM
  VisualToTokenHelper           hoi_forecast/model/visual_to_tokens.py  VisualToTokenHelper
  HandsOnVLMForCausalLM.prepare_inputs_labels_for_multimodal
                                handsonvlm/.../handsonvlm.py:212-451    HandsOnVLMMetaForCausalLM.prepare_inputs_...
  inline <hand_traj> gather     handsonvlm/.../handsonvlm.py:146-187    gather_hand_traj_states

A host model mixes these in exactly like the reference does and must provide ``get_model()`` (object with
``vision_tower`` / ``get_vision_tower()``, ``mm_projector`` (nn.Linear) and ``embed_tokens`` (nn.Embedding)) and
``config``.  All compute goes through ``ops`` (sm_100a kernels); nothing here falls back to torch eager.

Order of operations: mean and Linear commute, so the video path pools the 1024-d ViT features FIRST and projects
the 356 pooled rows (reference: projects 25 600 rows, then pools) -- identical up to fp32 rounding
(SURVEY.md section 8a notes), 72x fewer projector FLOPs and a 4x smaller pooling read.  ``encode_images`` still
returns every projected token, as its public signature requires.
"""
from __future__ import annotations

import os
from abc import ABC, abstractmethod
from typing import Optional

import torch

from . import _lib as L
from . import ops
from .constants import HAND_TRAJ_TOKEN_ID, IGNORE_INDEX, IMAGE_TOKEN_INDEX
from .tower import CLIPVisionTower

_POOLED = ("temporal_spatial_pool", "spatial_pool", "temporal", "spatial", "temporal_spatial")


def _project(projector, x2d: torch.Tensor) -> torch.Tensor:
    """mm_projector (nn.Linear 1024 -> D, with bias) on the tcgen05 GEMM; output in the projector's dtype."""
    pre = getattr(projector, "_hvlm_pre_forward", None)
    if pre is not None:
        pre()          # e.g. dist.ProjectorGradReducer.wait: the ViT in front of this call did not need the projector
    w, b = projector.weight, projector.bias
    out_f32 = w.dtype == torch.float32
    if torch.is_grad_enabled() and (w.requires_grad or b.requires_grad):
        w16, b32 = w.to(torch.bfloat16), b.to(torch.float32)         # conversions stay in the autograd graph
    else:
        # inference: convert the parameters once per version instead of once per call
        key = (w.data_ptr(), w._version, w.dtype, b.data_ptr(), b._version, b.dtype)
        cached = getattr(projector, "_hvlm_gemm_params", None)
        if cached is None or cached[0] != key:
            cached = (key, w.detach().to(torch.bfloat16), b.detach().to(torch.float32))
            projector._hvlm_gemm_params = cached
        w16, b32 = cached[1], cached[2]
    y = ops.linear(x2d, w16, b32, out_f32)
    return y if y.dtype == w.dtype else y.to(w.dtype)


class VisualTokenCache:
    """SURVEY.md section 8(f) item 1.  The reference re-runs the whole visual path for EVERY generated token
    (``generate(use_cache=False)``: handsonvlm_inference.py:99-109 -> handsonvlm.py:555-564, and HandsOnVLM dropped the
    ``T == 1`` early-out LLaVA has at llava_arch.py:117-120).  With ``config.hvlm_cache_visual_tokens = True`` the
    drop-in keeps the visual tokens of the last clip and reuses them while the SAME image tensor object, not modified in
    place since (version counter), is passed again under ``torch.no_grad()``: the ViT, pooling and projector then run once
    per sample instead of once per token.  The cache holds a strong reference to that tensor, so its storage cannot be
    freed and handed to the next sample's clip by the caching allocator while the entry is alive (an address-based key
    would then serve stale tokens).  Off by default (bit-identical either way)."""

    def __init__(self):
        self.images = None
        self.key = None
        self.value = None
        self.hits = 0

    @staticmethod
    def _key(images, extra):
        return (images._version, tuple(images.shape), images.dtype, images.device, extra)

    def get(self, images, extra):
        if torch.is_grad_enabled() or self.images is not images or self.key != self._key(images, extra):
            return None
        self.hits += 1
        return self.value

    def put(self, images, extra, value):
        if not torch.is_grad_enabled():
            self.images, self.key, self.value = images, self._key(images, extra), value

    def clear(self):
        self.images = self.key = self.value = None


def _dedup_setting(host, images):
    """``config.hvlm_dedup_frames``: False | True (dynamic: one 4-byte readback) | int k >= 1 (static: at most k distinct
    frames per clip, no host sync, violations reported by ``check_deferred_status``).  -> (enabled, capacity, deferred)"""
    v = getattr(host.config, "hvlm_dedup_frames", False)
    if v is True:
        return True, 0, None
    if not v:
        return False, 0, None
    d = _deferred(host, images.device)
    d.poll()
    return True, int(v) * images.shape[0], d


def _cached_video_tokens(host, tower, projector, images, mode):
    dedup = _dedup_setting(host, images)
    if not getattr(host.config, "hvlm_cache_visual_tokens", False):
        return video_tokens(tower, projector, images, mode, dedup)
    cache = host.__dict__.setdefault("_hvlm_visual_cache", VisualTokenCache())
    extra = (mode, projector.weight.data_ptr(), projector.weight._version, tower.weight_blob.data_ptr())
    tok = cache.get(images, extra)
    if tok is None:
        tok = video_tokens(tower, projector, images, mode, dedup)
        cache.put(images, extra, tok)
    return tok


class DeferredChecks:
    """Device-side status words that are looked at WITHOUT stalling the host: each check copies one int32 to pinned host
    memory behind the kernel that produced it and records an event; ``poll`` (called at the start of the next static
    splice / de-duplication) and ``flush`` (explicit, synchronising) raise for the completed ones.  This is how the
    sync-free modes (``hvlm_static_splice``, ``hvlm_dedup_frames=<capacity>``) report a broken contract: one step late
    instead of silently never."""

    SLOTS = 64

    def __init__(self, device):
        self.host = torch.zeros(self.SLOTS, dtype=torch.int32).pin_memory()
        self.pending = []          # (slot, event, raiser)
        self.next = 0

    def add(self, dev_word: torch.Tensor, raiser):
        if len(self.pending) >= self.SLOTS:
            self._finish(self.pending.pop(0), wait=True)
        slot = self.next
        self.next = (self.next + 1) % self.SLOTS
        self.host[slot:slot + 1].copy_(dev_word.reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.pending.append((slot, ev, raiser))

    def _finish(self, entry, wait):
        slot, ev, raiser = entry
        if wait:
            ev.synchronize()
        raiser(int(self.host[slot]))

    def poll(self):
        while self.pending and self.pending[0][1].query():
            self._finish(self.pending.pop(0), wait=False)

    def flush(self):
        while self.pending:
            self._finish(self.pending.pop(0), wait=True)


def _deferred(host, device) -> DeferredChecks:
    d = host.__dict__.get("_hvlm_deferred")
    if d is None:
        d = host.__dict__["_hvlm_deferred"] = DeferredChecks(device)
    return d


def check_deferred_status(host) -> None:
    """Synchronise and raise if any sync-free call on ``host`` since the last check broke its contract (static splice:
    not exactly one image token per sample / bad ids; static de-duplication: more distinct frames than the capacity)."""
    d = host.__dict__.get("_hvlm_deferred")
    if d is not None:
        d.flush()


def distinct_frames(flat: torch.Tensor, capacity: int = 0, deferred: Optional[DeferredChecks] = None):
    """SURVEY 8(f).2 frame de-duplication on the device (``hvlm_frame_dedup``: checksum pass, byte-wise confirmation,
    compaction -- three launches).  flat [N, ...] -> (distinct frames [U, ...], frame_map int32 [N]) with
    ``distinct[frame_map[i]]`` bit-identical to ``flat[i]``, or None when every frame is distinct.
    capacity == 0: U is read back (ONE 4-byte host sync).  capacity > 0 (a dataset contract: at most that many distinct
    frames in the batch): no host sync at all -- the tower runs on exactly ``capacity`` rows and a broken contract is
    reported through ``deferred``."""
    N = flat.shape[0]
    nbytes = flat[0].numel() * flat.element_size()
    if N < 2 or nbytes % 16 != 0:
        return None
    flat = flat.contiguous()
    if capacity > 0:
        cap = min(int(capacity), N)
        if cap == N:
            return None
        fmap, rep, n_unique = ops.frame_dedup(flat, cap)
        if deferred is not None:
            def raiser(u, cap=cap):
                if u > cap:
                    raise RuntimeError(f"hvlm_dedup_frames={cap}: the batch had {u} distinct frames (static capacity "
                                       "exceeded; its visual tokens were wrong)")
            deferred.add(n_unique, raiser)
        return ops.gather_rows(flat, rep, cap), fmap
    fmap, rep, n_unique = ops.frame_dedup(flat)
    U = int(n_unique.item())                                          # the one host sync of the dynamic mode
    if U == N:
        return None
    return ops.gather_rows(flat, rep, U), fmap


# HVLM_POOL_BEFORE_FC2=0 runs the last layer's fc2 on every token row (A/B runs)
_POOL_BEFORE_FC2 = os.environ.get("HVLM_POOL_BEFORE_FC2", "1") != "0"


def _video_tokens_pool_before_fc2(tower, projector, flat, fmap, b, t, mode):
    """The pooled video archs only need fixed averages over tokens of the tower's output, and the last thing the tower
    does is linear (fc2 + bias + residual add), so it is applied AFTER pooling:
        pool(h + f1 W2^T + b2) = pool(h) + pool(f1) W2^T + b2
    fc2 then runs on the 356 pooled rows of a clip instead of its 25 700 token rows (same idea as pool-before-projector)."""
    pm = L.POOL_MODES[mode]
    with torch.no_grad():
        hidden, f1 = tower.forward_hidden_open(flat)                # f32 [n,257,1024], bf16 [n,257,4096]
        if fmap is not None and mode in ("temporal_spatial_pool", "spatial_pool", "temporal"):
            p_h = ops.pool_slowfast_mapped(hidden, fmap, b, t, 257, 1, pm, False)
            p_f = ops.pool_slowfast_mapped(f1, fmap, b, t, 257, 1, pm, True)
        else:
            if fmap is not None:
                idx = fmap.to(torch.int64)
                hidden, f1 = hidden.index_select(0, idx), f1.index_select(0, idx)
            p_h = ops.pool_slowfast(hidden, b, t, 257, 1, pm, False)     # f32  [b,Nv,1024]
            p_f = ops.pool_slowfast(f1, b, t, 257, 1, pm, True)          # bf16 [b,Nv,4096]
        w2, b2 = tower.last_fc2()
        n_out = p_h.shape[1]
        p_h = p_h.reshape(-1, 1024)
        ops.gemm(p_f.reshape(-1, 4096), w2, b2, epilogue="residual", resid=p_h, out=p_h)   # p_h += p_f W2^T + b2
        pooled = p_h.to(torch.bfloat16)
    tok = _project(projector, pooled)
    return tok.reshape(b, n_out, -1)


def video_tokens(tower: CLIPVisionTower, projector, images: torch.Tensor, mode: str, dedup=False) -> torch.Tensor:
    """images [b,t,3,224,224] -> visual tokens [b,Nv,D] (encode -> pool -> project).  With ``dedup`` (True, or the
    (enabled, capacity, deferred) triple of ``_dedup_setting``) the tower only encodes the distinct frames of the batch and
    the pooling reads them through a frame map."""
    assert images.ndim == 5, "multiple videos per sample not supported yet"
    b, t = images.shape[:2]
    flat = images.reshape(b * t, *images.shape[2:])      # [b*t,3,224,224] float, or raw uint8 [b*t,224,224,3]
    fmap = None
    if dedup is True:
        dedup = (True, 0, None)
    if dedup and dedup[0]:
        d = distinct_frames(flat, dedup[1], dedup[2])
        if d is not None:
            flat, fmap = d
    if (mode in _POOLED and tower.select_feature == "patch" and tower.n_layers_needed >= 1 and _POOL_BEFORE_FC2
            and hasattr(tower, "forward_hidden_open")):
        return _video_tokens_pool_before_fc2(tower, projector, flat, fmap, b, t, mode)
    with torch.no_grad():
        hidden = tower.forward_hidden(flat)                         # f32 [n_distinct,257,1024]
    if mode in ("all", "none"):
        feats = tower.feature_select(hidden, torch.bfloat16)        # [n_distinct,256(+1),1024]
        if fmap is not None:
            feats = feats.index_select(0, fmap.to(torch.int64))
        tok = _project(projector, feats.reshape(-1, feats.shape[-1]))
        return tok.reshape(b, t * feats.shape[1], -1)
    if mode not in _POOLED:
        raise ValueError(f"unknown video arch {mode}")
    if tower.select_feature != "patch":
        raise AssertionError(f"tokens.shape = {(b * t, 257, projector.out_features)}")   # reference asserts 256 tokens
    if fmap is not None and mode in ("temporal_spatial_pool", "spatial_pool", "temporal"):
        pooled = ops.pool_slowfast_mapped(hidden, fmap, b, t, 257, 1, L.POOL_MODES[mode], True)
    else:
        if fmap is not None:
            hidden = hidden.index_select(0, fmap.to(torch.int64))
        pooled = ops.pool_slowfast(hidden, b, t, 257, 1, L.POOL_MODES[mode], True)   # bf16 [b,Nv,1024]
    tok = _project(projector, pooled.reshape(-1, 1024))
    return tok.reshape(b, pooled.shape[1], -1)


class VisualToTokenHelper:
    """Same constructor and ``pipeline`` contract as hoi_forecast/model/visual_to_tokens.py:7-37.
    Supported: fuse_input_mode='origin' (the released config; the other modes are ablations on pre-extracted
    lmdb features and are out of scope), video_compress_mode in {'temporal_spatial_pool','spatial_pool','none'}."""

    def __init__(self, images_raw_encode, images_mm_projector, fuse_input_mode, video_compress_mode,
                 mm_hidden_size, token_dim, cache_host=None):
        self.cache_host = cache_host          # optional: object whose config enables the visual-token cache
        self.images_raw_encode = images_raw_encode
        self.images_mm_projector = images_mm_projector
        self.fuse_input_mode = fuse_input_mode
        self.video_compress_mode = video_compress_mode
        self.mm_hidden_size = mm_hidden_size
        self.token_dim = token_dim
        self.b = self.t = self.c = self.h = self.w = None

    def pipeline(self, **kwargs):
        images = kwargs.get("images")
        if images is None:
            raise ValueError("VisualToTokenHelper.pipeline: only images=... (fuse_input_mode='origin') is supported")
        if self.fuse_input_mode != "origin":
            raise ValueError(f"Unknown fuse_input_mode: {self.fuse_input_mode}")
        if self.video_compress_mode not in ("temporal_spatial_pool", "spatial_pool", "none"):
            raise ValueError(f"unsupported video_compress_mode: {self.video_compress_mode}")
        self.b, self.t, self.c, self.h, self.w = images.shape
        if self.cache_host is not None:
            out = _cached_video_tokens(self.cache_host, self.images_raw_encode, self.images_mm_projector, images,
                                       self.video_compress_mode)
        else:
            out = video_tokens(self.images_raw_encode, self.images_mm_projector, images, self.video_compress_mode)
        n = out.shape[1]
        assert out.shape == torch.Size([self.b, n, self.token_dim]), \
            f"output_tokens.shape = {out.shape}, expected shape is {torch.Size([self.b, n, self.token_dim])}"
        if kwargs.get("_hvlm_no_mask", False):        # internal: the splice treats a missing visual mask as all-True
            return out, None
        attention_mask = torch.ones(self.b, n, dtype=torch.bool, device=out.device)
        return out, attention_mask

    def encode_images(self, images):
        assert images.shape == torch.Size([self.b, self.t, self.c, self.h, self.w]), images.shape
        out = video_tokens(self.images_raw_encode, self.images_mm_projector, images, "none")
        return out.reshape(self.b, self.t, -1, self.token_dim)

    def compress_tokens(self, tokens, attention_mask):
        mode = self.video_compress_mode
        if mode == "none":
            b, t, s, d = tokens.shape
            return tokens.reshape(b, t * s, d), attention_mask.reshape(b, t * s)
        out = ops.pool_tokens(tokens, mode)
        return out, torch.ones(out.shape[:-1], dtype=torch.bool, device=out.device)


# ------------------------------------------------------------------------------------------------
# splice
# ------------------------------------------------------------------------------------------------
class SplicePre:
    """Image-token census of a batch of ids, taken EARLY: ``hvlm_splice_info`` + an asynchronous copy to pinned host memory
    are enqueued at the top of ``prepare_inputs_labels_for_multimodal``, before the visual pipeline of the same call.  By the
    time the splice needs the numbers (output length, error checks) the host has enqueued ~170 launches behind them, so
    ``get()`` finds the copy complete and does not stall -- the reference's per-sample host syncs without draining the GPU."""

    RING = 4

    def __init__(self, host, input_ids: torch.Tensor, vocab: int):
        B = input_ids.shape[0]
        self.ids, self.version = input_ids, input_ids._version        # the census is only valid for THIS content
        self.info = ops.splice_info(input_ids, vocab)                 # device int32 [4, B]
        pool = host.__dict__.setdefault("_hvlm_pre_pool", {})
        ring = pool.setdefault(B, [[], 0])
        if len(ring[0]) < self.RING:
            ring[0].append(torch.empty(4, B, dtype=torch.int32).pin_memory())
        self.host = ring[0][ring[1] % len(ring[0])]
        ring[1] += 1
        self.host.copy_(self.info, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()
        self._lists = None

    def get(self):
        """-> (counts, last image position, <hand_traj> tokens after it, bad-id flag), python lists per sample."""
        if self._lists is None:
            self.event.synchronize()
            self._lists = tuple(self.host.tolist())
        return self._lists


def _raise_on_status(bits: int, static: bool = False):
    if static and bits & (L.PLAN_NOT_UNIFORM | L.PLAN_ERR_LEN_OVERFLOW):
        raise RuntimeError("hvlm_static_splice contract violated: every sample must carry exactly one image token "
                           "(hybrid_dataset.py:155-158); the previous batch's spliced outputs were wrong")
    if bits & L.PLAN_ERR_BAD_ID:
        raise IndexError("index out of range in self")                 # nn.Embedding's error for a bad token id
    if bits & L.PLAN_ERR_IMG_OVERFLOW:
        raise IndexError("index out of bounds: more image tokens than image features")
    if bits & L.PLAN_ERR_HAND_COUNT:
        raise AssertionError("number of <hand_traj> tokens does not match future_hands")
    if bits & L.PLAN_ERR_LEN_OVERFLOW:
        raise RuntimeError("hvlm splice: spliced sequence longer than the planned output (static splice "
                           "contract violated: a sample has more than one image token)")


def splice_tokens(host, variant: int, input_ids, attention_mask, labels, visual, visual_mask=None,
                  future_hands=None, is_evaluate: bool = False, im_start_end: bool = False, last_visual_end=None,
                  pre: Optional[SplicePre] = None):
    """Core of both prepare_inputs_labels_for_multimodal variants -> (attention_mask', embeds, labels').
    ``im_start_end``: the ``tune_mm_mlp_adapter and mm_use_im_start_end`` branch of llava_arch.py:146-161,172-181."""
    table = host.get_model().embed_tokens.weight
    B, T = input_ids.shape
    slot_offsets = None
    if isinstance(visual, (list, tuple)):
        # per-slot token blocks of different lengths (list path of images_to_tokens): row-concatenate and describe the
        # slots by their row offsets (sizes are host-known shapes: no sync)
        sizes = [int(v.shape[0]) for v in visual]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + n)
        slot_offsets = torch.tensor(offs, dtype=torch.int32, device=input_ids.device)
        visual = torch.cat(list(visual), dim=0).unsqueeze(0)            # [1, sum n, D]
        n_img, Nv, D = len(sizes), 1, visual.shape[2]
    else:
        n_img, Nv, D = visual.shape
        sizes = None
    if visual.dtype != table.dtype:
        visual = visual.to(table.dtype)
    static = bool(getattr(host.config, "hvlm_static_splice", False)) and sizes is None
    hand_mode, n_hand = 0, 0
    if variant == L.SPLICE_HANDSONVLM:
        if not is_evaluate:
            if future_hands is None:
                raise TypeError("'NoneType' object is not subscriptable")   # kwargs.get('future_hands')[batch_idx]
            assert tuple(future_hands.shape[1:]) == (2, 4, 2), future_hands.shape
            hand_mode, n_hand = 1, 4
        elif future_hands is not None:
            hand_mode, n_hand = 2, int(future_hands.shape[2])
    if static:
        # collator contract (hybrid_dataset.py:155-158): exactly one image token per sample, equal T.  No host sync at all;
        # a broken contract of an EARLIER call raises here, this call's status is looked at by the next poll
        _deferred(host, input_ids.device).poll()
        counts = ops.splice_count(input_ids)
        Lout, uniform = T - 1 + Nv, True
    else:
        # the census was taken before the visual pipeline was enqueued (prepare_inputs_labels_for_multimodal); a direct
        # call takes it now.  Everything the reference would raise is raised HERE, from host numbers, before the plan
        if pre is None or pre.ids is not input_ids or pre.version != input_ids._version:
            pre = SplicePre(host, input_ids, table.shape[0])
        ks, last_pos, hand_tail, bad = pre.get()
        counts = pre.info[0]
        if any(bad):
            raise IndexError("index out of range in self")             # nn.Embedding's error for a bad token id
        slot = 0
        for k in ks:                                                   # cur_image_idx bookkeeping
            if k > 0 and slot + k > n_img:
                raise IndexError("index out of bounds: more image tokens than image features")
            slot += max(k, 1)
        if hand_mode:
            for k, lp, nh in zip(ks, last_pos, hand_tail):
                if k > 0 and ((hand_mode == 1 and nh > 4) or (hand_mode == 2 and lp + 1 < T and nh != n_hand)):
                    raise AssertionError("number of <hand_traj> tokens does not match future_hands")
        if sizes is None:
            lens = [T + k * (Nv - 1) for k in ks]
        else:
            lens, slot = [], 0
            for k in ks:                                             # a sample without image token still uses a slot
                lens.append(T - k + sum(sizes[slot:slot + k]))
                slot += max(k, 1)
        Lout, uniform = max(lens), all(n == lens[0] for n in lens)
    src_index, hand_code, lens_d, hand_scale, status = ops.splice_plan(
        input_ids, counts, Nv, n_img, Lout, table.shape[0], variant, hand_mode, n_hand, slot_offsets, last_visual_end)
    if static:
        host._hvlm_splice_status = status
        _deferred(host, input_ids.device).add(status, lambda bits: _raise_on_status(bits, static=True))
    embeds, new_labels, new_mask = ops.splice_gather(
        src_index, hand_code, lens_d, hand_scale, input_ids, labels, attention_mask, table, visual, visual_mask,
        future_hands if hand_mode else None, variant | (L.SPLICE_FLAG_IM_START_END if im_start_end else 0))
    if labels is None:
        new_labels = None
    if attention_mask is None:
        new_mask = None
    elif variant == L.SPLICE_HANDSONVLM and not uniform:
        # bug-compatible with handsonvlm.py:441: a ragged batch pads the MASK with IGNORE_INDEX in the labels' dtype
        ar = torch.arange(Lout, device=new_mask.device).unsqueeze(0)
        new_mask = torch.where(ar < lens_d.unsqueeze(1), new_mask.to(torch.int64),
                               torch.full((), IGNORE_INDEX, dtype=torch.int64, device=new_mask.device))
    elif attention_mask.dtype != torch.bool:
        new_mask = new_mask.to(attention_mask.dtype)
    return new_mask, embeds, new_labels


def _early_census(host, input_ids):
    """The splice's image-token census, launched before the visual pipeline (None in static mode: no readback at all)."""
    if bool(getattr(host.config, "hvlm_static_splice", False)) or not input_ids.is_cuda:
        return None
    return SplicePre(host, input_ids, host.get_model().embed_tokens.weight.shape[0])


class LlavaMetaForCausalLM(ABC):
    """llava/model/llava_arch.py:73-234"""

    @abstractmethod
    def get_model(self):
        pass

    def get_vision_tower(self):
        return self.get_model().get_vision_tower()

    def encode_images(self, images):
        """images [N,3,224,224] -> [N,256,D]  (llava_arch.py:81-93)"""
        tower = self.get_model().get_vision_tower()
        with torch.no_grad():
            hidden = tower.forward_hidden(images)
        feats = tower.feature_select(hidden, torch.bfloat16)
        tok = _project(self.get_model().mm_projector, feats.reshape(-1, feats.shape[-1]))
        return tok.reshape(feats.shape[0], feats.shape[1], -1)

    def images_to_tokens(self, images):
        """llava_arch.py:95-106.  A plain batch [N,3,H,W] -> [N,256,D].  A list (or a 5-d stack) of per-sample image
        groups is encoded in ONE tower call and handed back as one flat [n_i*256, D] token block per sample."""
        grouped = isinstance(images, (list, tuple)) or images.ndim == 5
        if not grouped:
            return self.encode_images(images)
        groups = list(images)
        feats = self.encode_images(torch.cat(groups, dim=0))                 # [sum n_i, 256, D]
        out, start = [], 0
        for g in groups:
            n = g.shape[0]
            out.append(feats[start:start + n].reshape(n * feats.shape[1], feats.shape[2]))
            start += n
        return out

    def visual_to_tokens(self, images):
        return self.images_to_tokens(images)

    def prepare_inputs_labels_for_multimodal(self, input_ids, attention_mask, past_key_values, labels, images):
        """llava_arch.py:110-234, including the ``tune_mm_mlp_adapter and mm_use_im_start_end`` branch (:146-161)."""
        vision_tower = self.get_vision_tower()
        if vision_tower is None or images is None or input_ids.shape[1] == 1:
            if past_key_values is not None and vision_tower is not None and images is not None and input_ids.shape[1] == 1:
                attention_mask = torch.ones((attention_mask.shape[0], past_key_values[-1][-1].shape[-2] + 1),
                                            dtype=attention_mask.dtype, device=attention_mask.device)
            return input_ids, attention_mask, past_key_values, None, labels
        ise = bool(getattr(self.config, "tune_mm_mlp_adapter", False) and getattr(self.config, "mm_use_im_start_end", False))
        pre = _early_census(self, input_ids)
        image_features = self.visual_to_tokens(images)
        if isinstance(image_features, (list, tuple)):
            if all(x.shape == image_features[0].shape for x in image_features):
                image_features = torch.stack(list(image_features), 0)
            # else: token blocks of different lengths go to the splice as a list (ragged plan)
        new_mask, embeds, new_labels = splice_tokens(self, L.SPLICE_LLAVA, input_ids, attention_mask, labels,
                                                     image_features, im_start_end=ise, pre=pre)
        return None, new_mask, past_key_values, embeds, new_labels


class LitaMetaForCausalLM(LlavaMetaForCausalLM):
    """lita/model/lita_arch.py:17-85"""

    def videos_to_tokens(self, images):
        assert images.ndim == 5, "multiple videos per sample not supported yet"
        video_arch = getattr(self.config, "video_arch", "temporal")
        return _cached_video_tokens(self, self.get_model().get_vision_tower(), self.get_model().mm_projector, images,
                                    video_arch)

    def visual_to_tokens(self, images):
        input_type = getattr(self.config, "input_type", "image")
        if input_type == "image":
            return self.images_to_tokens(images)
        elif input_type == "video":
            return self.videos_to_tokens(images)


def gather_hand_traj_states(hidden_states: torch.Tensor, labels: torch.Tensor, hand_token_id: int = HAND_TRAJ_TOKEN_ID,
                            future_valid: Optional[torch.Tensor] = None, strict: bool = True):
    """handsonvlm.py:146-187 as one kernel: returns (pred_hand_embeddings [B,2,4,D/2], valid [B] bool).
    ``future_valid`` ([B,2] bool) is cleared in place for samples without hand tokens, like the reference.
    strict=True reproduces the reference's failure when a sample has a hand-token count other than 0 or 4
    (costs one 4-byte-per-sample readback); strict=False leaves the check to the caller (no host sync)."""
    out, valid, rows, counts = ops.hand_gather(hidden_states, labels, hand_token_id)
    if strict:
        D = hidden_states.shape[-1]
        for c in counts.tolist():
            if c not in (0, 4):
                raise RuntimeError(f"shape '[4, {D // 2}, 2]' is invalid for input of size {c * D}")
    if future_valid is not None:
        future_valid.logical_and_(valid.unsqueeze(-1))
    return out, valid


class HandsOnVLMMetaForCausalLM(LitaMetaForCausalLM):
    """handsonvlm/model/handsonvlm_arch.py:8-17 + the released model's splice (handsonvlm.py:212-451)."""

    def visual_to_tokens(self, images):
        input_type = getattr(self.config, "input_type", "image")
        if input_type == "image":
            return self.images_to_tokens(images)
        elif input_type == "video":
            return self.videos_to_tokens(images)

    def prepare_inputs_labels_for_multimodal(self, input_ids, attention_mask, past_key_values, labels, images, **kwargs):
        pre = _early_census(self, input_ids)
        helper = VisualToTokenHelper(images_raw_encode=self.get_model().get_vision_tower(),
                                     images_mm_projector=self.get_model().mm_projector,
                                     fuse_input_mode=self.config.fuse_input_mode,
                                     video_compress_mode=self.config.video_compress_mode,
                                     mm_hidden_size=self.config.mm_hidden_size, token_dim=self.token_dim,
                                     cache_host=self)
        visual_tokens, visual_mask = helper.pipeline(images=images, _hvlm_no_mask=True, **kwargs)
        assert visual_tokens.shape == torch.Size([self.B, visual_tokens.shape[1], self.token_dim]), visual_tokens.shape
        if getattr(self.config, "tune_mm_mlp_adapter", False) and getattr(self.config, "mm_use_im_start_end", False):
            # handsonvlm.py:263-286,343-344: <im_start>/<im_end> branch -- no hand embeddings on the tail, the <im_end>
            # slot takes the image position's label and mask entry, last_visual_token_index is left untouched
            new_mask, embeds, new_labels = splice_tokens(
                self, L.SPLICE_HANDSONVLM, input_ids, attention_mask, labels, visual_tokens, None, None, True,
                im_start_end=True, pre=pre)
            return None, new_mask, past_key_values, embeds, new_labels
        # side effect of the reference (handsonvlm.py:288): a 0-d int64 tensor, written by the plan kernel for the last
        # image token of the last sample that has one (left at its previous value when no sample has an image token)
        lve = self.__dict__.get("last_visual_token_index")
        if not (isinstance(lve, torch.Tensor) and lve.dtype == torch.int64 and lve.dim() == 0
                and lve.device == input_ids.device):
            lve = torch.full((), -1, dtype=torch.int64, device=input_ids.device)
        new_mask, embeds, new_labels = splice_tokens(
            self, L.SPLICE_HANDSONVLM, input_ids, attention_mask, labels, visual_tokens, None,
            kwargs.get("future_hands"), kwargs.get("is_evaluate", False), last_visual_end=lve, pre=pre)
        self.__dict__["last_visual_token_index"] = lve
        return None, new_mask, past_key_values, embeds, new_labels

    def clear_visual_token_cache(self):
        cache = self.__dict__.get("_hvlm_visual_cache")
        if cache is not None:
            cache.clear()

    def gather_hand_traj_states(self, hidden_states, labels, future_valid=None, strict=True):
        return gather_hand_traj_states(hidden_states, labels, HAND_TRAJ_TOKEN_ID, future_valid, strict)
