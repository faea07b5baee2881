"""TEST INFRASTRUCTURE ONLY -- fp32 torch-CPU restatement of the HandsOnVLM visual-token path.

Every function restates one stage of the reference and cites the lines it follows
(paths relative to the reference checkout).  It is written independently of the
reference's control flow (index plans instead of Python cat-loops) so that it can be compared
both against the reference itself (``tests/golden`` fixtures made by ``make_golden.py``) and
against the CUDA kernels, which use the same index-plan formulation.

The ViT arithmetic lives in a third-party dependency that is not vendored in the reference:
HuggingFace ``transformers==4.31.0`` (reference ``setup.py:16``),
``models/clip/modeling_clip.py`` (``CLIPVisionModel``).  ``vit_hidden`` restates its published
forward; it is pinned against the installed transformers' CLIPVisionModel through fixtures.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from .synth import HAND_TRAJ_TOKEN_ID, IGNORE_INDEX, IMAGE_TOKEN_INDEX, VIT_L14, VitCfg


# ----------------------------------------------------------------------------------------
# a1: CLIPVisionTower.forward  (llava/model/multimodal_encoder/clip_encoder.py:39-51)
# ----------------------------------------------------------------------------------------

def _r(x: torch.Tensor, emulate: str | None) -> torch.Tensor:
    """Round a GEMM operand to bf16 (debug regime that mirrors the CUDA path's operand
    precision); identity in the fp32 oracle regime."""
    if emulate in ("bf16", "bf16_fold"):
        return x.to(torch.bfloat16).to(torch.float32)
    return x


def folded_layernorm_linear(h: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, W: torch.Tensor, b: torch.Tensor,
                            eps: float = 1e-5, shift: torch.Tensor | None = None) -> torch.Tensor:
    """``LayerNorm(h) W^T + b`` in the arithmetic of the CUDA path's FOLDED LayerNorm (debug regime "bf16_fold"):
    the GEMM reads the un-normalised rows and the gamma-scaled weights as bf16 and the row statistics are applied to the
    product,  rstd * (bf16(h) bf16(gamma*W)^T - mean * c) + (b + W beta)  with c = row sums of the rounded weights and
    one-pass fp32 statistics.  Algebraically identical to LayerNorm followed by Linear (modeling_clip.py CLIPEncoderLayer).
    ``shift`` [..., 1]: the per-row constant the producer subtracts before rounding / accumulating (LayerNorm does not see
    it); the CUDA path uses the row's running mean."""
    if shift is not None:
        h = h - shift
    E = h.shape[-1]
    mean = h.sum(-1, keepdim=True) / E
    var = ((h * h).sum(-1, keepdim=True) / E - mean * mean).clamp_min(0.0)
    rstd = torch.rsqrt(var + eps)
    w_f = _r(W * gamma.unsqueeze(0), "bf16")
    c = w_f.double().sum(-1).float()
    b_f = (b.double() + W.double() @ beta.double()).float()
    return rstd * (_r(h, "bf16") @ w_f.t() - mean * c) + b_f


def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    """HF ``QuickGELUActivation``: x * sigmoid(1.702 x) (CLIP config hidden_act='quick_gelu')."""
    return x * torch.sigmoid(1.702 * x)


def patch_matrix(pixels: torch.Tensor, patch: int = 14) -> torch.Tensor:
    """[N,3,H,W] -> [N, (H/p)*(W/p), 3*p*p]; column order (c, i, j) matches
    ``Conv2d.weight.reshape(E, -1)`` so conv == matmul (CLIPVisionEmbeddings.patch_embedding)."""
    N, C, H, W = pixels.shape
    g = H // patch
    x = pixels.reshape(N, C, g, patch, W // patch, patch)
    return x.permute(0, 2, 4, 1, 3, 5).reshape(N, g * (W // patch), C * patch * patch)


def vit_hidden(pixels: torch.Tensor, sd: dict, n_layers_run: int, cfg: VitCfg = VIT_L14,
               emulate: str | None = None, return_all: bool = False):
    """Residual stream after ``n_layers_run`` encoder layers == HF ``hidden_states[n_layers_run]``
    (index 0 is the post-``pre_layrnorm`` embedding).  [N, 1+P, E] fp32.

    Follows HF CLIPVisionTransformer.forward: embeddings (class token ++ conv patches + position
    embedding) -> pre_layrnorm -> encoder layers (pre-LN attention + pre-LN quick-GELU MLP)."""
    p = "vision_model."
    E, H = cfg.hidden, cfg.heads
    dh = E // H
    N = pixels.shape[0]
    pm = patch_matrix(pixels.float(), cfg.patch)                                   # [N,P,588]
    w = sd[p + "embeddings.patch_embedding.weight"].reshape(E, -1)
    x = _r(pm, emulate) @ _r(w, emulate).t()                                       # [N,P,E]
    cls = sd[p + "embeddings.class_embedding"].reshape(1, 1, E).expand(N, 1, E)
    x = torch.cat([cls, x], dim=1) + sd[p + "embeddings.position_embedding.weight"].unsqueeze(0)
    h = F.layer_norm(x, (E,), sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"], 1e-5)
    hs = [h]
    fold = emulate == "bf16_fold"
    shift = h.mean(-1, keepdim=True)      # "bf16_fold": the running row mean (pre_layrnorm writes the exact one)
    for l in range(n_layers_run):
        q_ = f"{p}encoder.layers.{l}."
        g1, b1 = sd[q_ + "layer_norm1.weight"], sd[q_ + "layer_norm1.bias"]
        y = None if fold else _r(F.layer_norm(h, (E,), g1, b1, 1e-5), emulate)

        def lin(t, nm):
            return t @ _r(sd[f"{q_}{nm}.weight"], emulate).t() + sd[f"{q_}{nm}.bias"]

        def ln_lin(t, y_, g, b, nm, scale=1.0):      # LayerNorm + Linear, in the regime's arithmetic
            if fold:
                return folded_layernorm_linear(t, g, b, sd[f"{q_}{nm}.weight"] * scale, sd[f"{q_}{nm}.bias"] * scale,
                                               shift=shift)
            return lin(y_, nm) * scale

        S = h.shape[1]
        qh = _r(ln_lin(h, y, g1, b1, "self_attn.q_proj", dh ** -0.5), emulate).reshape(N, S, H, dh).transpose(1, 2)
        kh = _r(ln_lin(h, y, g1, b1, "self_attn.k_proj"), emulate).reshape(N, S, H, dh).transpose(1, 2)
        vh = _r(ln_lin(h, y, g1, b1, "self_attn.v_proj"), emulate).reshape(N, S, H, dh).transpose(1, 2)
        sc = qh @ kh.transpose(-1, -2)
        if emulate is not None:
            # the kernel normalises after the PV product: P~ = exp(s - max) in bf16, / rowsum(fp32)
            m = sc.max(-1, keepdim=True).values
            pe = torch.exp(sc - m)
            a = (_r(pe, emulate) @ vh) / pe.sum(-1, keepdim=True)
        else:
            a = torch.softmax(sc, dim=-1) @ vh
        a = _r(a.transpose(1, 2).reshape(N, S, E), emulate)
        shift = shift + (h - shift).mean(-1, keepdim=True)        # the QKV GEMM's epilogue updates the running mean ...
        h = h + lin(a, "self_attn.out_proj")
        g2, b2 = sd[q_ + "layer_norm2.weight"], sd[q_ + "layer_norm2.bias"]
        y = None if fold else _r(F.layer_norm(h, (E,), g2, b2, 1e-5), emulate)
        f = _r(quick_gelu(ln_lin(h, y, g2, b2, "mlp.fc1")), emulate)
        shift = shift + (h - shift).mean(-1, keepdim=True)        # ... and so does fc1's
        h = h + lin(f, "mlp.fc2")
        hs.append(h)
    return hs if return_all else h


def tower_forward(pixels: torch.Tensor, sd: dict, select_layer: int = -2, cfg: VitCfg = VIT_L14,
                  select_feature: str = "patch", emulate: str | None = None) -> torch.Tensor:
    """``CLIPVisionTower.forward`` + ``feature_select`` (clip_encoder.py:29-51): take
    ``hidden_states[select_layer]`` of the (cfg.layers)-deep tower, drop CLS for 'patch', cast
    back to ``images.dtype``."""
    n_run = select_layer if select_layer >= 0 else cfg.layers + 1 + select_layer
    h = vit_hidden(pixels, sd, n_run, cfg, emulate)
    if select_feature == "patch":
        h = h[:, 1:]
    elif select_feature != "cls_patch":
        raise ValueError(f"Unexpected select feature: {select_feature}")
    return h.to(pixels.dtype)


# ----------------------------------------------------------------------------------------
# a2: encode_images  (llava/model/llava_arch.py:81-93; visual_to_tokens.py:274-284)
# ----------------------------------------------------------------------------------------

def project(x: torch.Tensor, W: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """``mm_projector = nn.Linear(1024, D)`` with bias (llava_arch.py:33,92)."""
    return x @ W.t() + b


# ----------------------------------------------------------------------------------------
# a3 / a4: slow-fast pooling  (lita/model/lita_arch.py:41-77; visual_to_tokens.py:230-272)
# ----------------------------------------------------------------------------------------

def selected_frames(t: int, n: int = 4) -> np.ndarray:
    """``np.round(np.linspace(0, t-1, 4)).astype(int)`` (lita_arch.py:56, visual_to_tokens.py:254)."""
    return np.round(np.linspace(0, t - 1, n)).astype(int)


def pool_tokens(tokens: torch.Tensor, mode: str) -> torch.Tensor:
    """tokens [b,t,256,d] -> pooled tokens.  Index-formula restatement:
    fast[b,f,:]          = mean_s tokens[b,f,s,:]                         (lita_arch.py:67)
    slow[b,64k+8h+w,:]   = mean_{i,j in {0,1}} tokens[b,sel[k],(2h+i)*16+(2w+j),:]   (:56-63)
    out = cat([fast, slow], 1) for 'temporal_spatial_pool' (:69)."""
    b, t, s, d = tokens.shape
    if mode in ("all", "none"):
        return tokens.reshape(b, t * s, d)
    if mode == "temporal":
        return tokens.mean(dim=2)
    if mode == "spatial":
        return tokens.mean(dim=1)
    if mode == "temporal_spatial":
        return torch.cat([tokens.mean(dim=2), tokens.mean(dim=1)], dim=1)
    if mode in ("temporal_spatial_pool", "spatial_pool"):
        assert s == 256
        sel = torch.as_tensor(selected_frames(t), dtype=torch.long)
        g = tokens[:, sel].reshape(b, 4, 8, 2, 8, 2, d)          # [b,k,h',i,w',j,d]
        slow = g.mean(dim=(3, 5)).reshape(b, 256, d)
        if mode == "spatial_pool":
            return slow
        return torch.cat([tokens.mean(dim=2), slow], dim=1)
    raise ValueError(f"unknown video arch {mode}")


def pool_tokens_backward(dout: torch.Tensor, t: int, mode: str = "temporal_spatial_pool") -> torch.Tensor:
    """d_tok[b,f,s,:] = d_fast[b,f,:]/256 + sum_{k: sel[k]==f} d_slow[b,64k+8(h//2)+(w//2),:]/4
    (autograd of pool_tokens).  Duplicated ``sel`` entries (t<4) accumulate."""
    b, n, d = dout.shape
    dtok = torch.zeros(b, t, 256, d, dtype=dout.dtype)
    if mode == "temporal_spatial_pool":
        dtok += dout[:, :t].unsqueeze(2) / 256.0
        ds = dout[:, t:]
    elif mode == "spatial_pool":
        ds = dout
    else:
        raise ValueError(mode)
    sel = selected_frames(t)
    ds = ds.reshape(b, 4, 8, 1, 8, 1, d).expand(b, 4, 8, 2, 8, 2, d).reshape(b, 4, 256, d) / 4.0
    for k, f in enumerate(sel):
        dtok[:, int(f)] += ds[:, k]
    return dtok


# ----------------------------------------------------------------------------------------
# a6: hand positional embedding  (handsonvlm.py:310-338)
# ----------------------------------------------------------------------------------------

def hand_pos_embedding(gt_hand: torch.Tensor, D: int) -> torch.Tensor:
    """gt_hand [2,n,2] -> [n,D]; out[k, 2c+h] = enc(hand h, point k)[c] with
    enc = cat[sin(x f), cos(y f), sin(x f), cos(y f)], f = 10000^(-arange(0,D/4,2)/(D/4))."""
    ch = D // 4
    n = gt_hand.shape[1]
    inv_freq = 1.0 / (10000 ** (torch.arange(0, ch, 2, dtype=gt_hand.dtype) / ch))
    flat = gt_hand.reshape(-1, 2)
    xe = flat[:, 0:1] * inv_freq
    ye = flat[:, 1:2] * inv_freq
    enc = torch.cat([torch.sin(xe), torch.cos(ye), torch.sin(xe), torch.cos(ye)], dim=-1)   # [2n, D/2]
    return enc.reshape(2, n, D // 2).permute(1, 2, 0).reshape(n, D)


# ----------------------------------------------------------------------------------------
# a5 / a6: splice plans
# ----------------------------------------------------------------------------------------

def splice_plan(ids_row: torch.Tensor, Nv):
    """Source plan for one sample: list of (kind, index) per output row.
    kind 0 = text row (index = position in ids_row), kind 1 = visual row (index = j-th image
    token of this sample * Nv + offset; with a list ``Nv`` of per-image row counts the index is the pair
    (j, offset)).  A sample with no IMAGE_TOKEN_INDEX maps 1:1."""
    out = []
    j = 0
    for pos, tok in enumerate(ids_row.tolist()):
        if tok == IMAGE_TOKEN_INDEX:
            if isinstance(Nv, int):
                out.extend((1, j * Nv + r) for r in range(Nv))
            else:
                out.extend((1, (j, r)) for r in range(Nv[j]))
            j += 1
        else:
            out.append((0, pos))
    return out, j


def splice(ids, attention_mask, labels, visual, embed_w, variant: str,
           visual_mask=None, future_hands=None, is_evaluate: bool = False, im_start_end: bool = False):
    """Restates both splices:

    variant 'llava'      -- ``LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal``
                            (llava/model/llava_arch.py:110-234; released flags: no im_start_end)
    variant 'handsonvlm' -- ``HandsOnVLMForCausalLM.prepare_inputs_labels_for_multimodal``
                            (handsonvlm/model/language_model/handsonvlm.py:212-451)

    im_start_end         -- the ``tune_mm_mlp_adapter and mm_use_im_start_end`` branch
                            (llava_arch.py:146-161,172-173,180-181): same embedding rows; the token right after an
                            image token takes the label of the image-token position (:159); text embeddings other than
                            the two tokens around an image are detached (see ``splice_backward``).

    ids [B,T] i64, attention_mask [B,T] bool|None, labels [B,T] i64|None,
    visual [n_img, Nv, D], embed_w [V, D].  Returns (attention_mask', embeds, labels').

    Image-slot bookkeeping (``cur_image_idx``): a sample with k image tokens consumes k
    consecutive slots; a sample with none still consumes one (llava_arch.py:135, handsonvlm.py:243).
    """
    B, T = ids.shape
    ragged = isinstance(visual, (list, tuple))          # per-slot blocks [n_g, D] (list path of images_to_tokens)
    Nv, D = (None, visual[0].shape[-1]) if ragged else (visual.shape[1], visual.shape[2])
    rows_e, rows_l, rows_m = [], [], []
    slot = 0
    for b in range(B):
        if ragged:
            k_here = int((ids[b] == IMAGE_TOKEN_INDEX).sum())
            plan, k_img = splice_plan(ids[b], [int(visual[slot + j].shape[0]) for j in range(k_here)])
        else:
            plan, k_img = splice_plan(ids[b], Nv)
        e = torch.empty(len(plan), D, dtype=embed_w.dtype)
        lab = torch.empty(len(plan), dtype=torch.int64) if labels is not None else None
        msk = torch.empty(len(plan), dtype=torch.bool) if attention_mask is not None else None
        for r, (kind, idx) in enumerate(plan):
            if kind == 0:
                e[r] = embed_w[ids[b, idx]]
                after_img = im_start_end and idx > 0 and int(ids[b, idx - 1]) == IMAGE_TOKEN_INDEX
                if lab is not None:
                    lab[r] = labels[b, idx - 1 if after_img else idx]
                if msk is not None:
                    # only the HandsOnVLM splice builds its mask by position (handsonvlm.py:283)
                    msk[r] = attention_mask[b, idx - 1 if (after_img and variant == "handsonvlm") else idx]
            else:
                img, off = (slot + idx[0], idx[1]) if ragged else (slot + idx // Nv, idx % Nv)
                e[r] = visual[img][off]
                if lab is not None:
                    lab[r] = IGNORE_INDEX
                if msk is not None:
                    msk[r] = True if visual_mask is None else visual_mask[img][off]
        if variant == "handsonvlm" and k_img > 0 and not im_start_end:     # (:343-344: that branch adds nothing)
            # tail segment = text after the last image token (handsonvlm.py:342-396)
            last_img_pos = int(torch.where(ids[b] == IMAGE_TOKEN_INDEX)[0][-1])
            tail = ids[b, last_img_pos + 1:]
            if tail.numel() > 0:
                tail_row0 = len(plan) - tail.numel()
                hand_rel = torch.where(tail == HAND_TRAJ_TOKEN_ID)[0]
                cnt = int(hand_rel.numel())
                if not is_evaluate:
                    assert future_hands is not None and tuple(future_hands[b].shape) == (2, 4, 2)
                    assert cnt <= 4
                    emb = hand_pos_embedding(future_hands[b].to(torch.float32), D) * (cnt / 4)
                    idx = hand_rel.tolist() + [0] * (4 - cnt)
                    # zero.scatter(0, idx, emb): duplicate indices -> last writer wins on CPU
                    add = {}
                    for k, rr in enumerate(idx):
                        add[rr] = emb[k]
                    for rr, v in add.items():
                        e[tail_row0 + rr] = (e[tail_row0 + rr].float() + v).to(e.dtype)
                elif future_hands is not None:
                    gh = future_hands[b]
                    assert gh.shape[1] == cnt, f"gt_hand_num: {gh.shape[1]}, hand_token_cnt: {cnt}"
                    emb = hand_pos_embedding(gh.to(torch.float32), D)
                    for k, rr in enumerate(hand_rel.tolist()):
                        e[tail_row0 + rr] = (e[tail_row0 + rr].float() + emb[k]).to(e.dtype)
        slot += max(k_img, 1)
        rows_e.append(e)
        rows_l.append(lab)
        rows_m.append(msk)

    lens = [x.shape[0] for x in rows_e]
    L = max(lens)
    ragged = any(n != lens[0] for n in lens)
    embeds = torch.zeros(B, L, D, dtype=embed_w.dtype)
    for b in range(B):
        embeds[b, : lens[b]] = rows_e[b]
    new_labels = None
    if labels is not None:
        new_labels = torch.full((B, L), IGNORE_INDEX, dtype=torch.int64)
        for b in range(B):
            new_labels[b, : lens[b]] = rows_l[b]
    new_mask = None
    if attention_mask is not None:
        if variant == "handsonvlm":
            if ragged:
                # handsonvlm.py:441 pads the *mask* with IGNORE_INDEX in the labels' dtype
                new_mask = torch.full((B, L), IGNORE_INDEX, dtype=torch.int64)
                for b in range(B):
                    new_mask[b, : lens[b]] = rows_m[b].to(torch.int64)
            else:
                new_mask = torch.stack(rows_m, 0)
        else:
            # llava_arch.py:215-232: the mask is NOT position-spliced: True x (len-T) on the left,
            # the original mask, False right-pad.
            new_mask = torch.zeros(B, L, dtype=torch.bool)
            for b in range(B):
                new_mask[b, : lens[b] - T] = True
                new_mask[b, lens[b] - T: lens[b]] = attention_mask[b]
    return new_mask, embeds, new_labels


def last_visual_token_index(ids: torch.Tensor, Nv: int, previous: int = -1) -> int:
    """handsonvlm.py:288 side effect: ``self.last_visual_token_index = image_token_start + n_visual_tokens`` is assigned for
    every image token of every sample in order; ``image_token_start`` indexes ``cur_input_ids`` AFTER the slices of
    handsonvlm.py:300-303 (everything up to and including the previous image token cut off).  Samples without an image
    token leave it alone (:234-245)."""
    val = previous
    for row in ids:
        pos = (row == IMAGE_TOKEN_INDEX).nonzero().flatten().tolist()
        prev = -1
        for p in pos:
            val = (p - prev - 1) + Nv
            prev = p
    return val


def splice_backward(d_embeds, ids, Nv: int, n_img: int, vocab: int, im_start_end: bool = False):
    """Backward of the copy part of the splice: visual rows -> d_visual [n_img,Nv,D];
    text rows scatter-add into d_embed_table [vocab, D].  With ``im_start_end`` only the tokens
    directly before / after an image token (<im_start>, <im_end>) reach the table: every other text segment is
    ``.detach()``ed by the reference (llava_arch.py:150,181)."""
    B, L, D = d_embeds.shape
    d_visual = torch.zeros(n_img, Nv, D, dtype=torch.float32)
    d_table = torch.zeros(vocab, D, dtype=torch.float32)
    slot = 0
    for b in range(B):
        plan, k_img = splice_plan(ids[b], Nv)
        for r, (kind, idx) in enumerate(plan):
            if kind == 0:
                T = ids.shape[1]
                near = (idx + 1 < T and int(ids[b, idx + 1]) == IMAGE_TOKEN_INDEX) or \
                       (idx > 0 and int(ids[b, idx - 1]) == IMAGE_TOKEN_INDEX)
                if not im_start_end or near:
                    d_table[ids[b, idx]] += d_embeds[b, r].float()
            else:
                d_visual[slot + idx // Nv, idx % Nv] += d_embeds[b, r].float()
        slot += max(k_img, 1)
    return d_visual, d_table


# ----------------------------------------------------------------------------------------
# a7 / a8: <hand_traj> hidden-state gather  (handsonvlm.py:146-187, 609-622)
# ----------------------------------------------------------------------------------------

def gather_hand_traj(hidden: torch.Tensor, labels: torch.Tensor, hand_id: int = HAND_TRAJ_TOKEN_ID):
    """Rows that *predict* each <hand_traj> label (mask shifted left by one).  Returns
    (out [B,2,4,D/2] with out[b,h,k,j] = hidden[b,row_k,2j+h], valid [B] bool, rows [B,4] i32 (-1 = none))."""
    B, L, D = hidden.shape
    out = torch.zeros(B, 2, 4, D // 2, dtype=hidden.dtype)
    valid = torch.zeros(B, dtype=torch.bool)
    rows = torch.full((B, 4), -1, dtype=torch.int32)
    for b in range(B):
        m = labels[b] == hand_id
        pos = torch.where(m[1:])[0]                    # m_shift[i] = m[i+1]
        if pos.numel() == 0:
            continue
        if pos.numel() != 4:
            raise RuntimeError(f"shape '[4, {D // 2}, 2]' is invalid for input of size {pos.numel() * D}")
        valid[b] = True
        rows[b] = pos.to(torch.int32)
        out[b] = hidden[b, pos].reshape(4, D // 2, 2).permute(2, 0, 1)
    return out, valid, rows


def gather_hand_traj_backward(dout: torch.Tensor, rows: torch.Tensor, L: int) -> torch.Tensor:
    B, _, _, half = dout.shape
    dh = torch.zeros(B, L, 2 * half, dtype=torch.float32)
    for b in range(B):
        for k in range(4):
            r = int(rows[b, k])
            if r >= 0:
                dh[b, r] += dout[b, :, k, :].float().t().reshape(-1)   # [half,2] -> interleave
    return dh


def gather_hand_traj_step(hidden_last: torch.Tensor) -> torch.Tensor:
    """Generation-time gather (handsonvlm.py:613-616): hidden[:, -1, :] [B,D] ->
    reshape(B, D/2, 2).permute(0,2,1).unsqueeze(2) -> [B,2,1,D/2]."""
    B, D = hidden_last.shape
    return hidden_last.reshape(B, D // 2, 2).permute(0, 2, 1).unsqueeze(2)


# ----------------------------------------------------------------------------------------
# trajectory head after the gather, generation side
# ----------------------------------------------------------------------------------------

def traj_cvae_inference(cond: torch.Tensor, z: torch.Tensor, sd: dict) -> torch.Tensor:
    """``TrajCVAE.inference`` with ``condition_contact=False`` (hoi_forecast/architecture/traj_decoder.py:75-91) ->
    ``VAE.inference`` (decoder_modules.py:56-60): ``dec_MLP(cat(z, cond))`` with
    ``dec_MLP = Linear(latent + token_dim, hidden), ELU, Linear(hidden, 2)`` (decoder_modules.py:26-29).
    cond [R, token_dim], z [R, latent] -- the ALREADY SCALED noise (the reference draws
    ``z_scale * torch.randn([R, latent])`` itself, traj_decoder.py:87) -> [R, 2].  fp32 throughout."""
    pre = "hand_traj_decoder.cvae.dec_MLP."
    x = torch.cat([z.float(), cond.float()], dim=-1)
    h = x @ sd[pre + "0.weight"].float().t() + sd[pre + "0.bias"].float()
    h = torch.where(h > 0, h, torch.expm1(h))                                   # nn.ELU(alpha=1)
    return h @ sd[pre + "2.weight"].float().t() + sd[pre + "2.bias"].float()


def traj_decoder_inference(pred_hand_embeddings: torch.Tensor, z: torch.Tensor, sd: dict) -> torch.Tensor:
    """``TrajDecoder.inference`` (handsonvlm/model/language_model/traj_decoder.py:39-47):
    [B,2,T_pred,token_dim] -> rows (b,hand,k) -> [B,2,T_pred,2]."""
    B, _, T_pred, Dc = pred_hand_embeddings.shape
    return traj_cvae_inference(pred_hand_embeddings.reshape(-1, Dc), z, sd).reshape(B, 2, T_pred, 2)


def traj_mlp_inference(pred_hand_embeddings: torch.Tensor, sd: dict) -> torch.Tensor:
    """``MLPTrajDecoder.inference`` (traj_decoder.py:39-47 -> ``TrajMLP.inference``,
    hoi_forecast/architecture/traj_decoder.py:139-147): Linear-ReLU-Linear-ReLU-Linear on the rows (b, hand, k).
    [B,2,T_pred,token_dim] -> [B,2,T_pred,2]."""
    pre = "hand_traj_decoder.mlp."
    B, _, T_pred, Dc = pred_hand_embeddings.shape
    h = pred_hand_embeddings.reshape(-1, Dc).float()
    h = torch.relu(h @ sd[pre + "0.weight"].float().t() + sd[pre + "0.bias"].float())
    h = torch.relu(h @ sd[pre + "2.weight"].float().t() + sd[pre + "2.bias"].float())
    h = h @ sd[pre + "4.weight"].float().t() + sd[pre + "4.bias"].float()
    return h.reshape(B, 2, T_pred, 2)


def traj_decode_step(hidden_last: torch.Tensor, z: torch.Tensor, sd: dict) -> torch.Tensor:
    """The generation loop's ``<hand_traj>`` branch in one call (handsonvlm.py:609-622): gather the last hidden
    row (even/odd de-interleave) and decode it -> [B,2,2] (the ``.squeeze(2)`` of handsonvlm.py:620)."""
    return traj_decoder_inference(gather_hand_traj_step(hidden_last), z, sd).squeeze(2)


# ----------------------------------------------------------------------------------------
# whole path (VisualToTokenHelper.pipeline, visual_to_tokens.py:23-37)
# ----------------------------------------------------------------------------------------

def pipeline(images: torch.Tensor, sd: dict, proj_w, proj_b, mode: str = "temporal_spatial_pool",
             select_layer: int = -2, cfg: VitCfg = VIT_L14, emulate: str | None = None):
    """images [b,t,3,H,W] -> (tokens [b,Nv,D], mask [b,Nv] all-True).  Reference order:
    project every token, then pool (visual_to_tokens.py:179-183, 252-271)."""
    b, t = images.shape[:2]
    feats = tower_forward(images.reshape(b * t, *images.shape[2:]), sd, select_layer, cfg, emulate=emulate)
    tok = project(feats.float(), proj_w, proj_b).reshape(b, t, feats.shape[1], -1)
    out = pool_tokens(tok, mode)
    return out, torch.ones(out.shape[:2], dtype=torch.bool)


def projector_grads(feats_pooled: torch.Tensor, d_tokens: torch.Tensor):
    """Pool-first identity (SURVEY.md section 8a notes): tokens = pool(X) W^T + b  =>
    dW = d_tokens^T pool(X), db = sum_rows d_tokens."""
    x = feats_pooled.reshape(-1, feats_pooled.shape[-1]).float()
    dy = d_tokens.reshape(-1, d_tokens.shape[-1]).float()
    return dy.t() @ x, dy.sum(0)


# ----------------------------------------------------------------------------------------
# f3: CLIPImageProcessor resize + centre crop (hoi_forecast/dataset/video_utils.py:28-53 ->
#     processor.preprocess; transformers==4.31.0 models/clip/image_processing_clip.py: resize
#     shortest edge -> 224 with PIL BICUBIC, centre crop 224, rescale 1/255, normalise).
# Third party: Pillow's 8-bit resampler (src/libImaging/Resample.c), restated from its published
# algorithm in integer numpy; pinned against PIL.Image.resize itself by tests/test_oracle_golden.py.
# ----------------------------------------------------------------------------------------

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_table(in_size: int, out_size: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for a bicubic in_size -> out_size pass:
    (bounds int [out,2] = (xmin, n_taps), coef int [out, ksize])."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    coef = np.zeros((out_size, ksize), np.int64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            coef[xx, x] = int(-0.5 + v * (1 << _PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, coef


def _resample_axis_last(img: np.ndarray, bounds: np.ndarray, coef: np.ndarray) -> np.ndarray:
    """img uint8 [..., in] -> uint8 [..., out] along the last axis (one 8-bit Pillow pass)."""
    out = np.empty(img.shape[:-1] + (bounds.shape[0],), np.uint8)
    src = img.astype(np.int64)
    for xx in range(bounds.shape[0]):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (1 << (_PRECISION_BITS - 1)) + (src[..., x0:x0 + n] * coef[xx, :n]).sum(-1)
        out[..., xx] = np.clip(acc >> _PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def clip_resize_output_size(h: int, w: int, shortest: int = 224):
    """transformers get_resize_output_image_size(size=int, default_to_square=False) -> (new_h, new_w)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = shortest, int(shortest * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def clip_resize_center_crop_u8(frames: np.ndarray, size: int = 224) -> np.ndarray:
    """frames uint8 [N,H,W,3] -> uint8 [N,size,size,3]: PIL BICUBIC resize of the shortest edge to `size` (horizontal
    pass first, then vertical, uint8 in between), then transformers' centre crop."""
    N, H, W, _ = frames.shape
    nh, nw = clip_resize_output_size(H, W, size)
    xb, xc = resample_table(W, nw)
    yb, yc = resample_table(H, nh)
    tmp = _resample_axis_last(np.ascontiguousarray(frames.transpose(0, 1, 3, 2)), xb, xc)      # [N,H,3,nw]
    out = _resample_axis_last(np.ascontiguousarray(tmp.transpose(0, 2, 3, 1)), yb, yc)         # [N,3,nw,nh]
    out = out.transpose(0, 3, 2, 1)                                                            # [N,nh,nw,3]
    top, left = (nh - size) // 2, (nw - size) // 2
    return np.ascontiguousarray(out[:, top:top + size, left:left + size])


def expand2square_u8(frames: np.ndarray, background=None) -> np.ndarray:
    """hoi_forecast/dataset/video_utils.py:13-25 (the `image_aspect_ratio == 'pad'` branch of load_image, :30-31):
    paste at (0, (w-h)//2) for landscape, ((h-w)//2, 0) for portrait, on a canvas of int(255 * image_mean)."""
    N, H, W, _ = frames.shape
    if H == W:
        return frames
    bg = tuple(int(x * 255) for x in CLIP_MEAN) if background is None else background
    S = max(H, W)
    out = np.empty((N, S, S, 3), np.uint8)
    out[...] = np.asarray(bg, np.uint8)
    if W > H:
        y0 = (W - H) // 2
        out[:, y0:y0 + H, :, :] = frames
    else:
        x0 = (H - W) // 2
        out[:, :, x0:x0 + W, :] = frames
    return out


def clip_normalize_u8(frames: np.ndarray) -> torch.Tensor:
    """uint8 [N,h,w,3] -> float32 [N,3,h,w]: rescale 1/255 + CLIP mean/std normalisation."""
    x = torch.from_numpy(frames).float().mul(1.0 / 255.0)
    x = (x - torch.tensor(CLIP_MEAN)) / torch.tensor(CLIP_STD)
    return x.permute(0, 3, 1, 2).contiguous()


# ----------------------------------------------------------------------------------------
# f2: frame de-duplication -- the reference has none (it encodes every repeated frame:
#     epic_dataset.py:90-95, hybrid_dataset.py:141-142); the specification is "same visual
#     tokens as without it", plus this definition of the frame map.
# ----------------------------------------------------------------------------------------

def frame_dedup(frames: torch.Tensor):
    """frames [N, ...] -> (frame_map int32 [N], rep int32 [U]): byte-wise equal frames share the
    unique index of their first occurrence; unique indices are numbered in order of first occurrence."""
    N = frames.shape[0]
    raw = frames.contiguous().reshape(N, -1).view(torch.uint8).numpy()
    first, fmap, rep = {}, [], []
    for i in range(N):
        key = raw[i].tobytes()
        if key not in first:
            first[key] = len(rep)
            rep.append(i)
        fmap.append(first[key])
    return torch.tensor(fmap, dtype=torch.int32), torch.tensor(rep, dtype=torch.int32)
