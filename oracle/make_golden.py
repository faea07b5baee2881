"""TEST INFRASTRUCTURE ONLY -- freeze golden fixtures from the real reference (dev container).

Run:  python -m oracle.make_golden          (needs /root/reference; CPU only, ~1-2 min)

Writes ``tests/golden/*.npz``.  Every fixture stores only *outputs* (and the tiny inputs that
are not regenerable); big inputs/weights are regenerated from ``oracle.synth`` seeds by the tests.
The fixtures pin ``oracle.restate`` (tests/test_oracle_golden.py) and, through it, the CUDA path.

Sources exercised (reference checkout, unmodified):
  * HF ``CLIPVisionModel`` inside ``CLIPVisionTower.forward``   llava/model/multimodal_encoder/clip_encoder.py:39-51
  * ``LitaMetaForCausalLM.videos_to_tokens``                   lita/model/lita_arch.py:30-77
  * ``VisualToTokenHelper.pipeline`` / ``compress_tokens``     hoi_forecast/model/visual_to_tokens.py:23-37,230-272
  * ``LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal``     llava/model/llava_arch.py:110-234
  * ``HandsOnVLMForCausalLM.prepare_inputs_labels_for_multimodal``    handsonvlm/.../handsonvlm.py:212-451
  * the inline <hand_traj> gather of ``HandsOnVLMForCausalLM.forward`` handsonvlm.py:146-187, executed
    from the reference's own source text (it is not a callable).
"""
from __future__ import annotations

import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn as nn

from . import ref_shim, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SMALL = synth.VitCfg(hidden=64, inter=128, layers=3, heads=4, image=224, patch=14)
SMALL_D = 64


def hf_model(cfg: synth.VitCfg, sd: dict):
    from transformers import CLIPVisionConfig, CLIPVisionModel
    c = CLIPVisionConfig(hidden_size=cfg.hidden, intermediate_size=cfg.inter, num_hidden_layers=cfg.layers,
                         num_attention_heads=cfg.heads, image_size=cfg.image, patch_size=cfg.patch,
                         projection_dim=32)
    m = CLIPVisionModel(c).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in k for k in missing), missing
    return m


def small_parts(seed=0):
    sd = synth.clip_state_dict(SMALL, seed=seed, profile="strong")
    ps = synth.projector_state(SMALL_D, SMALL.hidden, seed=1)
    proj = nn.Linear(SMALL.hidden, SMALL_D)
    proj.weight.data.copy_(ps["mm_projector.weight"])
    proj.bias.data.copy_(ps["mm_projector.bias"])
    emb = nn.Embedding(synth.VOCAB, SMALL_D)
    emb.weight.data.copy_(synth.embed_table(SMALL_D))
    return sd, proj, emb


def make_host(ns, mixin, tower, proj, emb, config, B):
    cls = type("Host", (ref_shim.FakeHost, mixin), {})
    return cls(tower, proj, emb, config, B)


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    conv = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **conv)
    print(f"wrote {name}.npz  ({sum(a.nbytes for a in conv.values()) / 1e3:.1f} kB raw)")


# ----------------------------------------------------------------------------------------

@torch.no_grad()
def golden_vit_full(ns):
    """Full ViT-L/14 (24 layers) through the reference tower; 2 frames; both weight profiles.
    Stores a strided subsample of hidden_states[-2][:,1:] plus per-token L2 norms."""
    for profile in ("hf", "strong"):
        sd = synth.clip_state_dict(synth.VIT_L14, seed=0, profile=profile)
        tower = ref_shim.build_tower(ns, hf_model(synth.VIT_L14, sd))
        px = synth.pixels((2, 3, 224, 224), seed=3)
        feats = tower(px)                                   # [2,256,1024]
        assert feats.shape == (2, 256, 1024)
        save(f"vit_l14_{profile}", sub=feats[:, ::8, ::4].contiguous(), norms=feats.norm(dim=-1),
             absmax=feats.abs().max())


@torch.no_grad()
def golden_pool(ns):
    """compress_tokens / videos_to_tokens pooling on raw random tokens."""
    H = ns.v2t.VisualToTokenHelper
    for t, d in ((100, 32), (10, 16), (2, 8), (5, 8)):
        tok = synth.gen(f"pooltok{t}", (2, t, 256, d), 1.0, seed=5)
        for mode in ("temporal_spatial_pool", "spatial_pool", "none"):
            h = H(None, None, "origin", mode, 0, d)
            h.b, h.t = 2, t
            out, mask = h.compress_tokens(tokens=tok, attention_mask=torch.ones(2, t, 256, dtype=torch.bool))
            if mode == "none":
                out = out[:, ::97]
            save(f"pool_{mode}_t{t}", out=out, mask_all_true=bool(mask.all()), mask_shape=np.array(mask.shape))
    # autograd of the reference pooling = oracle for the backward kernel
    tok = synth.gen("pooltok_bwd", (1, 10, 256, 8), 1.0, seed=6).double().requires_grad_(True)
    with torch.enable_grad():
        h = H(None, None, "origin", "temporal_spatial_pool", 0, 8)
        h.b, h.t = 1, 10
        out, _ = h.compress_tokens(tokens=tok, attention_mask=None)
        dout = synth.gen("pooldout_bwd", tuple(out.shape), 1.0, seed=6).double()
        (out * dout).sum().backward()
    save("pool_bwd_t10", dtok=tok.grad.float())


@torch.no_grad()
def golden_lita(ns):
    """LitaMetaForCausalLM.videos_to_tokens, every video_arch, small tower."""
    sd, proj, emb = small_parts()
    tower = ref_shim.build_tower(ns, hf_model(SMALL, sd), select_layer=-2)
    px = synth.pixels((1, 6, 3, 224, 224), seed=7)
    for arch in ("all", "temporal", "spatial", "temporal_spatial", "temporal_spatial_pool", "spatial_pool"):
        cfg = types.SimpleNamespace(video_arch=arch, input_type="video")
        host = make_host(ns, ns.lita_arch.LitaMetaForCausalLM, tower, proj, emb, cfg, 1)
        out = host.visual_to_tokens(px)
        save(f"lita_{arch}", out=out[:, ::7] if arch == "all" else out, shape=np.array(out.shape))
    # VisualToTokenHelper.pipeline == videos_to_tokens
    h = ns.v2t.VisualToTokenHelper(tower, proj, "origin", "temporal_spatial_pool", SMALL.hidden, SMALL_D)
    out, mask = h.pipeline(images=px)
    save("v2t_pipeline", out=out, mask=mask)


def _hvlm_host(ns, tower, proj, emb, B, mode="temporal_spatial_pool"):
    cfg = types.SimpleNamespace(fuse_input_mode="origin", video_compress_mode=mode, mm_hidden_size=SMALL.hidden,
                                input_type="video")
    host = ref_shim.FakeHost(tower, proj, emb, cfg, B)
    return host


@torch.no_grad()
def golden_splice(ns):
    sd, proj, emb = small_parts()
    tower = ref_shim.build_tower(ns, hf_model(SMALL, sd), select_layer=-2)
    HV = ns.handsonvlm.HandsOnVLMForCausalLM.prepare_inputs_labels_for_multimodal
    t = 4

    def run_hvlm(name, ids, mask, labels, fh, fv, is_eval=False, B=None, px_seed=11):
        B = ids.shape[0] if B is None else B
        px = synth.pixels((B, t, 3, 224, 224), seed=px_seed)
        host = _hvlm_host(ns, tower, proj, emb, B)
        kw = {}
        if fh is not None:
            kw["future_hands"] = fh
        if fv is not None:
            kw["future_valid"] = fv
        _, m2, _, e2, l2 = HV(host, ids, mask, None, labels, px, is_evaluate=is_eval, **kw)
        arrs = dict(embeds=e2, ids=ids, px_seed=px_seed, t=t, last_visual_token_index=int(host.last_visual_token_index)
                    if hasattr(host, "last_visual_token_index") else -1)
        if m2 is not None:
            arrs["mask"] = m2
            arrs["mask_dtype"] = str(m2.dtype)
        if l2 is not None:
            arrs["labels"] = l2
        if mask is not None:
            arrs["in_mask"] = mask
        if labels is not None:
            arrs["in_labels"] = labels
        if fh is not None:
            arrs["future_hands"] = fh
        save(name, **arrs)

    # (a) equal-length batch, 4 hand tokens, training mode
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=3, seed=21, n_pre=7, n_post=5)
    run_hvlm("splice_hvlm_train_b3", ids, mask, labels, fh, fv)
    # (b) collator-padded batch (ragged prompts right-padded to equal T before the splice)
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=3, seed=22, ragged=True)
    run_hvlm("splice_hvlm_train_padded", ids, mask, labels, fh, fv)
    # (c) 2 hand tokens: cnt/4 scaling + the row-0 duplicate-index quirk
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=1, seed=23, n_pre=5, n_post=4, n_hand=2)
    run_hvlm("splice_hvlm_2hand", ids, mask, labels, fh, fv)
    # (d) 0 hand tokens
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=1, seed=24, n_pre=5, n_post=4, n_hand=0)
    run_hvlm("splice_hvlm_0hand", ids, mask, labels, fh, fv)
    # (e) truly ragged output: sample 1 has no image token -> int64 mask padded with -100
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=2, seed=25, n_pre=6, n_post=3)
    ids[1, 6] = 1234
    run_hvlm("splice_hvlm_ragged", ids, mask, labels, fh, fv)
    # (f) evaluation: labels=None, mask=None, future_hands [B,2,n,2] with n == #hand tokens
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=1, seed=26, n_pre=6, n_post=3, n_hand=3)
    run_hvlm("splice_hvlm_eval_hands", ids, None, None, fh[:, :, :3], None, is_eval=True)
    # (g) evaluation without future_hands
    run_hvlm("splice_hvlm_eval_nohands", ids, mask, None, None, None, is_eval=True)
    # (h) image token is the last token (empty tail)
    ids = torch.cat([synth._rand_ids("h", 9, 27), torch.tensor([synth.IMAGE_TOKEN_INDEX])]).unsqueeze(0)
    run_hvlm("splice_hvlm_empty_tail", ids, torch.ones_like(ids, dtype=torch.bool), ids.clone(),
             synth.gen("fh", (1, 2, 4, 2), 1.0, 27), torch.ones(1, 2, dtype=torch.bool))

    # (i) two image tokens in sample 0 (it consumes visual slots 0 and 1), none in sample 1: the surviving
    #     last_visual_token_index is the SECOND token's position relative to the ids left after the first one
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=2, seed=28, n_pre=6, n_post=5)
    ids[0, 10] = synth.IMAGE_TOKEN_INDEX
    ids[1, 6] = 4321
    run_hvlm("splice_hvlm_two_images", ids, mask, labels, fh, fv)

    # LLaVA / LITA variant (image input: [B,3,224,224] -> 256 tokens), config-1 shaped
    LV = ns.llava_arch.LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal

    def run_llava(name, ids, mask, labels, px, cfg):
        host = make_host(ns, ns.lita_arch.LitaMetaForCausalLM, tower, proj, emb, cfg, ids.shape[0])
        r_ids, m2, _, e2, l2 = LV(host, ids, mask, None, labels, px)
        arrs = dict(ids=ids, in_mask=mask, mask=m2)
        if e2 is not None:
            arrs["embeds"] = e2
        if l2 is not None:
            arrs["labels"] = l2
        if labels is not None:
            arrs["in_labels"] = labels
        arrs["returned_ids_is_none"] = r_ids is None
        save(name, **arrs)

    cfg_img = types.SimpleNamespace(input_type="image")
    ids, mask, labels = synth.prompt_llava(seed=31)
    run_llava("splice_llava_cfg1", ids, mask, labels, synth.pixels((1, 3, 224, 224), seed=12), cfg_img)
    # ragged: sample 1 has no image token (still consumes an image slot)
    ids2 = torch.cat([ids, ids], 0).clone()
    ids2[1, 35] = 77
    mask2 = torch.ones_like(ids2, dtype=torch.bool)
    mask2[1, -3:] = False
    run_llava("splice_llava_ragged", ids2, mask2, ids2.clone(), synth.pixels((2, 3, 224, 224), seed=13), cfg_img)
    # two image tokens in one sample, image batch of 2
    ids3 = ids.clone()
    ids3[0, 10] = synth.IMAGE_TOKEN_INDEX
    run_llava("splice_llava_two_images", ids3, torch.ones_like(ids3, dtype=torch.bool), ids3.clone(),
              synth.pixels((2, 3, 224, 224), seed=14), cfg_img)
    # video input through the LLaVA splice (LITA caller, lita_llama.py:85)
    cfg_vid = types.SimpleNamespace(input_type="video", video_arch="temporal_spatial_pool")
    run_llava("splice_llava_video", ids, mask, labels, synth.pixels((1, t, 3, 224, 224), seed=15), cfg_vid)
    # T == 1 early-out (llava_arch.py:117-120)
    one = torch.tensor([[5]])
    host = make_host(ns, ns.lita_arch.LitaMetaForCausalLM, tower, proj, emb, cfg_img, 1)
    pkv = [[torch.zeros(1, 2, 9, 4), torch.zeros(1, 2, 9, 4)]]
    r_ids, m2, _, e2, _ = LV(host, one, torch.ones(1, 1, dtype=torch.bool), pkv, None,
                             synth.pixels((1, 3, 224, 224), seed=12))
    save("splice_llava_t1", mask=m2, embeds_is_none=e2 is None, ids=r_ids)


def golden_splice_im_start_end(ns):
    """The ``tune_mm_mlp_adapter and mm_use_im_start_end`` branch of llava_arch.py:146-161,172-181, run as is: outputs and
    the gradient that reaches ``embed_tokens.weight`` (only the two tokens around the image token are not detached)."""
    sd, proj, emb = small_parts()
    tower = ref_shim.build_tower(ns, hf_model(SMALL, sd), select_layer=-2)
    LV = ns.llava_arch.LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal
    cfg = types.SimpleNamespace(input_type="image", tune_mm_mlp_adapter=True, mm_use_im_start_end=True)
    ids, mask, _ = synth.prompt_llava(seed=33)
    ids = torch.cat([ids, ids], 0).clone()
    ids[1, 20:30] = torch.arange(100, 110)
    labels = torch.arange(ids.numel(), dtype=torch.int64).reshape(ids.shape) + 1000      # all distinct
    mask = torch.ones_like(ids, dtype=torch.bool)
    host = make_host(ns, ns.lita_arch.LitaMetaForCausalLM, tower, proj, emb, cfg, ids.shape[0])
    emb.weight.grad = None
    with torch.enable_grad():
        r_ids, m2, _, e2, l2 = LV(host, ids, mask, None, labels, synth.pixels((2, 3, 224, 224), seed=16))
        w = synth.gen("ise.dout", tuple(e2.shape), 1.0, seed=16)
        (e2 * w).sum().backward()
    g = emb.weight.grad
    rows = torch.nonzero(g.abs().sum(1) > 0).flatten()
    save("splice_llava_im_start_end", ids=ids, in_mask=mask, in_labels=labels, mask=m2, embeds=e2.detach(), labels=l2,
         grad_rows=rows, grad_vals=g[rows])
    emb.weight.grad = None


def golden_splice_hvlm_im_start_end(ns):
    """HandsOnVLM's own splice with tune_mm_mlp_adapter + mm_use_im_start_end (handsonvlm.py:263-286,343-344)."""
    sd, proj, emb = small_parts()
    tower = ref_shim.build_tower(ns, hf_model(SMALL, sd), select_layer=-2)
    HV = ns.handsonvlm.HandsOnVLMForCausalLM.prepare_inputs_labels_for_multimodal
    t = 4
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=2, seed=23, n_pre=7, n_post=5)
    labels = torch.arange(ids.numel(), dtype=torch.int64).reshape(ids.shape) + 1000
    mask = torch.ones_like(ids, dtype=torch.bool)
    s = int((ids[0] == synth.IMAGE_TOKEN_INDEX).nonzero()[0])
    mask[0, s] = False                  # visible in the output only through the <im_end> slot rule
    mask[1, -2:] = False
    host = _hvlm_host(ns, tower, proj, emb, 2)
    host.config.tune_mm_mlp_adapter = True
    host.config.mm_use_im_start_end = True
    px = synth.pixels((2, t, 3, 224, 224), seed=17)
    _, m2, _, e2, l2 = HV(host, ids, mask, None, labels, px, is_evaluate=False, future_hands=fh, future_valid=fv)
    save("splice_hvlm_im_start_end", ids=ids, in_mask=mask, in_labels=labels, mask=m2, mask_dtype=str(m2.dtype),
         embeds=e2.detach(), labels=l2, t=t, has_last_visual_token_index=hasattr(host, "last_visual_token_index"))


@torch.no_grad()
def golden_splice_llava_list(ns):
    """The list path of images_to_tokens (llava_arch.py:95-106): per-sample image groups of DIFFERENT sizes; each
    sample's single image token expands to its whole group (n_i * 256 tokens), so the batch is ragged."""
    sd, proj, emb = small_parts()
    tower = ref_shim.build_tower(ns, hf_model(SMALL, sd), select_layer=-2)
    LV = ns.llava_arch.LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal
    cfg = types.SimpleNamespace(input_type="image")
    ids, mask, labels = synth.prompt_llava(seed=34)
    ids = torch.cat([ids, ids, ids], 0).clone()
    ids[1, 5:9] = torch.arange(200, 204)
    ids[2, 35] = 99                                   # sample 2 has no image token (still consumes a slot)
    labels = torch.arange(ids.numel(), dtype=torch.int64).reshape(ids.shape) + 2000
    mask = torch.ones_like(ids, dtype=torch.bool)
    mask[1, -4:] = False
    groups = [synth.pixels((1, 3, 224, 224), seed=18), synth.pixels((3, 3, 224, 224), seed=19),
              synth.pixels((2, 3, 224, 224), seed=20)]
    host = make_host(ns, ns.lita_arch.LitaMetaForCausalLM, tower, proj, emb, cfg, ids.shape[0])
    r_ids, m2, _, e2, l2 = LV(host, ids, mask, None, labels, groups)
    save("splice_llava_list_ragged", ids=ids, in_mask=mask, in_labels=labels, mask=m2, embeds=e2, labels=l2,
         group_sizes=[g.shape[0] for g in groups])


def golden_gather(ns):
    """Execute the reference's inline gather (handsonvlm.py, inside forward) from its source text."""
    import inspect
    src = inspect.getsource(ns.handsonvlm.HandsOnVLMForCausalLM.forward).splitlines()
    start = next(i for i, l in enumerate(src) if "hand_traj_token_idx = 32100" in l)
    end = next(i for i, l in enumerate(src) if "pred_hand_embeddings = torch.stack" in l)
    code = textwrap.dedent("\n".join(src[start:end + 1]))

    def run(hidden, labels, future_valid):
        B, L, D = hidden.shape
        env = dict(torch=torch, labels=labels, hidden_states=hidden, B=B, T_modified=L,
                   self=types.SimpleNamespace(token_dim=D), future_valid=future_valid)
        exec(code, env)
        return env["pred_hand_embeddings"], env["future_valid"]

    # toy known-answer from SURVEY.md section 8a (a7)
    hidden = torch.arange(2 * 10 * 8, dtype=torch.float32).reshape(2, 10, 8)
    labels = torch.full((2, 10), -100, dtype=torch.int64)
    labels[0, 5:9] = 32100
    fv = torch.ones(2, 2, dtype=torch.bool)
    out, fv2 = run(hidden, labels, fv)
    save("gather_toy", out=out, future_valid=fv2, labels=labels)
    # random, D=64, L=40, B=4: positions differ per sample, one sample without hand tokens
    hidden = synth.gen("gather_hidden", (4, 40, 64), 1.0, seed=41)
    labels = torch.full((4, 40), -100, dtype=torch.int64)
    labels[0, 30:34] = 32100
    labels[1, [3, 9, 17, 39]] = 32100
    labels[3, 1:5] = 32100
    labels[3, 20] = 5
    fv = torch.ones(4, 2, dtype=torch.bool)
    out, fv2 = run(hidden, labels, fv)
    save("gather_rand", out=out, future_valid=fv2, labels=labels)
    # hand token at position 0 only counts through the shift (label[0] has no predictor)
    labels = torch.full((1, 12), -100, dtype=torch.int64)
    labels[0, [0, 2, 3, 4, 5]] = 32100
    hidden = synth.gen("gather_hidden0", (1, 12, 16), 1.0, seed=42)
    out, fv2 = run(hidden, labels, torch.ones(1, 2, dtype=torch.bool))
    save("gather_pos0", out=out, future_valid=fv2, labels=labels)


def golden_traj(ns):
    """The reference's CVAETrajDecoder.inference (generation side of the trajectory head), run as
    is; the noise it draws internally is reproduced by reseeding torch's global generator and stored with the output."""
    import importlib
    mod = importlib.import_module("handsonvlm.model.language_model.traj_decoder")
    Dc = 32
    dec = mod.CVAETrajDecoder(token_dim=Dc)
    sd = synth.traj_cvae_state(Dc, seed=3)
    dec.load_state_dict(sd, strict=True)
    with torch.no_grad():
        # general inference: [B,2,T_pred,Dc]
        emb = synth.gen("traj_emb", (3, 2, 4, Dc), 1.0, seed=51)
        torch.manual_seed(1234)
        out = dec.inference(pred_hand_embeddings=emb)
        torch.manual_seed(1234)
        z = dec.hand_traj_decoder.z_scale * torch.randn([3 * 2 * 4, 256])
        save("traj_infer", out=out, z=z)
        # the generation step exactly as handsonvlm.py:613-620 does it (batch 1)
        hidden_last = synth.gen("traj_hidden_last", (1, 2 * Dc), 1.0, seed=52)
        e = hidden_last.reshape(1, Dc, 2).permute(0, 2, 1).unsqueeze(2)
        torch.manual_seed(77)
        out = dec.inference(pred_hand_embeddings=e).squeeze(2)
        torch.manual_seed(77)
        z = dec.hand_traj_decoder.z_scale * torch.randn([2, 256])
        save("traj_step", out=out, z=z)
        # the MLP decoder (traj_decoder 'MLP'): deterministic
        mlp = mod.MLPTrajDecoder(token_dim=Dc)
        mlp.load_state_dict(synth.traj_mlp_state(Dc, seed=4), strict=True)
        emb = synth.gen("trajmlp_emb", (3, 2, 4, Dc), 1.0, seed=53)
        save("traj_mlp_infer", out=mlp.inference(pred_hand_embeddings=emb))


def golden_preprocess(ns):
    """hoi_forecast/dataset/video_utils.py:28-53 (load_image) on in-memory frames: the reference's own expand2square for the
    'pad' branch, then `processor.preprocess` with the PIL-backed CLIPImageProcessor (what transformers==4.31.0 runs)."""
    import importlib.util
    import transformers
    from PIL import Image
    spec = importlib.util.spec_from_file_location("ref_video_utils", os.path.join(ref_shim.REF, "hoi_forecast/dataset/video_utils.py"))
    vu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vu)
    proc = getattr(transformers.models.clip, "CLIPImageProcessorPil", transformers.CLIPImageProcessor)()
    rng = np.random.RandomState(77)
    for name, (H, W) in (("landscape", (128, 228)), ("portrait", (150, 100))):
        frame = rng.randint(0, 256, (H, W, 3), dtype=np.uint8)
        yy, xx = np.mgrid[0:H, 0:W]
        frame[..., 1] = ((xx * 3 + yy * 5) % 256).astype(np.uint8)             # some structure next to the noise
        img = Image.fromarray(frame)
        sq = proc.preprocess(img, return_tensors="pt")["pixel_values"][0]
        padded = vu.expand2square(img, tuple(int(x * 255) for x in proc.image_mean))
        pd = proc.preprocess(padded, return_tensors="pt")["pixel_values"][0]
        save(f"preprocess_{name}", frame=frame, square=sq[:, ::2, ::2].contiguous(), pad=pd[:, ::2, ::2].contiguous(),
             padded_size=np.array(padded.size))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    ns = ref_shim.load()
    if ns.handsonvlm is None:
        raise ns.handsonvlm_error
    only = sys.argv[1:]
    if only:                                  # e.g. `python -m oracle.make_golden golden_splice`
        for name in only:
            globals()[name](ns)
        return
    golden_pool(ns)
    golden_gather(ns)
    golden_traj(ns)
    golden_lita(ns)
    golden_splice(ns)
    golden_splice_im_start_end(ns)
    golden_splice_hvlm_im_start_end(ns)
    golden_splice_llava_list(ns)
    golden_vit_full(ns)
    golden_preprocess(ns)


if __name__ == "__main__":
    main()
