"""TEST INFRASTRUCTURE ONLY -- seeded synthetic weights and inputs.

There is no network, so no pretrained CLIP / Vicuna checkpoint exists anywhere in this
project.  Every tensor is generated from ``(name, seed)`` with a private CPU
``torch.Generator`` so that (a) the same tensors can be loaded into HuggingFace's
``CLIPVisionModel`` (golden generation, dev container), into ``oracle.restate`` and into the
CUDA path's packed weight blob, and (b) adding/removing layers never shifts other tensors.

Tensor names follow the HuggingFace ``CLIPVisionModel.state_dict()`` keys (the model the
reference instantiates at ``llava/model/multimodal_encoder/clip_encoder.py:24``), the
projector keys ``mm_projector.{weight,bias}`` (``llava/model/llava_arch.py:33``) and
``embed_tokens.weight``.

Input builders reproduce the concrete configs of SURVEY.md section 8(d).
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass

import torch

IGNORE_INDEX = -100          # handsonvlm/constants.py:12
IMAGE_TOKEN_INDEX = -200     # handsonvlm/constants.py:13
HAND_TRAJ_TOKEN_ID = 32100   # handsonvlm/model/language_model/handsonvlm.py:146,349,609
VOCAB = 32101                # 32000 Vicuna + 100 <t*> + <hand_traj>  (handsonvlm/train/train.py:365-368)


@dataclass(frozen=True)
class VitCfg:
    hidden: int = 1024
    inter: int = 4096
    layers: int = 24
    heads: int = 16
    image: int = 224
    patch: int = 14

    @property
    def grid(self) -> int:
        return self.image // self.patch

    @property
    def n_patches(self) -> int:
        return self.grid * self.grid

    @property
    def seq(self) -> int:
        return self.n_patches + 1


VIT_L14 = VitCfg()


def _seed_for(name: str, seed: int) -> int:
    return (zlib.crc32(name.encode()) ^ ((seed + 1) * 0x9E3779B1)) & 0x7FFFFFFF


def gen(name: str, shape, std: float, seed: int = 0, mean: float = 0.0) -> torch.Tensor:
    """Deterministic fp32 N(mean, std^2) tensor keyed by (name, seed)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(_seed_for(name, seed))
    t = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    return t.mul_(std).add_(mean)


def clip_state_dict(cfg: VitCfg = VIT_L14, seed: int = 0, profile: str = "hf",
                    n_layers: int | None = None) -> dict:
    """HF-named fp32 state dict for CLIPVisionModel.

    profile 'hf'     : weight stds of HF's CLIP ``_init_weights`` (factor 1.0); biases and LN
                       affine parameters are *not* left at 0/1 but perturbed, so bias / gamma /
                       beta code paths are exercised by every parity test.
    profile 'strong' : unit-variance q/k/v and O(1) residual updates per layer -- attention and
                       MLP contribute as much as the residual, so a wrong layer cannot hide
                       under the tolerance.
    """
    E, FF, L = cfg.hidden, cfg.inter, cfg.layers if n_layers is None else n_layers
    if profile == "hf":
        s_qkv = (E ** -0.5) * ((2 * cfg.layers) ** -0.5)
        s_out = E ** -0.5
        s_fc1 = (E ** -0.5) * ((2 * cfg.layers) ** -0.5)
        s_fc2 = (2 * E) ** -0.5
        s_bias, s_lnw, s_lnb = 0.02, 0.1, 0.05
    elif profile == "strong":
        s_qkv = E ** -0.5
        s_out = 0.5 * E ** -0.5
        s_fc1 = E ** -0.5
        s_fc2 = 0.5 * FF ** -0.5
        s_bias, s_lnw, s_lnb = 0.1, 0.2, 0.1
    else:
        raise ValueError(profile)
    p = "vision_model."
    sd = {
        p + "embeddings.class_embedding": gen("cls", (E,), E ** -0.5, seed),
        p + "embeddings.patch_embedding.weight": gen("patch", (E, 3, cfg.patch, cfg.patch), 0.02, seed),
        p + "embeddings.position_embedding.weight": gen("pos", (cfg.seq, E), 0.02, seed),
        p + "pre_layrnorm.weight": gen("preln.w", (E,), s_lnw, seed, 1.0),
        p + "pre_layrnorm.bias": gen("preln.b", (E,), s_lnb, seed),
        p + "post_layernorm.weight": gen("postln.w", (E,), s_lnw, seed, 1.0),
        p + "post_layernorm.bias": gen("postln.b", (E,), s_lnb, seed),
    }
    for l in range(L):
        q = f"{p}encoder.layers.{l}."
        for nm, std in (("q_proj", s_qkv), ("k_proj", s_qkv), ("v_proj", s_qkv), ("out_proj", s_out)):
            sd[f"{q}self_attn.{nm}.weight"] = gen(f"l{l}.{nm}.w", (E, E), std, seed)
            sd[f"{q}self_attn.{nm}.bias"] = gen(f"l{l}.{nm}.b", (E,), s_bias, seed)
        sd[q + "layer_norm1.weight"] = gen(f"l{l}.ln1.w", (E,), s_lnw, seed, 1.0)
        sd[q + "layer_norm1.bias"] = gen(f"l{l}.ln1.b", (E,), s_lnb, seed)
        sd[q + "layer_norm2.weight"] = gen(f"l{l}.ln2.w", (E,), s_lnw, seed, 1.0)
        sd[q + "layer_norm2.bias"] = gen(f"l{l}.ln2.b", (E,), s_lnb, seed)
        sd[q + "mlp.fc1.weight"] = gen(f"l{l}.fc1.w", (FF, E), s_fc1, seed)
        sd[q + "mlp.fc1.bias"] = gen(f"l{l}.fc1.b", (FF,), s_bias, seed)
        sd[q + "mlp.fc2.weight"] = gen(f"l{l}.fc2.w", (E, FF), s_fc2, seed)
        sd[q + "mlp.fc2.bias"] = gen(f"l{l}.fc2.b", (E,), s_bias, seed)
    return sd


def projector_state(D: int, E: int = 1024, seed: int = 1) -> dict:
    """``mm_projector = nn.Linear(E, D)`` (llava_arch.py:33) -- nn.Linear-like scale."""
    bound = E ** -0.5
    return {"mm_projector.weight": gen("proj.w", (D, E), bound * 0.577, seed),
            "mm_projector.bias": gen("proj.b", (D,), bound * 0.577, seed)}


def traj_cvae_state(token_dim: int, hidden: int = 512, latent: int = 256, in_dim: int = 2, seed: int = 1) -> dict:
    """State dict of ``CVAETrajDecoder(token_dim)`` (handsonvlm/model/language_model/traj_decoder.py:60-70 ->
    TrajCVAE -> VAE, hoi_forecast/architecture/decoder_modules.py:5-30), keys as the reference names them,
    nn.Linear-like scale.  ``token_dim`` here is the per-hand width D/2."""
    pre = "hand_traj_decoder.cvae."
    shapes = {"enc_MLP.0": (hidden, in_dim + token_dim), "linear_means": (latent, hidden),
              "linear_log_var": (latent, hidden), "dec_MLP.0": (hidden, latent + token_dim), "dec_MLP.2": (in_dim, hidden)}
    sd = {}
    for k, (o, i) in shapes.items():
        sd[pre + k + ".weight"] = gen("traj." + k + ".w", (o, i), 0.577 * i ** -0.5, seed)
        sd[pre + k + ".bias"] = gen("traj." + k + ".b", (o,), 0.577 * i ** -0.5, seed)
    return sd


def traj_mlp_state(token_dim: int, hidden: int = 512, seed: int = 1) -> dict:
    """State dict of ``MLPTrajDecoder(token_dim)`` (handsonvlm/model/language_model/traj_decoder.py:50-57 -> TrajMLP,
    hoi_forecast/architecture/traj_decoder.py:94-104): ``hand_traj_decoder.mlp.{0,2,4}``."""
    pre = "hand_traj_decoder.mlp."
    sd = {}
    for k, (o, i) in {"0": (hidden, token_dim), "2": (hidden, hidden), "4": (2, hidden)}.items():
        sd[pre + k + ".weight"] = gen("trajmlp." + k + ".w", (o, i), 0.577 * i ** -0.5, seed)
        sd[pre + k + ".bias"] = gen("trajmlp." + k + ".b", (o,), 0.577 * i ** -0.5, seed)
    return sd


def embed_table(D: int, vocab: int = VOCAB, seed: int = 1) -> torch.Tensor:
    """``embed_tokens = nn.Embedding(vocab, D)`` -- N(0,1) like nn.Embedding's default."""
    return gen("embed_tokens", (vocab, D), 1.0, seed)


def pixels(shape, seed: int = 0) -> torch.Tensor:
    return gen("pixels", shape, 1.0, seed)


# ----------------------------------------------------------------------------------------
# prompts
# ----------------------------------------------------------------------------------------

def _rand_ids(name: str, n: int, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(_seed_for(name, seed))
    return torch.randint(1, 32000, (n,), generator=g, dtype=torch.int64)


def prompt_llava(seed: int = 0, n_pre: int = 35, n_post: int = 20):
    """Config 1: 35 random ids ++ [-200] ++ 20 random ids; mask all True; labels = ids."""
    ids = torch.cat([_rand_ids("pre", n_pre, seed), torch.tensor([IMAGE_TOKEN_INDEX]),
                     _rand_ids("post", n_post, seed)]).unsqueeze(0)
    mask = torch.ones_like(ids, dtype=torch.bool)
    labels = ids.clone()
    return ids, mask, labels


def prompt_handsonvlm(B: int = 1, seed: int = 0, n_pre: int = 35, n_post: int = 20,
                      n_hand: int = 4, ragged: bool = False):
    """Configs 2-5: 35 ++ [-200] ++ 20 ++ [32100]x4 ++ [869, 2]; labels = ids with the first
    n_pre+1+n_post positions set to -100.  With ``ragged=True`` every sample gets a random
    prompt length and the batch is right-padded the way the collator does it
    (handsonvlm/dataset/hybrid_dataset.py:155-158: ids 0, labels -100, mask False)."""
    rows, labs = [], []
    g = torch.Generator(device="cpu")
    g.manual_seed(_seed_for("ragged", seed))
    for b in range(B):
        npre = n_pre if not ragged else int(torch.randint(10, 40, (1,), generator=g))
        npost = n_post if not ragged else int(torch.randint(8, 40, (1,), generator=g))
        ids = torch.cat([_rand_ids(f"pre{b}", npre, seed), torch.tensor([IMAGE_TOKEN_INDEX]),
                         _rand_ids(f"post{b}", npost, seed),
                         torch.full((n_hand,), HAND_TRAJ_TOKEN_ID, dtype=torch.int64),
                         torch.tensor([869, 2])])
        lab = ids.clone()
        lab[: npre + 1 + npost] = IGNORE_INDEX
        rows.append(ids)
        labs.append(lab)
    T = max(r.numel() for r in rows)
    ids = torch.zeros(B, T, dtype=torch.int64)
    labels = torch.full((B, T), IGNORE_INDEX, dtype=torch.int64)
    mask = torch.zeros(B, T, dtype=torch.bool)
    for b, (r, l) in enumerate(zip(rows, labs)):
        ids[b, : r.numel()] = r
        labels[b, : r.numel()] = l
        mask[b, : r.numel()] = True
    future_hands = gen("future_hands", (B, 2, 4, 2), 1.0, seed).abs().clamp_(max=1.0)
    future_valid = torch.ones(B, 2, dtype=torch.bool)
    return ids, mask, labels, future_hands, future_valid
