"""HF-named CLIPVisionModel state dict  ->  the packed device blob defined by hvlm_vit_layout.

State-dict keys are the ones HuggingFace's ``CLIPVisionModel`` uses (the model the reference builds at
llava/model/multimodal_encoder/clip_encoder.py:24), so a real ``openai/clip-vit-large-patch14`` checkpoint loads
unchanged.  q_proj weight/bias are pre-multiplied by 64^-1/2 (a power of two: exact in bf16), which removes the
score scaling from the attention kernel.

ABI 3 adds the operands of the FOLDED LayerNorms (``hvlm_vit_layout.fold``): the tower runs LN1 / LN2 inside the QKV / fc1
GEMMs, ``LN(x) W^T + b = rstd * (bf16(x) (gamma*W)^T - mean * c) + (b + W beta)``, so per layer the blob also holds
``bf16(gamma*W)`` (rounded once, from the fp32 weights), ``c`` = the row sums of those rounded weights, and ``b + W beta``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

P = "vision_model."


def vit_layout(n_layers: int) -> L.VitLayout:
    lay = L.VitLayout()
    L.check(L.lib().hvlm_vit_l14_layout(n_layers, C.byref(lay)), "hvlm_vit_l14_layout")
    return lay


def fold_layernorm(w: torch.Tensor, b: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor):
    """Operands of ``LN(x) W^T + b`` with the LayerNorm folded into the GEMM (fp32 in):
    (bf16(gamma * W) [N,K], c [N] = row sums of the ROUNDED folded weights, b + W beta [N])."""
    w = w.detach().to("cpu", torch.float32)
    w_f = (w * gamma.to("cpu", torch.float32).unsqueeze(0)).to(torch.bfloat16)
    c = w_f.to(torch.float64).sum(1).to(torch.float32)
    b_f = (b.detach().to("cpu", torch.float64) + w.to(torch.float64) @ beta.to("cpu", torch.float64)).to(torch.float32)
    return w_f, c, b_f


def pack_vit_weights(sd: dict, n_layers: int = 24, device=None) -> torch.Tensor:
    """Returns a uint8 tensor (the blob).  Missing trailing layers are allowed: n_layers is clamped to the
    number of layers present in ``sd``."""
    have = 0
    while f"{P}encoder.layers.{have}.layer_norm1.weight" in sd:
        have += 1
    n_layers = min(n_layers, have)
    lay = vit_layout(n_layers)
    blob = torch.zeros(int(lay.total_bytes), dtype=torch.uint8)

    def put(off: int, t: torch.Tensor, dtype: torch.dtype):
        t = t.detach().to("cpu").to(dtype).contiguous()
        raw = t.view(torch.uint8).reshape(-1)
        blob[off: off + raw.numel()] = raw

    E = 1024
    pw = sd[P + "embeddings.patch_embedding.weight"].detach().float().reshape(E, -1)        # [1024, 588]
    assert pw.shape == (E, 588), pw.shape
    pw_pad = torch.zeros(E, 640)
    pw_pad[:, :588] = pw
    put(lay.patch_w, pw_pad, torch.bfloat16)
    put(lay.cls, sd[P + "embeddings.class_embedding"].reshape(E), torch.float32)
    pos = sd[P + "embeddings.position_embedding.weight"]
    assert tuple(pos.shape) == (257, E), pos.shape
    put(lay.pos, pos, torch.float32)
    put(lay.pre_ln_g, sd[P + "pre_layrnorm.weight"], torch.float32)
    put(lay.pre_ln_b, sd[P + "pre_layrnorm.bias"], torch.float32)
    for l in range(n_layers):
        q = f"{P}encoder.layers.{l}."
        y = lay.layer[l]
        put(y.ln1_g, sd[q + "layer_norm1.weight"], torch.float32)
        put(y.ln1_b, sd[q + "layer_norm1.bias"], torch.float32)
        wq = sd[q + "self_attn.q_proj.weight"].detach().float() * 0.125
        bq = sd[q + "self_attn.q_proj.bias"].detach().float() * 0.125
        w_qkv = torch.cat([wq, sd[q + "self_attn.k_proj.weight"].detach().float(),
                           sd[q + "self_attn.v_proj.weight"].detach().float()], 0)
        b_qkv = torch.cat([bq, sd[q + "self_attn.k_proj.bias"].detach().float(),
                           sd[q + "self_attn.v_proj.bias"].detach().float()], 0)
        assert w_qkv.shape == (3 * E, E)
        put(y.w_qkv, w_qkv, torch.bfloat16)
        put(y.b_qkv, b_qkv, torch.float32)
        put(y.w_o, sd[q + "self_attn.out_proj.weight"], torch.bfloat16)
        put(y.b_o, sd[q + "self_attn.out_proj.bias"], torch.float32)
        put(y.ln2_g, sd[q + "layer_norm2.weight"], torch.float32)
        put(y.ln2_b, sd[q + "layer_norm2.bias"], torch.float32)
        put(y.w_fc1, sd[q + "mlp.fc1.weight"], torch.bfloat16)
        put(y.b_fc1, sd[q + "mlp.fc1.bias"], torch.float32)
        put(y.w_fc2, sd[q + "mlp.fc2.weight"], torch.bfloat16)
        put(y.b_fc2, sd[q + "mlp.fc2.bias"], torch.float32)
        f = lay.fold[l]
        for w, b, g, beta, o_w, o_c, o_b in (
                (w_qkv, b_qkv, sd[q + "layer_norm1.weight"], sd[q + "layer_norm1.bias"], f.w_qkv_f, f.c_qkv, f.b_qkv_f),
                (sd[q + "mlp.fc1.weight"].detach().float(), sd[q + "mlp.fc1.bias"].detach().float(),
                 sd[q + "layer_norm2.weight"], sd[q + "layer_norm2.bias"], f.w_fc1_f, f.c_fc1, f.b_fc1_f)):
            w_f, c, b_f = fold_layernorm(w, b, g.detach().float(), beta.detach().float())
            put(o_w, w_f, torch.bfloat16)
            put(o_c, c, torch.float32)
            put(o_b, b_f, torch.float32)
    blob_layers = n_layers
    if device is not None:
        blob = blob.to(device)
    blob.hvlm_n_layers = blob_layers
    return blob
