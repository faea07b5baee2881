"""CPU-only checks of the folded LayerNorm's host side: the algebra the kernels rely on
   LN(x) W^T + b = rstd * (x (gamma*W)^T - mean * c) + (b + W beta),  c[n] = sum_k (gamma*W)[n,k]
(LayerNorm + Linear of HF CLIPEncoderLayer as run by llava/model/multimodal_encoder/clip_encoder.py:39-51), its invariance to a
per-row constant (what lets the producers centre the rows), the operands `pack_vit_weights` writes into the ABI-3 blob, and the
oracle's "bf16_fold" regime against its fp32 regime.  No kernel is launched."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

import hvlm_b200
from hvlm_b200 import _lib as L
from hvlm_b200.weights import fold_layernorm, pack_vit_weights, vit_layout
from oracle import restate, synth


def _operands(N=96, seed=3):
    W = synth.gen("fc.W", (N, 1024), 1024 ** -0.5, seed).double()
    b = synth.gen("fc.b", (N,), 0.3, seed).double()
    g = synth.gen("fc.g", (1024,), 0.2, seed, mean=1.0).double()
    beta = synth.gen("fc.beta", (1024,), 0.1, seed + 1).double()
    return W, b, g, beta


def test_fold_identity_in_fp64():
    W, b, g, beta = _operands()
    x = synth.gen("fc.x", (17, 1024), 2.0, 5, mean=1.5).double()
    ref = F.layer_norm(x, (1024,), g, beta, 1e-5) @ W.t() + b
    mean = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-5)
    Wg = W * g
    got = rstd * (x @ Wg.t() - mean * Wg.sum(-1)) + (b + W @ beta)
    assert float((got - ref).abs().max()) <= 1e-10


def test_fold_is_invariant_to_a_row_constant():
    """x -> x - s_i (any per-row s): mean moves by s, rstd does not move, the result does not move.  This is what allows the
    producers to centre the bf16 rows and the statistics on the running row mean."""
    W, b, g, beta = _operands()
    x = synth.gen("fc.x", (9, 1024), 2.0, 6, mean=40.0).double()
    s = synth.gen("fc.s", (9, 1), 0.5, 7, mean=40.0).double()          # "the previous mean": near the row mean, not equal
    a = restate.folded_layernorm_linear(x.float(), g.float(), beta.float(), W.float(), b.float()).double()
    c = restate.folded_layernorm_linear(x.float(), g.float(), beta.float(), W.float(), b.float(), shift=s.float()).double()
    ref = F.layer_norm(x, (1024,), g, beta, 1e-5) @ W.t() + b
    scale = float(ref.abs().max())
    assert float((c - ref).abs().max()) / scale <= 4e-3          # centred: bf16 operand rounding only
    assert float((a - ref).abs().max()) / scale > 2 * float((c - ref).abs().max()) / scale    # un-centred rows with a mean of 20 std lose precision


def test_fold_layernorm_operands():
    W, b, g, beta = _operands()
    w_f, c, b_f = fold_layernorm(W.float(), b.float(), g.float(), beta.float())
    assert w_f.dtype == torch.bfloat16 and c.dtype == torch.float32 and b_f.dtype == torch.float32
    assert torch.equal(w_f, (W.float() * g.float()).to(torch.bfloat16))              # rounded ONCE, from the fp32 weights
    assert float((c.double() - w_f.double().sum(-1)).abs().max()) <= 1e-6           # sums of the ROUNDED weights
    assert float((b_f.double() - (b + W.float().double() @ beta.float().double())).abs().max()) <= 1e-6


def test_packed_blob_carries_the_fold_operands():
    sd = synth.clip_state_dict(synth.VIT_L14, 0, "hf", n_layers=2)
    blob = pack_vit_weights(sd, n_layers=2)
    lay = vit_layout(2)
    assert blob.numel() == lay.total_bytes

    def view(off, n, dt):
        return blob[int(off): int(off) + n * torch.empty(0, dtype=dt).element_size()].view(dt)
    for l in range(2):
        q = f"vision_model.encoder.layers.{l}."
        f = lay.fold[l]
        w1, b1 = sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"]
        w_f, c, b_f = fold_layernorm(w1, b1, sd[q + "layer_norm2.weight"], sd[q + "layer_norm2.bias"])
        assert torch.equal(view(f.w_fc1_f, 4096 * 1024, torch.bfloat16).view(4096, 1024), w_f)
        assert torch.equal(view(f.c_fc1, 4096, torch.float32), c)
        assert torch.equal(view(f.b_fc1_f, 4096, torch.float32), b_f)
        wq = torch.cat([sd[q + "self_attn.q_proj.weight"] * 0.125, sd[q + "self_attn.k_proj.weight"],
                        sd[q + "self_attn.v_proj.weight"]], 0)
        bq = torch.cat([sd[q + "self_attn.q_proj.bias"] * 0.125, sd[q + "self_attn.k_proj.bias"],
                        sd[q + "self_attn.v_proj.bias"]], 0)
        w_f, c, b_f = fold_layernorm(wq, bq, sd[q + "layer_norm1.weight"], sd[q + "layer_norm1.bias"])
        assert torch.equal(view(f.w_qkv_f, 3072 * 1024, torch.bfloat16).view(3072, 1024), w_f)
        assert torch.equal(view(f.c_qkv, 3072, torch.float32), c)
        assert torch.equal(view(f.b_qkv_f, 3072, torch.float32), b_f)
        # the ABI-2 operands are untouched (HVLM_LN_FOLD=0 still runs from the same blob)
        assert torch.equal(view(lay.layer[l].w_fc1, 4096 * 1024, torch.bfloat16).view(4096, 1024), w1.to(torch.bfloat16))


@pytest.mark.parametrize("profile", ["hf", "strong"])
def test_oracle_bf16_fold_regime_tracks_fp32(profile):
    """The oracle's restatement of the folded arithmetic (2 layers, 1 frame) against its fp32 regime: same bar as the
    LayerNorm-then-round regime."""
    sd = synth.clip_state_dict(synth.VIT_L14, 0, profile, n_layers=2)
    px = synth.pixels((1, 3, 224, 224), seed=4)
    ref = restate.vit_hidden(px, sd, 2)
    fold = restate.vit_hidden(px, sd, 2, emulate="bf16_fold")
    emu = restate.vit_hidden(px, sd, 2, emulate="bf16")
    scale = float(ref.abs().max())
    e_fold, e_emu = float((fold - ref).abs().max()) / scale, float((emu - ref).abs().max()) / scale
    assert e_fold <= 4e-3 and e_emu <= 4e-3 and e_fold <= 2 * e_emu + 1e-4
