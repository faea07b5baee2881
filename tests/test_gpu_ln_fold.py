"""GPU parity tests of the FOLDED LayerNorms (the tower's default schedule): LN1 / LN2 run inside the QKV / fc1 GEMMs,
   LN(x) W^T + b = rstd * (bf16(x) bf16(gamma*W)^T - mean * c) + (b + W beta),
and the residual GEMMs in front of them emit the bf16 rows and the row statistics from their residual-load epilogue.
Replaces LayerNorm + Linear of HF CLIPEncoderLayer as run by llava/model/multimodal_encoder/clip_encoder.py:39-51.
Every call goes through the C ABI; the reference numbers are the oracle's (oracle/restate.py)."""
import types

import pytest
import torch
import torch.nn.functional as F

import hvlm_b200
from hvlm_b200 import ops
from hvlm_b200.tower import CLIPVisionTower
from hvlm_b200.weights import fold_layernorm
from oracle import restate, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def relmax(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def block_stats(x):
    """(sum, sum of squares) per row and 128-column block: [M,8,2] (fp64 -> fp32)."""
    xb = x.double().reshape(x.shape[0], 8, 128)
    return torch.stack([xb.sum(-1), (xb * xb).sum(-1)], -1).float()


@pytest.mark.parametrize("rows", [1, 7, 257, 2570])
def test_layernorm_with_stats(rows):
    """The tower's pre_layrnorm variant: fp32 LayerNorm output + its bf16 copy + whole-row statistics in block 0."""
    x = synth.gen("lnf.x", (rows, 1024), 2.0, rows, mean=0.7).to(DEV)
    g = synth.gen("lnf.g", (1024,), 0.2, 1, mean=1.0).to(DEV)
    b = synth.gen("lnf.b", (1024,), 0.1, 2).to(DEV)
    out, xb, stats, shift = ops.layernorm_1024_stats(x, g, b)
    ref = F.layer_norm(x.cpu(), (1024,), g.cpu(), b.cpu(), 1e-5)
    assert relmax(out, ref) <= 1e-5
    assert relmax(shift, out.double().mean(-1).float()) <= 1e-5 or float(shift.abs().max()) < 1e-6     # the row mean of `out`
    cen = out - shift[:, None]                                                    # what the fold hands over is centred
    assert torch.equal(xb, cen.to(torch.bfloat16))
    want = block_stats(cen).sum(1)                     # whole row
    assert float((stats[:, 0, 1].cpu() - want[:, 1].cpu()).abs().max() / want[:, 1].abs().max()) <= 1e-5
    assert float(stats[:, 0, 0].abs().max()) <= 2e-3 * float(out.abs().max())       # sum of the centred row ~ 0
    assert float(stats[:, 1:].abs().max()) == 0.0


@pytest.mark.parametrize("M,K", [(1, 1024), (300, 1024), (4099, 1024), (2570, 4096), (25700, 1024), (25700, 4096)])
def test_gemm_resid_stats(M, K):
    """Residual GEMM that feeds a folded LayerNorm (out_proj: K = 1024, fc2: K = 4096; ragged M, the one-tile case, the
    100-frame schedule with its half-tile tail): `hidden` is bit-identical to the TMA reduce-add residual GEMM, the bf16 copy
    is the rounding of `hidden`, the statistics are the per-128-column sums of `hidden`."""
    a = synth.gen("rs.A", (M, K), 1.0, 3).to(torch.bfloat16).to(DEV)
    w = synth.gen("rs.W", (1024, K), K ** -0.5, 3).to(torch.bfloat16).to(DEV)
    bias = synth.gen("rs.b", (1024,), 0.5, 3).to(DEV)
    res = synth.gen("rs.r", (M, 1024), 1.5, 4, mean=0.3).to(DEV)
    h_ref = res.clone()
    ops.gemm(a, w, bias, epilogue="residual", resid=h_ref, out=h_ref)
    h = res.clone()
    xb, stats = ops.gemm_resid_stats(a, w, bias, h)
    assert relmax(h, a.float() @ w.float().t() + bias + res) <= 2e-5
    assert torch.equal(h, h_ref)
    assert torch.equal(xb, h.to(torch.bfloat16))
    want = block_stats(h)
    assert float((stats.cpu() - want.cpu()).abs().max() / want.abs().max()) <= 2e-6
    h2 = res.clone()
    xb2, stats2 = ops.gemm_resid_stats(a, w, bias, h2)
    assert torch.equal(h2, h) and torch.equal(xb2, xb) and torch.equal(stats2, stats)       # bit-reproducible
    # with a per-row centre: same `hidden`, bf16 copy and statistics of the centred rows
    shift = synth.gen("rs.s", (M,), 0.4, 5, mean=0.3).to(DEV)
    h3 = res.clone()
    xb3, stats3 = ops.gemm_resid_stats(a, w, bias, h3, shift=shift)
    assert torch.equal(h3, h)
    cen = h - shift[:, None]
    assert torch.equal(xb3, cen.to(torch.bfloat16))
    want3 = block_stats(cen)
    assert float((stats3.cpu() - want3.cpu()).abs().max() / want3.abs().max()) <= 2e-6


@pytest.mark.parametrize("M,N,epilogue,qkv_hm", [(1, 3072, "bias", True), (300, 3072, "bias", True), (2570, 3072, "bias", True),
                                                 (300, 4096, "quick_gelu", False), (4099, 4096, "quick_gelu", False),
                                                 (356, 1024, "bias", False), (25700, 3072, "bias", True),
                                                 (25700, 4096, "quick_gelu", False)])
def test_gemm_ln_fold(M, N, epilogue, qkv_hm):
    """LayerNorm folded into its consumer GEMM against (a) the same arithmetic in fp32 torch (tight) and (b) the plain
    LayerNorm -> Linear of the reference model in fp32 (the 1e-2 bar of the bf16 regime).  Rows with a mean of ~2 standard
    deviations and per-128-column partial statistics spread over all eight blocks."""
    x = synth.gen("gf.x", (M, 1024), 1.5, M, mean=2.5).to(DEV)
    x[:, 7] += 30.0                                                  # an outlier channel, like real ViT residual streams
    W = synth.gen("gf.W", (N, 1024), 1024 ** -0.5, 5)
    b = synth.gen("gf.b", (N,), 0.3, 5)
    g = synth.gen("gf.g", (1024,), 0.2, 5, mean=1.0)
    beta = synth.gen("gf.beta", (1024,), 0.1, 6)
    w_f, c, b_f = fold_layernorm(W, b, g, beta)
    stats = block_stats(x).to(DEV)
    xb = x.to(torch.bfloat16)
    out = ops.gemm_ln_fold(xb, stats, w_f.to(DEV), c.to(DEV), b_f.to(DEV), epilogue=epilogue, qkv_hm=qkv_hm)
    if qkv_hm:
        assert out.shape == (48, M, 64)
        out = out.permute(1, 0, 2).reshape(M, N)
    act = restate.quick_gelu if epilogue == "quick_gelu" else (lambda t: t)
    xc = x.cpu()
    same = act(restate.folded_layernorm_linear(xc, g, beta, W, b))
    plain = act(F.layer_norm(xc, (1024,), g, beta, 1e-5) @ W.t() + b)
    assert relmax(out, same) <= 6e-3          # bf16 output rounding
    assert relmax(out, plain) <= 1e-2
    # statistics written as one whole-row block (the pre_layrnorm layout) give the same result
    s1 = torch.zeros_like(stats)
    s1[:, 0] = stats.sum(1)
    out1 = ops.gemm_ln_fold(xb, s1, w_f.to(DEV), c.to(DEV), b_f.to(DEV), epilogue=epilogue, qkv_hm=qkv_hm)
    if qkv_hm:
        out1 = out1.permute(1, 0, 2).reshape(M, N)
    assert relmax(out1, out) <= 4e-3
    # the running row mean kept for the next producer: only the tiles of the first column block write it, once per row
    sh = synth.gen("gf.s", (M,), 1.0, 8).to(DEV)
    sh0 = sh.clone()
    ops.gemm_ln_fold(xb, stats, w_f.to(DEV), c.to(DEV), b_f.to(DEV), epilogue=epilogue, qkv_hm=qkv_hm, shift_io=sh)
    assert relmax(sh - sh0, xc.double().mean(-1).float()) <= 1e-5


@pytest.mark.parametrize("dc", [0.0, 50.0, -400.0])
def test_fold_is_robust_to_a_row_offset(dc):
    """What separates the fold from LayerNorm-then-round is a constant added to a row: rounding x to bf16 instead of x - mean
    costs |mean| / std in relative precision, and the one-pass variance cancels.  The producers therefore centre the rows on
    their running mean (any per-row constant is invisible to the LayerNorm): residual GEMM -> folded GEMM with a row offset
    of `dc` standard deviations stays within the bf16 bar of the fp32 LayerNorm -> Linear."""
    M, N = 600, 3072
    res = synth.gen("dc.r", (M, 1024), 1.0, 3).to(DEV) + dc
    a = synth.gen("dc.A", (M, 1024), 1.0, 3).to(torch.bfloat16).to(DEV)
    w = synth.gen("dc.W", (1024, 1024), 0.5 * 1024 ** -0.5, 3).to(torch.bfloat16).to(DEV)
    bias = synth.gen("dc.b", (1024,), 0.1, 3).to(DEV)
    W = synth.gen("dc.W2", (N, 1024), 1024 ** -0.5, 5)
    b = synth.gen("dc.b2", (N,), 0.3, 5)
    g = synth.gen("dc.g", (1024,), 0.2, 5, mean=1.0)
    beta = synth.gen("dc.beta", (1024,), 0.1, 6)
    w_f, c, b_f = fold_layernorm(W, b, g, beta)
    shift = res.mean(-1) + 0.3                       # "the previous mean": close to, not equal to, the new row mean
    h = res.clone()
    xb, stats = ops.gemm_resid_stats(a, w, bias, h, shift=shift)
    sh = shift.clone()
    out = ops.gemm_ln_fold(xb, stats, w_f.to(DEV), c.to(DEV), b_f.to(DEV), shift_io=sh)
    plain = F.layer_norm(h.cpu(), (1024,), g, beta, 1e-5) @ W.t() + b
    assert relmax(out, plain) <= 1e-2
    assert relmax(sh, h.double().mean(-1).float()) <= 1e-4        # the running mean is the row mean again


@pytest.fixture(scope="module")
def tower_strong():
    sd = synth.clip_state_dict(synth.VIT_L14, 0, "strong", n_layers=23)
    tw = CLIPVisionTower("synthetic", types.SimpleNamespace(mm_vision_select_layer=-2), delay_load=True)
    tw.load_model(sd)
    return tw.to(DEV), sd


def test_tower_fold_on_off_agree(tower_strong):
    """The two LayerNorm schedules are the same tower: each within the bf16 bar of the fp32 oracle, within 6e-3 of each
    other, each bit-reproducible; the folded schedule launches two kernels fewer per layer."""
    tw, sd = tower_strong
    px = synth.pixels((3, 3, 224, 224), seed=9).to(DEV)
    prev = ops.vit_set_ln_fold(1)
    try:
        n0 = ops.launch_count()
        h_fold = tw.forward_hidden(px)
        n_fold = ops.launch_count() - n0
        assert torch.equal(h_fold, tw.forward_hidden(px))
        assert ops.vit_set_ln_fold(0) == 1
        n0 = ops.launch_count()
        h_plain = tw.forward_hidden(px)
        n_plain = ops.launch_count() - n0
        assert torch.equal(h_plain, tw.forward_hidden(px))
    finally:
        ops.vit_set_ln_fold(prev)
    assert n_plain - n_fold == 2 * 23, (n_plain, n_fold)
    ref = restate.vit_hidden(px.cpu(), sd, 23)
    assert relmax(h_fold, ref) <= 1e-2 and relmax(h_plain, ref) <= 1e-2
    assert relmax(h_fold, h_plain) <= 6e-3


@pytest.mark.parametrize("n_layers_run", [0, 1, 5])
def test_tower_fold_prefix_of_layers(tower_strong, n_layers_run):
    """A blob packed for 23 layers runs any prefix of them (select_layer variants) through the folded schedule."""
    tw, sd = tower_strong
    px = synth.pixels((2, 3, 224, 224), seed=11)
    hid = ops.vit_l14_hidden(tw.weight_blob, px.to(DEV), n_layers_run)
    emu = restate.vit_hidden(px, sd, n_layers_run, emulate="bf16_fold")
    assert relmax(hid, emu) <= 3e-3
