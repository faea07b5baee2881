"""GPU parity tests: every CUDA stage against the CPU oracle (oracle/restate.py, pinned to the real reference by
tests/golden) and directly against the golden fixtures.  All calls go through the C ABI (ctypes -> libhvlm_b200.so).

Tolerances (BASELINE.json north_star):
  * ints / bools / copied rows : bit-exact (torch.equal)
  * fp32 pooling               : max|a-b| / max|b| <= 1e-5
  * bf16-GEMM-backed outputs   : max|a-b| / max|b_fp32| <= 1e-2 against the fp32 oracle
"""
import types

import numpy as np
import pytest
import torch

import hvlm_b200
from hvlm_b200 import _lib as L
from hvlm_b200 import arch, ops
from hvlm_b200.tower import CLIPVisionTower
from oracle import restate, synth
from oracle.make_golden import SMALL, SMALL_D, small_parts

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_POOL = 1e-5
TOL_BF16 = 1e-2


def T(a):
    return torch.from_numpy(np.asarray(a))


def relmax(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ------------------------------------------------------------------ GEMM (tcgen05)
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 1024, 1024), (356, 4096, 1024), (356, 5120, 1024),
                                   (1000, 1024, 4096), (200, 128, 64), (2571, 3072, 1024), (1024, 1024, 1424)])
def test_gemm_epilogues(M, N, K):
    a = synth.gen(f"A{M}", (M, K), 1.0, 1).to(torch.bfloat16).to(DEV)
    w = synth.gen(f"W{N}", (N, K), K ** -0.5, 1).to(torch.bfloat16).to(DEV)
    bias = synth.gen(f"b{N}", (N,), 0.5, 1).to(DEV)
    res = synth.gen(f"r{M}", (M, N), 1.0, 2).to(DEV)
    ref = a.float() @ w.float().t() + bias                     # same bf16 operands, fp32 math
    assert relmax(ops.gemm(a, w, bias, out_dtype=torch.float32), ref) <= 2e-5
    assert relmax(ops.gemm(a, w, None, out_dtype=torch.float32), ref - bias) <= 2e-5
    assert relmax(ops.gemm(a, w, bias, out_dtype=torch.bfloat16), ref) <= 5e-3
    assert relmax(ops.gemm(a, w, bias, epilogue="quick_gelu", out_dtype=torch.float32), restate.quick_gelu(ref)) <= 2e-5
    assert relmax(ops.gemm(a, w, bias, epilogue="residual", resid=res, out_dtype=torch.float32), ref + res) <= 2e-5
    # in-place residual (how the tower uses it)
    r2 = res.clone()
    ops.gemm(a, w, bias, epilogue="residual", resid=r2, out=r2)
    assert relmax(r2, ref + res) <= 2e-5


def test_gemm_randomised_shapes():
    """Seeded fuzz of the tcgen05 GEMMs: ragged M (partial row blocks, single rows), every N class (128-multiples go to the
    1-CTA kernel, 256-multiples to the 2-CTA kernel with whole / half-tile tails), K not a multiple of the 64-wide stage."""
    rng = np.random.RandomState(7)
    for case in range(24):
        M = int(rng.choice([1, 2, 127, 128, 129, 255, 257, 300, 511, 777, 1025, 2311, 4099]))
        N = int(rng.choice([128, 256, 384, 512, 768, 1024, 1280, 2048, 3072]))
        K = int(rng.choice([8, 64, 72, 128, 200, 320, 640, 1024, 1096]))
        a = synth.gen(f"fz.A{case}", (M, K), 1.0, case).to(torch.bfloat16).to(DEV)
        w = synth.gen(f"fz.W{case}", (N, K), K ** -0.5, case).to(torch.bfloat16).to(DEV)
        bias = synth.gen(f"fz.b{case}", (N,), 0.5, case).to(DEV)
        res = synth.gen(f"fz.r{case}", (M, N), 1.0, case).to(DEV)
        ref = a.float() @ w.float().t() + bias
        tag = (case, M, N, K)
        assert relmax(ops.gemm(a, w, bias, out_dtype=torch.float32), ref) <= 2e-5, tag
        assert relmax(ops.gemm(a, w, bias, out_dtype=torch.bfloat16), ref) <= 5e-3, tag
        assert relmax(ops.gemm(a, w, bias, epilogue="quick_gelu", out_dtype=torch.bfloat16), restate.quick_gelu(ref)) <= 6e-3, tag
        r2 = res.clone()
        ops.gemm(a, w, bias, epilogue="residual", resid=r2, out=r2)
        assert relmax(r2, ref + res) <= 2e-5, tag


def test_gemm_rejects_bad_arguments():
    a = torch.zeros(8, 64, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(L.HvlmError):
        ops.gemm(a, torch.zeros(100, 64, dtype=torch.bfloat16, device=DEV))      # N % 128 != 0
    with pytest.raises(AssertionError):
        ops.gemm(a.float(), a.float())


def test_projector_autograd_matches_torch():
    M, D = 712, 4096
    x = synth.gen("proj.x", (M, 1024), 1.0, 3).to(DEV)
    ps = synth.projector_state(D)
    w = ps["mm_projector.weight"].to(DEV).requires_grad_(True)
    b = ps["mm_projector.bias"].to(DEV).requires_grad_(True)
    dy = synth.gen("proj.dy", (M, D), 1.0, 3).to(DEV)
    y = ops.linear(x, w.to(torch.bfloat16), b, True)
    y.backward(dy)
    xr = x.to(torch.bfloat16).float()
    wr = w.detach().to(torch.bfloat16).float()
    assert relmax(y, xr @ wr.t() + b.detach()) <= 2e-5
    assert relmax(w.grad, dy.t() @ x) <= TOL_BF16              # fp32 oracle
    assert relmax(b.grad, dy.sum(0)) <= 1e-5


# ------------------------------------------------------------------ LayerNorm
def test_layernorm():
    x = synth.gen("ln.x", (1030, 1024), 2.0, 1, 0.3)
    g = synth.gen("ln.g", (1024,), 0.2, 1, 1.0)
    b = synth.gen("ln.b", (1024,), 0.1, 1)
    ref = torch.nn.functional.layer_norm(x, (1024,), g, b, 1e-5)
    assert relmax(ops.layernorm_1024(x.to(DEV), g.to(DEV), b.to(DEV), torch.float32), ref) <= 1e-5
    assert relmax(ops.layernorm_1024(x.to(DEV), g.to(DEV), b.to(DEV), torch.bfloat16), ref) <= 5e-3


# ------------------------------------------------------------------ pooling
@pytest.mark.parametrize("t,d", [(100, 32), (10, 16), (2, 8), (5, 8)])
@pytest.mark.parametrize("mode", ["temporal_spatial_pool", "spatial_pool"])
def test_pool_matches_reference_fixture(golden, t, d, mode):
    g = golden(f"pool_{mode}_t{t}")
    tok = synth.gen(f"pooltok{t}", (2, t, 256, d), 1.0, seed=5)
    out = ops.pool_tokens(tok.to(DEV), mode)
    assert relmax(out, T(g["out"])) <= TOL_POOL


@pytest.mark.parametrize("mode", ["temporal_spatial_pool", "spatial_pool", "temporal", "spatial", "temporal_spatial"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pool_modes_vs_oracle(mode, dtype):
    tok = synth.gen("pool.modes", (2, 7, 256, 1024), 1.0, 2).to(dtype)
    ref = restate.pool_tokens(tok.float(), mode)
    assert relmax(ops.pool_tokens(tok.to(DEV), mode, torch.float32), ref) <= TOL_POOL


def test_pool_and_gather_randomised_against_oracle():
    """Seeded fuzz: pooling over odd frame counts (t < 4 duplicates the selected frames, lita_arch.py:54-58), channel
    counts and modes; <hand_traj> gather with 0 or 4 hand tokens at random positions (incl. position 0 and the end)."""
    rng = np.random.RandomState(99)
    for case in range(12):
        b = int(rng.randint(1, 4))
        t = int(rng.choice([1, 2, 3, 4, 5, 9, 17]))
        c = int(rng.choice([64, 320, 1024]))
        mode = str(rng.choice(["temporal_spatial_pool", "spatial_pool", "temporal", "spatial", "temporal_spatial"]))
        tok = synth.gen(f"fuzz.pool{case}", (b, t, 256, c), 1.0, case)
        ref = restate.pool_tokens(tok, mode)
        out = ops.pool_tokens(tok.to(DEV), mode, torch.float32)
        assert out.shape == ref.shape and relmax(out, ref) <= TOL_POOL, (case, b, t, c, mode)
    for case in range(20):
        B = int(rng.randint(1, 6))
        Lh = int(rng.randint(6, 90))
        D = int(rng.choice([16, 64, 4096]))
        hidden = synth.gen(f"fuzz.gh{case}", (B, Lh, D), 1.0, case).to(torch.bfloat16)
        labels = torch.from_numpy(rng.randint(-100, 32000, size=(B, Lh))).long()
        for bb in range(B):
            if rng.rand() < 0.7:
                pos = rng.choice(np.arange(1, Lh), size=4, replace=False)   # label[0] has no predictor row
                labels[bb, torch.from_numpy(np.sort(pos)).long()] = 32100
        out, valid, rows, counts = ops.hand_gather(hidden.to(DEV), labels.to(DEV), 32100)
        ro, rv, rr = restate.gather_hand_traj(hidden, labels)
        assert torch.equal(out.cpu(), ro) and torch.equal(valid.cpu(), rv) and torch.equal(rows.cpu(), rr), case


def test_pool_reads_tower_layout_in_place():
    hid = synth.gen("hid", (6, 257, 1024), 1.0, 4)
    ref = restate.pool_tokens(hid[:, 1:].reshape(2, 3, 256, 1024), "temporal_spatial_pool")
    assert relmax(ops.pool_slowfast(hid.to(DEV), 2, 3, 257, 1, 0, False), ref) <= TOL_POOL


def test_pool_backward_fixture_and_autograd(golden):
    g = golden("pool_bwd_t10")
    dout = synth.gen("pooldout_bwd", (1, 266, 8), 1.0, seed=6)
    assert relmax(ops.pool_slowfast_bwd(dout.to(DEV), 10, 0, False), T(g["dtok"])) <= TOL_POOL
    tok = synth.gen("pooltok_bwd", (1, 10, 256, 8), 1.0, seed=6).to(DEV).requires_grad_(True)
    out = ops.pool_tokens(tok, "temporal_spatial_pool")
    (out * dout.to(DEV)).sum().backward()
    assert relmax(tok.grad, T(g["dtok"])) <= TOL_POOL


def test_pool_full_size_properties():
    """BASELINE size (fp32, 100 frames, C=4096): linearity and mean-of-means, size-independent checks."""
    a = torch.randn(1, 100, 256, 4096, device=DEV)
    b = torch.randn(1, 100, 256, 4096, device=DEV)
    pa, pb = ops.pool_tokens(a, "temporal_spatial_pool"), ops.pool_tokens(b, "temporal_spatial_pool")
    pab = ops.pool_tokens(a + 2 * b, "temporal_spatial_pool")
    assert pa.shape == (1, 356, 4096)
    assert relmax(pab, pa + 2 * pb) <= 1e-5
    # each selected frame's 64 slow tokens average to that frame's fast token
    for k, f in enumerate([0, 33, 66, 99]):
        assert relmax(pa[0, 100 + 64 * k: 100 + 64 * (k + 1)].mean(0), pa[0, f]) <= 1e-5
    assert relmax(pa[0, :100], a[0].mean(1)) <= 1e-5


# ------------------------------------------------------------------ gather
@pytest.mark.parametrize("name,shape,seedname,seed", [("gather_toy", None, None, None),
                                                      ("gather_rand", (4, 40, 64), "gather_hidden", 41),
                                                      ("gather_pos0", (1, 12, 16), "gather_hidden0", 42)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gather_matches_reference_fixture(golden, name, shape, seedname, seed, dtype):
    g = golden(name)
    labels = T(g["labels"])
    hidden = (torch.arange(2 * 10 * 8, dtype=torch.float32).reshape(2, 10, 8) if shape is None
              else synth.gen(seedname, shape, 1.0, seed=seed)).to(dtype)
    fv = torch.ones(labels.shape[0], 2, dtype=torch.bool, device=DEV)
    out, valid = arch.gather_hand_traj_states(hidden.to(DEV), labels.to(DEV), future_valid=fv)
    assert torch.equal(out.cpu().float(), T(g["out"]).to(dtype).float())
    assert torch.equal(fv.cpu(), T(g["future_valid"]))


def test_gather_wrong_count_raises_like_reference():
    labels = torch.full((1, 10), -100, dtype=torch.int64)
    labels[0, 3:6] = 32100
    with pytest.raises(RuntimeError, match="is invalid for input of size"):
        arch.gather_hand_traj_states(torch.zeros(1, 10, 8, device=DEV), labels.to(DEV))


def test_gather_backward_and_step():
    hidden = synth.gen("gh", (3, 50, 4096), 1.0, 7).to(DEV).requires_grad_(True)
    labels = torch.full((3, 50), -100, dtype=torch.int64)
    labels[0, 40:44] = 32100
    labels[2, [5, 6, 30, 49]] = 32100
    out, valid, rows, counts = ops.hand_gather(hidden, labels.to(DEV), 32100)
    dout = synth.gen("gdout", tuple(out.shape), 1.0, 7)
    out.backward(dout.to(DEV))
    ro, rv, rr = restate.gather_hand_traj(hidden.detach().cpu(), labels)
    assert torch.equal(out.detach().cpu(), ro) and torch.equal(rows.cpu(), rr)
    assert torch.equal(hidden.grad.cpu(), restate.gather_hand_traj_backward(dout, rr, 50))
    last = synth.gen("glast", (2, 4096), 1.0, 8)
    assert torch.equal(ops.hand_gather_step(last.to(DEV)).cpu(), restate.gather_hand_traj_step(last))


# ------------------------------------------------------------------ trajectory head (8f item 4)
def _traj_module(Dc, seed, dtype):
    from hvlm_b200.traj_decoder import CVAETrajDecoder
    sd = synth.traj_cvae_state(Dc, seed=seed)
    dec = CVAETrajDecoder(token_dim=Dc)
    dec.load_state_dict(sd, strict=True)            # the reference's own key names
    return dec.to(DEV).to(dtype), sd


def test_traj_head_golden():
    """fp32 weights, the reference's noise from the fixture: the fused kernel vs the reference's own output."""
    dec, sd = _traj_module(32, 3, torch.float32)
    g = np.load("tests/golden/traj_infer.npz")
    emb = synth.gen("traj_emb", (3, 2, 4, 32), 1.0, seed=51).to(DEV)
    out = dec.inference(pred_hand_embeddings=emb, z=T(g["z"]).to(DEV))
    assert out.shape == (3, 2, 4, 2)
    assert relmax(out, T(g["out"])) <= 3e-5, relmax(out, T(g["out"]))
    g = np.load("tests/golden/traj_step.npz")
    hl = synth.gen("traj_hidden_last", (1, 64), 1.0, seed=52).to(DEV)
    out = dec.inference_step(hl, z=T(g["z"]).to(DEV))
    assert out.shape == (1, 2, 2)
    assert relmax(out, T(g["out"])) <= 3e-5, relmax(out, T(g["out"]))


@pytest.mark.parametrize("D,B", [(4096, 1), (5120, 1), (4096, 3)])
def test_traj_head_step_full_size(D, B):
    """7B / 13B widths in bf16: fused gather+decode vs the fp32 oracle on the same (bf16-rounded) operands; internally
    drawn noise must reproduce torch.randn under the same seed; bit-reproducible run to run."""
    Dc = D // 2
    dec, sd = _traj_module(Dc, 5, torch.bfloat16)
    sd16 = {k: v.to(torch.bfloat16).float() for k, v in sd.items()}
    hl = synth.gen("traj_hl", (B, D), 1.0, seed=9).to(torch.bfloat16)
    torch.manual_seed(99)
    out = dec.inference_step(hl.to(DEV))
    torch.manual_seed(99)
    z = (2.0 * torch.randn([2 * B, 256], device=DEV)).to(torch.bfloat16)
    ref = restate.traj_decode_step(hl.float(), z.float().cpu(), sd16)
    assert out.dtype == torch.bfloat16 and out.shape == (B, 2, 2)
    assert relmax(out, ref) <= TOL_BF16
    # fp32 result of the kernel itself (before the cast back to bf16) is tight
    d = dec.hand_traj_decoder.cvae.dec_MLP
    raw = ops.traj_decode(hl.to(DEV), z, d[0].weight, d[0].bias, d[2].weight, d[2].bias, interleaved=True)
    assert relmax(raw.reshape(B, 2, 2), ref) <= 1e-4
    raw2 = ops.traj_decode(hl.to(DEV), z, d[0].weight, d[0].bias, d[2].weight, d[2].bias, interleaved=True)
    assert torch.equal(raw, raw2)
    # the two-call form (gather, then decode) gives the same bits
    e = ops.hand_gather_step(hl.to(DEV))
    raw3 = ops.traj_decode(e.reshape(-1, Dc), z, d[0].weight, d[0].bias, d[2].weight, d[2].bias)
    assert torch.equal(raw, raw3)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 3e-5), (torch.bfloat16, TOL_BF16)])
def test_traj_mlp_decoder(dtype, tol):
    """MLPTrajDecoder.inference (traj_decoder 'MLP'): three skinny linears, against the reference fixture (fp32, small) and
    the fp32 oracle at the 7B width on bf16-rounded operands."""
    from hvlm_b200.traj_decoder import MLPTrajDecoder
    if dtype == torch.float32:
        g = np.load("tests/golden/traj_mlp_infer.npz")
        sd = synth.traj_mlp_state(32, seed=4)
        dec = MLPTrajDecoder(token_dim=32)
        dec.load_state_dict(sd, strict=True)
        emb = synth.gen("trajmlp_emb", (3, 2, 4, 32), 1.0, seed=53)
        out = dec.to(DEV).inference(pred_hand_embeddings=emb.to(DEV))
        assert out.shape == (3, 2, 4, 2) and relmax(out, T(g["out"])) <= tol
        return
    Dc = 2048
    sd = synth.traj_mlp_state(Dc, seed=7)
    dec = MLPTrajDecoder(token_dim=Dc)
    dec.load_state_dict(sd, strict=True)
    dec = dec.to(DEV).to(dtype)
    hl = synth.gen("trajmlp_hl", (2, 2 * Dc), 1.0, seed=8).to(dtype)
    out = dec.inference_step(hl.to(DEV))
    sd16 = {k: v.to(dtype).float() for k, v in sd.items()}
    ref = restate.traj_mlp_inference(restate.gather_hand_traj_step(hl.float()), sd16).squeeze(2)
    assert out.shape == (2, 2, 2) and out.dtype == dtype
    assert relmax(out, ref) <= tol
    with pytest.raises(NotImplementedError):
        dec(pred_hand_embeddings=hl)


def test_traj_head_many_rows_and_errors():
    dec, sd = _traj_module(64, 6, torch.float32)
    emb = synth.gen("traj_emb2", (5, 2, 3, 64), 1.0, seed=10)      # R = 30: several row chunks, ragged tail
    z = synth.gen("traj_z2", (30, 256), 2.0, seed=10)
    out = dec.inference(pred_hand_embeddings=emb.to(DEV), z=z.to(DEV))
    assert relmax(out, restate.traj_decoder_inference(emb, z, sd)) <= 1e-5
    # more rows than one launch covers (4096): the C entry splits the rows over launches
    emb = synth.gen("traj_emb3", (2051, 2, 1, 64), 1.0, seed=11)
    z = synth.gen("traj_z3", (4102, 256), 2.0, seed=11)
    out = dec.inference(pred_hand_embeddings=emb.to(DEV), z=z.to(DEV))
    assert relmax(out, restate.traj_decoder_inference(emb, z, sd)) <= 1e-5
    emb = synth.gen("traj_emb", (3, 2, 4, 64), 1.0, seed=51)
    with pytest.raises(AssertionError):
        dec.inference(pred_hand_embeddings=emb[..., :32].to(DEV))
    with pytest.raises(RuntimeError):
        dec.inference(pred_hand_embeddings=emb)                     # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        dec(pred_hand_embeddings=emb.to(DEV))


# ------------------------------------------------------------------ splice
class _Inner(torch.nn.Module):
    def __init__(self, tower, proj, emb):
        super().__init__()
        self.vision_tower, self.mm_projector, self.embed_tokens = tower, proj, emb

    def get_vision_tower(self):
        return self.vision_tower


def _host(mixin, tower, proj, emb, config, B):
    class Host(torch.nn.Module, mixin):
        def __init__(self):
            super().__init__()
            self.model = _Inner(tower, proj, emb)
            self.config = config
            self.token_dim, self.B = proj.out_features, B

        def get_model(self):
            return self.model
    return Host().to(DEV)


@pytest.fixture(scope="module")
def small():
    sd, proj, emb = small_parts()
    return sd, proj, emb


def _small_visual(small, px, mode="temporal_spatial_pool"):
    sd, proj, _ = small
    return restate.pipeline(px, sd, proj.weight.data, proj.bias.data, mode, select_layer=-2, cfg=SMALL)[0]


def _splice_cuda(variant, ids, mask, labels, vis, table, fh=None, is_eval=False, static=False):
    host = types.SimpleNamespace(
        get_model=lambda: types.SimpleNamespace(embed_tokens=types.SimpleNamespace(weight=table.to(DEV))),
        config=types.SimpleNamespace(hvlm_static_splice=static))
    dv = lambda x: None if x is None else x.to(DEV)
    return arch.splice_tokens(host, variant, dv(ids), dv(mask), dv(labels), dv(vis), None, dv(fh), is_eval)


@pytest.mark.parametrize("name,is_eval", [
    ("splice_hvlm_train_b3", False), ("splice_hvlm_train_padded", False), ("splice_hvlm_2hand", False),
    ("splice_hvlm_0hand", False), ("splice_hvlm_ragged", False), ("splice_hvlm_eval_hands", True),
    ("splice_hvlm_eval_nohands", True), ("splice_hvlm_empty_tail", False), ("splice_hvlm_two_images", False)])
def test_splice_handsonvlm_vs_reference_fixture(golden, small, name, is_eval):
    g = golden(name)
    ids = T(g["ids"])
    px = synth.pixels((ids.shape[0], int(g["t"]), 3, 224, 224), seed=int(g["px_seed"]))
    vis = _small_visual(small, px)                       # visual tokens from the oracle: isolates the splice
    mask = T(g["in_mask"]) if "in_mask" in g.files else None
    labels = T(g["in_labels"]) if "in_labels" in g.files else None
    fh = T(g["future_hands"]) if "future_hands" in g.files else None
    m2, e2, l2 = _splice_cuda(L.SPLICE_HANDSONVLM, ids, mask, labels, vis, small[2].weight.data, fh, is_eval)
    ge = T(g["embeds"])
    assert e2.shape == ge.shape
    assert relmax(e2, ge) <= 2e-5
    # copied rows are bit-exact against the oracle's spliced tensor
    rm, re_, rl = restate.splice(ids, mask, labels, vis, small[2].weight.data, "handsonvlm", future_hands=fh,
                                 is_evaluate=is_eval)
    diff_rows = (e2.cpu() != re_).any(-1).sum().item()
    assert diff_rows <= 4 * ids.shape[0]                 # only <hand_traj> rows carry fp adds
    if "labels" in g.files:
        assert torch.equal(l2.cpu(), T(g["labels"]))
    else:
        assert l2 is None
    if "mask" in g.files:
        assert str(m2.dtype) == str(g["mask_dtype"]) and torch.equal(m2.cpu(), T(g["mask"]))
    else:
        assert m2 is None


@pytest.mark.parametrize("name,pxshape,pxseed,cfg", [
    ("splice_llava_cfg1", (1, 3, 224, 224), 12, "image"), ("splice_llava_ragged", (2, 3, 224, 224), 13, "image"),
    ("splice_llava_two_images", (2, 3, 224, 224), 14, "image"), ("splice_llava_video", (1, 4, 3, 224, 224), 15, "video")])
def test_splice_llava_vs_reference_fixture(golden, small, name, pxshape, pxseed, cfg):
    g = golden(name)
    sd, proj, emb = small
    px = synth.pixels(pxshape, seed=pxseed)
    if cfg == "image":
        vis = restate.project(restate.tower_forward(px, sd, -2, SMALL), proj.weight.data, proj.bias.data)
    else:
        vis = _small_visual(small, px)
    m2, e2, l2 = _splice_cuda(L.SPLICE_LLAVA, T(g["ids"]), T(g["in_mask"]), T(g["in_labels"]), vis, emb.weight.data)
    # visual rows come from the CPU oracle run on THIS host (thread-count dependent rounding vs the fixture);
    # every row is a pure copy of its source, which is checked bit-exactly against the oracle's own splice
    assert relmax(e2, T(g["embeds"])) <= 2e-5
    rm, re_, rl = restate.splice(T(g["ids"]), T(g["in_mask"]), T(g["in_labels"]), vis, emb.weight.data, "llava")
    assert torch.equal(e2.cpu(), re_)
    assert torch.equal(l2.cpu(), T(g["labels"]))
    assert m2.dtype == torch.bool and torch.equal(m2.cpu(), T(g["mask"]))


def test_splice_llava_im_start_end_variant(golden, small):
    """tune_mm_mlp_adapter + mm_use_im_start_end (llava_arch.py:146-161) through the drop-in method: labels / mask
    bit-exact against the reference fixture, embeddings exact copies, and only the <im_start>/<im_end> rows of the
    embedding table receive gradient (the reference detaches every other text segment)."""
    g = golden("splice_llava_im_start_end")
    sd, proj, emb = small
    feats = restate.tower_forward(synth.pixels((2, 3, 224, 224), seed=16), sd, -2, SMALL)
    vis = restate.project(feats, proj.weight.data, proj.bias.data)
    table = emb.weight.data.clone().to(DEV).requires_grad_(True)
    host = types.SimpleNamespace(
        get_model=lambda: types.SimpleNamespace(embed_tokens=types.SimpleNamespace(weight=table)),
        config=types.SimpleNamespace(hvlm_static_splice=False))
    ids = T(g["ids"]).to(DEV)
    m2, e2, l2 = arch.splice_tokens(host, L.SPLICE_LLAVA, ids, T(g["in_mask"]).to(DEV), T(g["in_labels"]).to(DEV),
                                    vis.to(DEV), im_start_end=True)
    assert torch.equal(l2.cpu(), T(g["labels"])) and torch.equal(m2.cpu(), T(g["mask"]))
    assert relmax(e2, T(g["embeds"])) <= 2e-5
    ro = restate.splice(T(g["ids"]), T(g["in_mask"]), T(g["in_labels"]), vis, emb.weight.data, "llava", im_start_end=True)
    assert torch.equal(e2.detach().cpu(), ro[1])
    dout = synth.gen("ise.dout", tuple(e2.shape), 1.0, seed=16)
    e2.backward(dout.to(DEV))
    gt = table.grad.cpu()
    rows = torch.nonzero(gt.abs().sum(1) > 0).flatten()
    assert torch.equal(rows, T(g["grad_rows"]))
    assert relmax(gt[rows], T(g["grad_vals"])) <= 1e-6
    # and the whole method honours the config flags
    class H(torch.nn.Module, arch.LlavaMetaForCausalLM):
        def get_model(self):
            return None
    assert "mm_use_im_start_end" in arch.LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal.__doc__


def test_splice_hvlm_im_start_end_variant(golden, small):
    """HandsOnVLM splice with tune_mm_mlp_adapter + mm_use_im_start_end (handsonvlm.py:263-286,343-344)."""
    g = golden("splice_hvlm_im_start_end")
    sd, proj, emb = small
    vis = _small_visual(small, synth.pixels((2, int(g["t"]), 3, 224, 224), seed=17))
    host = types.SimpleNamespace(
        get_model=lambda: types.SimpleNamespace(embed_tokens=types.SimpleNamespace(weight=emb.weight.data.to(DEV))),
        config=types.SimpleNamespace(hvlm_static_splice=False))
    m2, e2, l2 = arch.splice_tokens(host, L.SPLICE_HANDSONVLM, T(g["ids"]).to(DEV), T(g["in_mask"]).to(DEV),
                                    T(g["in_labels"]).to(DEV), vis.to(DEV), None, None, True, im_start_end=True)
    assert torch.equal(l2.cpu(), T(g["labels"]))
    assert m2.dtype == torch.bool and torch.equal(m2.cpu(), T(g["mask"]))
    assert relmax(e2, T(g["embeds"])) <= 2e-5
    ro = restate.splice(T(g["ids"]), T(g["in_mask"]), T(g["in_labels"]), vis, emb.weight.data, "handsonvlm",
                        im_start_end=True)
    assert torch.equal(e2.cpu(), ro[1])


def test_splice_llava_list_of_image_groups(golden, small):
    """Ragged visual token blocks (list path of images_to_tokens, llava_arch.py:95-106) through hvlm_splice_plan_ragged:
    labels / mask bit-exact against the reference fixture, rows exact copies, gradient reaches every block."""
    g = golden("splice_llava_list_ragged")
    sd, proj, emb = small
    blocks = []
    for n, sd_ in zip(g["group_sizes"], (18, 19, 20)):
        feats = restate.tower_forward(synth.pixels((int(n), 3, 224, 224), seed=sd_), sd, -2, SMALL)
        blocks.append(restate.project(feats, proj.weight.data, proj.bias.data).reshape(-1, SMALL_D))
    host = types.SimpleNamespace(
        get_model=lambda: types.SimpleNamespace(embed_tokens=types.SimpleNamespace(weight=emb.weight.data.to(DEV))),
        config=types.SimpleNamespace(hvlm_static_splice=True))       # static mode is ignored for ragged blocks
    dblocks = [b.to(DEV).requires_grad_(True) for b in blocks]
    m2, e2, l2 = arch.splice_tokens(host, L.SPLICE_LLAVA, T(g["ids"]).to(DEV), T(g["in_mask"]).to(DEV),
                                    T(g["in_labels"]).to(DEV), dblocks)
    assert e2.shape == tuple(g["embeds"].shape)
    assert torch.equal(l2.cpu(), T(g["labels"])) and torch.equal(m2.cpu(), T(g["mask"]))
    assert relmax(e2, T(g["embeds"])) <= 2e-5
    ro = restate.splice(T(g["ids"]), T(g["in_mask"]), T(g["in_labels"]), blocks, emb.weight.data, "llava")
    assert torch.equal(e2.detach().cpu(), ro[1])
    # backward: sample 0 uses block 0, sample 1 block 1; block 2 belongs to the sample without an image token
    e2.sum().backward()
    assert float(dblocks[0].grad.min()) == 1.0 and float(dblocks[1].grad.min()) == 1.0
    assert float(dblocks[2].grad.abs().max()) == 0.0


def test_splice_randomised_against_oracle():
    """Seeded fuzz of both splice variants: random batch sizes, lengths, 0-3 image tokens per sample at random positions
    (including first / last / adjacent), random masks and labels, uniform and per-slot (ragged) visual token counts.
    Everything is a copy or an integer, so the CUDA result must equal the oracle bit for bit."""
    rng = np.random.RandomState(1234)
    D = 64
    table = synth.gen("fuzz.table", (500, D), 1.0, 3)
    for case in range(40):
        B = int(rng.randint(1, 6))
        T = int(rng.randint(2, 70))
        ragged_blocks = bool(case % 3 == 2)
        ids = torch.from_numpy(rng.randint(1, 500, size=(B, T))).long()
        n_slots = 0
        for b in range(B):
            k = int(rng.randint(0, 4)) if T >= 4 else int(rng.randint(0, 2))
            pos = rng.choice(T, size=min(k, T), replace=False)
            ids[b, torch.from_numpy(pos).long()] = -200
            n_slots += max(len(pos), 1)
        mask = torch.from_numpy(rng.rand(B, T) > 0.2)
        labels = torch.from_numpy(rng.randint(-100, 500, size=(B, T))).long()
        if ragged_blocks:
            vis = [synth.gen(f"fuzz.v{case}.{g}", (int(rng.randint(1, 20)), D), 1.0, case) for g in range(n_slots)]
            dvis = [v.to(DEV) for v in vis]
        else:
            Nv = int(rng.choice([1, 5, 17]))
            vis = synth.gen(f"fuzz.v{case}", (n_slots, Nv, D), 1.0, case)
            dvis = vis.to(DEV)
        for variant, name in ((L.SPLICE_LLAVA, "llava"), (L.SPLICE_HANDSONVLM, "handsonvlm")):
            host = types.SimpleNamespace(
                get_model=lambda: types.SimpleNamespace(embed_tokens=types.SimpleNamespace(weight=table.to(DEV))),
                config=types.SimpleNamespace(hvlm_static_splice=False))
            m2, e2, l2 = arch.splice_tokens(host, variant, ids.to(DEV), mask.to(DEV), labels.to(DEV), dvis, None, None, True)
            rm, re_, rl = restate.splice(ids, mask, labels, vis, table, name, is_evaluate=True)
            assert torch.equal(e2.cpu(), re_), (case, name)
            assert torch.equal(l2.cpu(), rl), (case, name)
            if m2.dtype == torch.bool:
                assert torch.equal(m2.cpu(), rm), (case, name)
            else:       # HandsOnVLM ragged-batch quirk: int64 mask padded with -100 (handsonvlm.py:441)
                assert m2.dtype == torch.int64 and torch.equal(m2.cpu(), rm.to(torch.int64)), (case, name)


def test_splice_static_mode_no_sync_and_status_flag():
    D = 256
    table = synth.embed_table(D)
    vis = synth.gen("vis", (2, 356, D), 1.0, 5)
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=2, seed=3)
    m2, e2, l2 = _splice_cuda(L.SPLICE_HANDSONVLM, ids, mask, labels, vis, table, fh, static=True)
    rm, re_, rl = restate.splice(ids, mask, labels, vis, table, "handsonvlm", future_hands=fh)
    assert torch.equal(l2.cpu(), rl) and torch.equal(m2.cpu(), rm) and relmax(e2, re_) <= 2e-5
    assert l2[0, 411:415].tolist() == [32100] * 4 and e2.shape == (2, 417, D)
    # bad token id -> IndexError like nn.Embedding
    bad = ids.clone()
    bad[0, 3] = 40000
    with pytest.raises(IndexError):
        _splice_cuda(L.SPLICE_HANDSONVLM, bad, mask, labels, vis, table, fh)
    # more hand tokens than future_hands points in eval mode -> AssertionError like the reference
    with pytest.raises(AssertionError):
        _splice_cuda(L.SPLICE_HANDSONVLM, ids, mask, labels, vis, table, fh[:, :, :3], is_eval=True)


def test_splice_backward_matches_oracle():
    D, B = 128, 2
    table = synth.embed_table(D).to(DEV).requires_grad_(True)
    vis = synth.gen("vis", (B, 356, D), 1.0, 5).to(DEV).requires_grad_(True)
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=B, seed=4)
    host = types.SimpleNamespace(
        get_model=lambda: types.SimpleNamespace(embed_tokens=types.SimpleNamespace(weight=table)),
        config=types.SimpleNamespace())
    m2, e2, l2 = arch.splice_tokens(host, L.SPLICE_HANDSONVLM, ids.to(DEV), mask.to(DEV), labels.to(DEV), vis, None,
                                    fh.to(DEV), False)
    de = synth.gen("de", tuple(e2.shape), 1.0, 6)
    e2.backward(de.to(DEV))
    dv, dt = restate.splice_backward(de, ids, 356, B, table.shape[0])
    assert torch.equal(vis.grad.cpu(), dv)
    assert relmax(table.grad, dt) <= 1e-6                 # atomics: order-dependent fp32 sums


# ------------------------------------------------------------------ ViT stages
def test_qkv_and_attention_stage():
    """QKV GEMM with the column-block-major epilogue (3-D TMA stores) + attention (tensor-core
    256x256 block, token-256 scores from the N=16 MMAs, 257th query row on CUDA cores) against fp32 torch.
    F = 20 / 40 give 320 / 640 (frame, head) items on 296 persistent CTAs: the multi-item path (next item's loads,
    token-256 blocks and first S tile issued behind the current item) is exercised with ragged item counts per CTA."""
    for F in (1, 3, 7, 20, 40):
        y = synth.gen("attn.y", (F * 257, 1024), 1.0, F).to(torch.bfloat16)
        w = synth.gen("attn.w", (3072, 1024), 1024 ** -0.5, 1).to(torch.bfloat16)
        b = synth.gen("attn.b", (3072,), 0.1, 1)
        qkv = ops.vit_qkv(y.to(DEV), w.to(DEV), b.to(DEV), F)
        ref = (y.float() @ w.float().t() + b).reshape(F * 257, 48, 64).permute(1, 0, 2)     # [48, M, 64]
        assert qkv.shape == (48, F * 257, 64) and relmax(qkv, ref) <= 5e-3
        for scale in (0.125, 1.0):
            q2 = qkv.clone()
            q2[:16] = (q2[:16].float() * scale).to(torch.bfloat16)
            o = ops.vit_attention(q2)
            t = q2.float().cpu().reshape(3, 16, F, 257, 64)
            oref = (torch.softmax(t[0] @ t[1].transpose(-1, -2), -1) @ t[2]).permute(1, 2, 0, 3).reshape(F * 257, 1024)
            assert relmax(o, oref) <= 8e-3
            # the CUDA-core paths: last query row of every frame, and sensitivity to the last key
            assert relmax(o.reshape(F, 257, 1024)[:, 256], oref.reshape(F, 257, 1024)[:, 256]) <= 8e-3


@pytest.fixture(scope="module")
def tower23():
    towers = {}

    def get(profile):
        if profile not in towers:
            sd = synth.clip_state_dict(synth.VIT_L14, 0, profile, n_layers=23)
            tw = CLIPVisionTower("synthetic", types.SimpleNamespace(mm_vision_select_layer=-2), delay_load=True)
            tw.load_model(sd)
            towers[profile] = (tw.to(DEV), sd)
        return towers[profile]
    return get


@pytest.mark.parametrize("profile", ["hf", "strong"])
def test_vit_l14_vs_hf_reference_fixture(golden, tower23, profile):
    """CUDA tower (bf16 operands, fp32 residual) vs HF CLIPVisionModel fp32 run through the reference's
    CLIPVisionTower.forward (fixture)."""
    g = golden(f"vit_l14_{profile}")
    tw, sd = tower23(profile)
    px = synth.pixels((2, 3, 224, 224), seed=3)
    feats = tw(px.to(DEV))
    assert feats.shape == (2, 256, 1024) and feats.dtype == torch.float32
    sub = T(g["sub"])
    err = float((feats[:, ::8, ::4].cpu() - sub).abs().max() / float(g["absmax"]))
    assert err <= TOL_BF16, err
    assert relmax(feats.norm(dim=-1), T(g["norms"])) <= TOL_BF16
    # list input -> list output, per image (clip_encoder.py:41-46), output dtype follows the image dtype
    lst = tw([px[0].to(DEV).half(), px[1].to(DEV).half()])
    assert isinstance(lst, list) and lst[0].shape == (1, 256, 1024) and lst[0].dtype == torch.float16


@pytest.mark.parametrize("ln_fold", [1, 0])
def test_vit_l14_tracks_bf16_operand_oracle(tower23, ln_fold):
    """Tight check against the oracle run with bf16-rounded GEMM operands (same numeric regime): catches
    structural bugs that the 1e-2 fp32 tolerance could hide.  Both LayerNorm schedules of the tower, each against the
    oracle in ITS arithmetic: folded into the QKV / fc1 GEMMs (default; the GEMM reads bf16(x) and bf16(gamma*W)) and the
    stand-alone LayerNorm kernels (the GEMM reads bf16(LN(x)) and bf16(W))."""
    tw, sd = tower23("strong")
    px = synth.pixels((1, 3, 224, 224), seed=5)
    prev = ops.vit_set_ln_fold(ln_fold)
    try:
        hid = tw.forward_hidden(px.to(DEV))
    finally:
        ops.vit_set_ln_fold(prev)
    emu = restate.vit_hidden(px, sd, 23, emulate="bf16_fold" if ln_fold else "bf16")
    assert relmax(hid, emu) <= 4e-3
    assert relmax(hid, restate.vit_hidden(px, sd, 23)) <= TOL_BF16


def test_tower_load_model_from_files(tower23, tmp_path):
    """load_model accepts what the reference's from_pretrained accepts (a save_pretrained directory) and plain
    state-dict files; the packed weight blob is identical to the one packed from the in-memory state dict."""
    import transformers
    tw, sd = tower23("hf")
    f = tmp_path / "clip_sd.pt"
    torch.save(sd, f)
    t2 = CLIPVisionTower(str(f), types.SimpleNamespace(mm_vision_select_layer=-2), delay_load=True)
    t2.load_model()
    assert t2.is_loaded and torch.equal(t2.weight_blob.cpu(), tw.weight_blob.cpu())
    cfg = transformers.CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24,
                                        num_attention_heads=16, image_size=224, patch_size=14, projection_dim=768)
    hf = transformers.CLIPVisionModel(cfg)
    hf.load_state_dict({k: v for k, v in sd.items()}, strict=False)
    d = tmp_path / "hf_dir"
    hf.save_pretrained(d)
    t3 = CLIPVisionTower(str(d), types.SimpleNamespace(mm_vision_select_layer=-2), delay_load=True)
    t3.load_model()
    assert torch.equal(t3.weight_blob.cpu(), tw.weight_blob.cpu())


@pytest.mark.parametrize("select_layer,select_feature", [(-1, "patch"), (-2, "cls_patch"), (12, "patch"), (0, "cls_patch")])
def test_tower_select_layer_and_feature_variants(select_layer, select_feature):
    """Every (select_layer, select_feature) the reference's feature_select accepts (clip_encoder.py:29-37): the full
    24-layer tower (-1), CLS kept, a positive layer index, and index 0 = the pre-LayerNorm embeddings."""
    sd = synth.clip_state_dict(synth.VIT_L14, 0, "hf", n_layers=24)
    args = types.SimpleNamespace(mm_vision_select_layer=select_layer, mm_vision_select_feature=select_feature)
    tw = CLIPVisionTower("synthetic", args, delay_load=True)
    tw.load_model(sd)
    tw = tw.to(DEV)
    px = synth.pixels((1, 3, 224, 224), seed=6)
    feats = tw(px.to(DEV))
    ref = restate.tower_forward(px, sd, select_layer, synth.VIT_L14, select_feature)
    assert feats.shape == ref.shape == (1, 257 if select_feature == "cls_patch" else 256, 1024)
    assert relmax(feats, ref) <= TOL_BF16
    with pytest.raises(ValueError):
        bad = CLIPVisionTower("synthetic", types.SimpleNamespace(mm_vision_select_layer=-2,
                                                                 mm_vision_select_feature="cls"), delay_load=True)
        bad.load_model(sd)
        bad.to(DEV)(px.to(DEV))


def test_tower_is_bit_reproducible(tower23):
    """Repeated forwards of the same frames give the same bits (no atomics-ordered sums, no races between the warp
    roles): 37 frames = ragged last tiles / 592 attention items on 296 CTAs; 3 frames = the small-batch path with
    programmatic dependent launch."""
    tw, _ = tower23("hf")
    g = torch.Generator(device=DEV)
    g.manual_seed(5)
    for n, reps in ((37, 12), (3, 25)):
        px = torch.randn(n, 3, 224, 224, device=DEV, generator=g).to(torch.bfloat16)
        ref = tw.forward_hidden(px).clone()
        assert torch.isfinite(ref).all()
        for _ in range(reps):
            assert torch.equal(tw.forward_hidden(px), ref)


def test_tower_rejects_wrong_image_size(tower23):
    tw, _ = tower23("hf")
    with pytest.raises(ValueError):
        tw(torch.zeros(1, 3, 112, 112, device=DEV))


# ------------------------------------------------------------------ whole path through the drop-in interface
def test_handsonvlm_path_end_to_end(tower23):
    """SURVEY 8d config 2 at reduced frame count: pixels -> ViT -> pool -> projector(4096) -> splice -> gather,
    through the reference's method signatures, against the fp32 oracle."""
    tw, sd = tower23("hf")
    D, t, B = 4096, 6, 2
    ps = synth.projector_state(D)
    proj = torch.nn.Linear(1024, D)
    proj.weight.data.copy_(ps["mm_projector.weight"])
    proj.bias.data.copy_(ps["mm_projector.bias"])
    emb = torch.nn.Embedding(synth.VOCAB, D)
    emb.weight.data.copy_(synth.embed_table(D))
    cfg = types.SimpleNamespace(fuse_input_mode="origin", video_compress_mode="temporal_spatial_pool",
                                mm_hidden_size=1024, input_type="video")
    host = _host(arch.HandsOnVLMMetaForCausalLM, tw, proj.to(torch.bfloat16), emb.to(torch.bfloat16), cfg, B)
    px = synth.pixels((B, t, 3, 224, 224), seed=9)
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=B, seed=9)
    with torch.no_grad():
        r = host.prepare_inputs_labels_for_multimodal(ids.to(DEV), mask.to(DEV), None, labels.to(DEV),
                                                      px.to(DEV).to(torch.bfloat16), future_hands=fh.to(DEV),
                                                      future_valid=fv.to(DEV), is_evaluate=False)
    assert r[0] is None and r[2] is None
    m2, e2, l2 = r[1], r[3], r[4]
    vis, _ = restate.pipeline(px.to(torch.bfloat16).float(), sd, ps["mm_projector.weight"], ps["mm_projector.bias"])
    table = synth.embed_table(D).to(torch.bfloat16).float()
    rm, re_, rl = restate.splice(ids, mask, labels, vis, table, "handsonvlm", future_hands=fh)
    assert e2.shape == (B, ids.shape[1] + 6 + 256 - 1, D) and e2.dtype == torch.bfloat16
    assert torch.equal(l2.cpu(), rl) and torch.equal(m2.cpu(), rm)
    assert relmax(e2, re_) <= TOL_BF16
    hidden = synth.gen("e2e.hidden", tuple(e2.shape), 1.0, 9).to(torch.bfloat16)
    gout, valid = host.gather_hand_traj_states(hidden.to(DEV), l2, future_valid=fv.to(DEV))
    ro, rv, _ = restate.gather_hand_traj(hidden, rl)
    assert torch.equal(gout.cpu(), ro) and torch.equal(valid.cpu(), rv)
    assert int(host.last_visual_token_index) == 35 + 262


def test_config3_full_size_16_clips_13b_shapes(tower23):
    """BASELINE configs[2] at full size (16 clips x 100 frames, projector 1024->5120) through the drop-in interface,
    checked by size-independent properties: (i) batch invariance -- a clip's visual tokens do not depend on what else is
    in the batch, bit for bit; (ii) labels / mask are bit-exact against the oracle's splice of the same prompts;
    (iii) in the spliced output the visual rows are exact copies of the tokens and the text rows exact copies of the
    embedding-table rows (SURVEY 8a note i)."""
    tw, sd = tower23("hf")
    D, t, B = 5120, 100, 16
    ps = synth.projector_state(D)
    proj = torch.nn.Linear(1024, D)
    proj.weight.data.copy_(ps["mm_projector.weight"])
    proj.bias.data.copy_(ps["mm_projector.bias"])
    emb = torch.nn.Embedding(synth.VOCAB, D)
    emb.weight.data.copy_(synth.embed_table(D))
    cfg = types.SimpleNamespace(fuse_input_mode="origin", video_compress_mode="temporal_spatial_pool",
                                mm_hidden_size=1024, input_type="video")
    host = _host(arch.HandsOnVLMMetaForCausalLM, tw, proj.to(torch.bfloat16), emb.to(torch.bfloat16), cfg, B)
    g = torch.Generator(device=DEV)
    g.manual_seed(31)
    px = torch.randn(B, t, 3, 224, 224, device=DEV, generator=g).to(torch.bfloat16)
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=B, seed=12)
    with torch.no_grad():
        r = host.prepare_inputs_labels_for_multimodal(ids.to(DEV), mask.to(DEV), None, labels.to(DEV), px,
                                                      future_hands=fh.to(DEV), future_valid=fv.to(DEV), is_evaluate=False)
        helper = arch.VisualToTokenHelper(images_raw_encode=tw, images_mm_projector=host.get_model().mm_projector,
                                          fuse_input_mode="origin", video_compress_mode="temporal_spatial_pool",
                                          mm_hidden_size=1024, token_dim=D)
        tok_all, vmask = helper.pipeline(images=px)
        tok_5, _ = helper.pipeline(images=px[5:6])
        tok_15, _ = helper.pipeline(images=px[15:16])
    m2, e2, l2 = r[1], r[3], r[4]
    T_in = ids.shape[1]
    assert e2.shape == (B, T_in + 355, D) and tok_all.shape == (B, 356, D) and bool(vmask.all())
    assert torch.equal(tok_all[5:6], tok_5) and torch.equal(tok_all[15:16], tok_15)
    assert torch.isfinite(tok_all.float()).all()
    # ints against the oracle (visual values do not influence them)
    rm, _, rl = restate.splice(ids, mask, labels, torch.zeros(B, 356, 8), torch.zeros(synth.VOCAB, 8), "handsonvlm",
                               future_hands=fh)
    assert torch.equal(l2.cpu(), rl) and torch.equal(m2.cpu(), rm)
    # exact-copy rows
    table = host.get_model().embed_tokens.weight
    for b in (0, 7, 15):
        start = int((ids[b] == -200).nonzero()[0])
        assert torch.equal(e2[b, start:start + 356], tok_all[b])
        assert torch.equal(e2[b, :start], table[ids[b, :start].to(DEV)])
        tail_ids = ids[b, start + 1:].to(DEV)
        plain = tail_ids != 32100                     # <hand_traj> rows carry the added GT-hand embedding
        assert torch.equal(e2[b, start + 356:][plain], table[tail_ids[plain]])


def test_llava_config1_single_image_fp32(tower23):
    """BASELINE config 1: single 224x224 image, fp32, 256 tokens -> projector 4096 -> LLaVA splice."""
    tw, sd = tower23("hf")
    D = 4096
    ps = synth.projector_state(D)
    proj = torch.nn.Linear(1024, D)
    proj.weight.data.copy_(ps["mm_projector.weight"])
    proj.bias.data.copy_(ps["mm_projector.bias"])
    emb = torch.nn.Embedding(synth.VOCAB, D)
    emb.weight.data.copy_(synth.embed_table(D))
    host = _host(arch.LitaMetaForCausalLM, tw, proj, emb, types.SimpleNamespace(input_type="image"), 1)
    px = synth.pixels((1, 3, 224, 224), seed=12)
    ids, mask, labels = synth.prompt_llava(seed=31)
    with torch.no_grad():
        r = host.prepare_inputs_labels_for_multimodal(ids.to(DEV), mask.to(DEV), None, labels.to(DEV), px.to(DEV))
    m2, e2, l2 = r[1], r[3], r[4]
    vis = restate.project(restate.tower_forward(px, sd, -2), ps["mm_projector.weight"], ps["mm_projector.bias"])
    rm, re_, rl = restate.splice(ids, mask, labels, vis, synth.embed_table(D), "llava")
    assert e2.shape == (1, 311, D) and e2.dtype == torch.float32
    assert torch.equal(l2.cpu(), rl) and torch.equal(m2.cpu(), rm)
    assert (l2[0, 35:291] == -100).all()
    assert relmax(e2[0, 35:291], re_[0, 35:291]) <= TOL_BF16
    assert torch.equal(e2[0, :35].cpu(), re_[0, :35])
    # T == 1 early-out (llava_arch.py:117-120)
    pkv = [[torch.zeros(1, 2, 9, 4, device=DEV), torch.zeros(1, 2, 9, 4, device=DEV)]]
    r = host.prepare_inputs_labels_for_multimodal(torch.tensor([[5]], device=DEV), torch.ones(1, 1, dtype=torch.bool, device=DEV),
                                                  pkv, None, px.to(DEV))
    assert r[3] is None and r[1].shape == (1, 10)


def test_llava_list_of_image_groups_end_to_end(tower23):
    """LLaVA list input (images_to_tokens list path): groups of 1 and 2 images through ONE tower call, ragged splice."""
    tw, sd = tower23("hf")
    D = 4096
    ps = synth.projector_state(D)
    proj = torch.nn.Linear(1024, D)
    proj.weight.data.copy_(ps["mm_projector.weight"])
    proj.bias.data.copy_(ps["mm_projector.bias"])
    emb = torch.nn.Embedding(synth.VOCAB, D)
    emb.weight.data.copy_(synth.embed_table(D))
    host = _host(arch.LitaMetaForCausalLM, tw, proj, emb, types.SimpleNamespace(input_type="image"), 2)
    groups = [synth.pixels((1, 3, 224, 224), seed=21), synth.pixels((2, 3, 224, 224), seed=22)]
    ids, mask, labels = synth.prompt_llava(seed=35)
    ids = torch.cat([ids, ids], 0)
    mask = torch.cat([mask, mask], 0)
    labels = torch.cat([labels, labels], 0)
    with torch.no_grad():
        r = host.prepare_inputs_labels_for_multimodal(ids.to(DEV), mask.to(DEV), None, labels.to(DEV),
                                                      [g.to(DEV) for g in groups])
    m2, e2, l2 = r[1], r[3], r[4]
    blocks = [restate.project(restate.tower_forward(g, sd, -2), ps["mm_projector.weight"],
                              ps["mm_projector.bias"]).reshape(-1, D) for g in groups]
    rm, re_, rl = restate.splice(ids, mask, labels, blocks, synth.embed_table(D), "llava")
    T_in = ids.shape[1]
    assert e2.shape == (2, T_in - 1 + 512, D) == tuple(re_.shape)
    assert torch.equal(l2.cpu(), rl) and torch.equal(m2.cpu(), rm)
    assert relmax(e2, re_) <= TOL_BF16


@pytest.mark.parametrize("mode", ["temporal_spatial_pool", "temporal", "spatial_pool"])
def test_pool_before_fc2_matches_full_tower(tower23, mode, monkeypatch):
    """The last layer's fc2 applied after pooling (linearity) vs fc2 on every token row, then pooling: same tokens to
    bf16 rounding, and both within the bar of the fp32 oracle."""
    tw, sd = tower23("strong")
    D, t, B = 4096, 5, 2
    ps = synth.projector_state(D)
    proj = torch.nn.Linear(1024, D)
    proj.weight.data.copy_(ps["mm_projector.weight"])
    proj.bias.data.copy_(ps["mm_projector.bias"])
    proj = proj.to(DEV).to(torch.bfloat16)
    px = synth.pixels((B, t, 3, 224, 224), seed=23)
    with torch.no_grad():
        monkeypatch.setattr(arch, "_POOL_BEFORE_FC2", True)
        fast = arch.video_tokens(tw, proj, px.to(DEV).to(torch.bfloat16), mode)
        monkeypatch.setattr(arch, "_POOL_BEFORE_FC2", False)
        full = arch.video_tokens(tw, proj, px.to(DEV).to(torch.bfloat16), mode)
    assert fast.shape == full.shape
    assert relmax(fast, full) <= 8e-3
    ref = restate.pipeline(px.to(torch.bfloat16).float(), sd, ps["mm_projector.weight"], ps["mm_projector.bias"], mode)[0]
    assert relmax(fast, ref) <= TOL_BF16 and relmax(full, ref) <= TOL_BF16


def test_lita_videos_to_tokens_archs(tower23):
    tw, sd = tower23("hf")
    D = 256
    ps = synth.projector_state(D)
    proj = torch.nn.Linear(1024, D)
    proj.weight.data.copy_(ps["mm_projector.weight"])
    proj.bias.data.copy_(ps["mm_projector.bias"])
    px = synth.pixels((1, 5, 3, 224, 224), seed=7)
    feats = restate.tower_forward(px[0], sd, -2)
    tok = restate.project(feats, ps["mm_projector.weight"], ps["mm_projector.bias"]).reshape(1, 5, 256, D)
    for arch_name in ("all", "temporal", "spatial", "temporal_spatial", "temporal_spatial_pool", "spatial_pool"):
        host = _host(arch.LitaMetaForCausalLM, tw, proj, torch.nn.Embedding(4, D),
                     types.SimpleNamespace(input_type="video", video_arch=arch_name), 1)
        with torch.no_grad():
            out = host.visual_to_tokens(px.to(DEV))
        ref = restate.pool_tokens(tok, arch_name)
        assert out.shape == ref.shape and relmax(out, ref) <= TOL_BF16
    with pytest.raises(ValueError):
        _host(arch.LitaMetaForCausalLM, tw, proj, torch.nn.Embedding(4, D),
              types.SimpleNamespace(input_type="video", video_arch="nope"), 1).visual_to_tokens(px.to(DEV))


def test_training_shaped_backward(tower23):
    """fwd + bwd through pooling / projector / splice / gather with upstream grads; projector and visual
    gradients against the fp32 oracle (SURVEY 8d config 5 at reduced size)."""
    tw, sd = tower23("hf")
    D, t, B = 512, 4, 2
    ps = synth.projector_state(D)
    proj = torch.nn.Linear(1024, D)
    proj.weight.data.copy_(ps["mm_projector.weight"])
    proj.bias.data.copy_(ps["mm_projector.bias"])
    emb = torch.nn.Embedding(synth.VOCAB, D)
    emb.weight.data.copy_(synth.embed_table(D))
    cfg = types.SimpleNamespace(fuse_input_mode="origin", video_compress_mode="temporal_spatial_pool",
                                mm_hidden_size=1024, input_type="video")
    host = _host(arch.HandsOnVLMMetaForCausalLM, tw, proj, emb, cfg, B)
    px = synth.pixels((B, t, 3, 224, 224), seed=19)
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=B, seed=19)
    r = host.prepare_inputs_labels_for_multimodal(ids.to(DEV), mask.to(DEV), None, labels.to(DEV), px.to(DEV),
                                                  future_hands=fh.to(DEV), future_valid=fv.to(DEV), is_evaluate=False)
    e2, l2 = r[3], r[4]
    gout, _ = host.gather_hand_traj_states(e2, l2)          # use the embeddings as stand-in hidden states
    de = synth.gen("tr.de", tuple(e2.shape), 1.0, 19)
    dg = synth.gen("tr.dg", tuple(gout.shape), 1.0, 19)
    ((e2 * de.to(DEV)).sum() + (gout * dg.to(DEV)).sum()).backward()
    # oracle: grads wrt spliced embeddings = de + scatter(dg); -> visual rows -> pooled-token grads -> dW, db
    _, _, rows = restate.gather_hand_traj(torch.zeros(B, e2.shape[1], D), l2.cpu())
    d_emb = de + restate.gather_hand_traj_backward(dg, rows, e2.shape[1])
    d_vis, d_tab = restate.splice_backward(d_emb, ids, t + 256, B, synth.VOCAB)
    feats = restate.tower_forward(px.reshape(B * t, 3, 224, 224), sd, -2).reshape(B, t, 256, 1024)
    pooled = restate.pool_tokens(feats, "temporal_spatial_pool")
    dW, db = restate.projector_grads(pooled, d_vis)
    pw, pb = host.model.mm_projector.weight, host.model.mm_projector.bias
    assert relmax(pw.grad, dW) <= TOL_BF16 and relmax(pb.grad, db) <= 1e-4
    assert relmax(host.model.embed_tokens.weight.grad, d_tab) <= 1e-5


def test_visual_token_cache_for_generation(tower23):
    """with config.hvlm_cache_visual_tokens the ViT/pool/projector run once per clip, not once per
    generated token; results are bit-identical and an in-place change of the clip invalidates the cache."""
    tw, sd = tower23("hf")
    D, t = 256, 4
    ps = synth.projector_state(D)
    proj = torch.nn.Linear(1024, D)
    proj.weight.data.copy_(ps["mm_projector.weight"])
    proj.bias.data.copy_(ps["mm_projector.bias"])
    emb = torch.nn.Embedding(synth.VOCAB, D)
    cfg = types.SimpleNamespace(fuse_input_mode="origin", video_compress_mode="temporal_spatial_pool", mm_hidden_size=1024,
                                input_type="video", hvlm_cache_visual_tokens=True)
    host = _host(arch.HandsOnVLMMetaForCausalLM, tw, proj, emb, cfg, 1)
    px = synth.pixels((1, t, 3, 224, 224), seed=5).to(DEV)
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=1, seed=5)
    args = (ids.to(DEV), mask.to(DEV), None, None, px)
    with torch.no_grad():
        n0 = ops.launch_count()
        r1 = host.prepare_inputs_labels_for_multimodal(*args, is_evaluate=True)
        n1 = ops.launch_count()
        # "next token": one more id appended, same clip tensor -> cache hit, no tower launches
        ids2 = torch.cat([ids, torch.tensor([[7]])], 1).to(DEV)
        m2 = torch.ones_like(ids2, dtype=torch.bool)
        r2 = host.prepare_inputs_labels_for_multimodal(ids2, m2, None, None, px, is_evaluate=True)
        n2 = ops.launch_count()
        assert n1 - n0 > 30 and n2 - n1 <= 4
        assert torch.equal(r2[3][:, : r1[3].shape[1]], r1[3])
        px.add_(0.0)                                   # in-place touch bumps the version counter -> recompute
        host.prepare_inputs_labels_for_multimodal(*args, is_evaluate=True)
        assert ops.launch_count() - n2 > 30
    # never cached when autograd is recording
    host.clear_visual_token_cache()
    r3 = host.prepare_inputs_labels_for_multimodal(ids.to(DEV), mask.to(DEV), None, labels.to(DEV), px,
                                                   future_hands=fh.to(DEV), future_valid=fv.to(DEV))
    assert host.__dict__["_hvlm_visual_cache"].key is None and r3[3].requires_grad


def test_uint8_frames_with_fused_clip_normalisation(tower23):
    """raw uint8 NHWC frames through the tower == the float path on (u8/255 - mean)/std pixels."""
    tw, sd = tower23("hf")
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (3, 224, 224, 3), generator=g, dtype=torch.uint8)
    mean = torch.tensor(ops.CLIP_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(ops.CLIP_STD).view(1, 3, 1, 1)
    px = (u8.permute(0, 3, 1, 2).float() / 255.0 - mean) / std
    h_u8 = tw.forward_hidden(u8.to(DEV))
    h_f = tw.forward_hidden(px.to(DEV))
    assert relmax(h_u8, h_f) <= 3e-3                       # same bf16 operands up to rounding of the normalised pixels
    ref = restate.vit_hidden(px, sd, 23)
    assert relmax(h_u8, ref) <= TOL_BF16
    feats = tw(u8.to(DEV))
    assert feats.dtype == torch.bfloat16 and feats.shape == (3, 256, 1024)
    with pytest.raises(ValueError):
        tw.forward_hidden(torch.zeros(1, 3, 224, 224, dtype=torch.uint8, device=DEV))


# frame de-duplication: tests/test_gpu_round2.py (kernel vs oracle frame map, tokens vs oracle numbers, static capacity)
