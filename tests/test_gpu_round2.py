"""GPU parity tests added in round 2: the sizes the bench times (BASELINE configs[1] and [2] against oracle NUMBERS, the
M = 25 700 GEMM schedule), the NCCL path of the training-shaped variant on two ranks, frame de-duplication and the
uint8 resize / crop against the oracle, and the sync-free modes' deferred contract checks.  Every call goes through the
C ABI; tolerances as in test_gpu_parity.py (BASELINE.json north_star)."""
import os
import socket
import types

import numpy as np
import pytest
import torch

import hvlm_b200
from hvlm_b200 import _lib as L
from hvlm_b200 import arch, ops
from hvlm_b200 import dist as hd
from hvlm_b200.tower import CLIPVisionTower
from oracle import restate, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_BF16 = 1e-2


def relmax(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


class _Inner(torch.nn.Module):
    def __init__(self, tower, proj, emb):
        super().__init__()
        self.vision_tower, self.mm_projector, self.embed_tokens = tower, proj, emb

    def get_vision_tower(self):
        return self.vision_tower


def make_host(tower, proj, emb, config, B, dev=DEV):
    class Host(torch.nn.Module, arch.HandsOnVLMMetaForCausalLM):
        def __init__(self):
            super().__init__()
            self.model = _Inner(tower, proj, emb)
            self.config = config
            self.token_dim, self.B = proj.out_features, B

        def get_model(self):
            return self.model
    return Host().to(dev)


def video_cfg(**kw):
    return types.SimpleNamespace(fuse_input_mode="origin", video_compress_mode="temporal_spatial_pool", mm_hidden_size=1024,
                                 input_type="video", **kw)


def projector(D, dtype=torch.bfloat16):
    ps = synth.projector_state(D)
    proj = torch.nn.Linear(1024, D)
    proj.weight.data.copy_(ps["mm_projector.weight"])
    proj.bias.data.copy_(ps["mm_projector.bias"])
    return proj.to(dtype), ps


@pytest.fixture(scope="module")
def tower_hf():
    sd = synth.clip_state_dict(synth.VIT_L14, 0, "hf", n_layers=23)
    tw = CLIPVisionTower("synthetic", types.SimpleNamespace(mm_vision_select_layer=-2), delay_load=True)
    tw.load_model(sd)
    return tw.to(DEV), sd


# ------------------------------------------------------------------ BASELINE configs[1] at FULL size, oracle numbers
@pytest.fixture(scope="module")
def config2_oracle(tower_hf):
    """One 100-frame clip through the fp32 CPU oracle in the reference's order (project all 25 600 tokens, then pool)."""
    _, sd = tower_hf
    D = 4096
    ps = synth.projector_state(D)
    px = synth.pixels((1, 100, 3, 224, 224), seed=41).to(torch.bfloat16)
    torch.set_num_threads(os.cpu_count() or 1)
    vis, _ = restate.pipeline(px.float(), sd, ps["mm_projector.weight"], ps["mm_projector.bias"])
    return px, vis


@pytest.mark.parametrize("pool_before_fc2", [True, False])
def test_config2_full_size_vs_oracle(tower_hf, config2_oracle, pool_before_fc2, monkeypatch):
    """HandsOnVLM-7B clip: 100 frames -> ViT-L/14 -> slow-fast pool -> projector 4096 -> splice (T=62 -> 417) + gather,
    through prepare_inputs_labels_for_multimodal, against the oracle's numbers.  M = 25 700 token rows: the tile schedule
    the bench times (fc1 1 616 tiles; fc2 / out_proj 404 tiles on 74 pairs with the half-tile tail)."""
    tw, sd = tower_hf
    px, vis = config2_oracle
    D, B = 4096, 1
    monkeypatch.setattr(arch, "_POOL_BEFORE_FC2", pool_before_fc2)
    proj, ps = projector(D)
    emb = torch.nn.Embedding(synth.VOCAB, D)
    emb.weight.data.copy_(synth.embed_table(D))
    host = make_host(tw, proj, emb.to(torch.bfloat16), video_cfg(), B)
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=B, seed=41)
    assert ids.shape[1] == 62
    with torch.no_grad():
        r = host.prepare_inputs_labels_for_multimodal(ids.to(DEV), mask.to(DEV), None, labels.to(DEV), px.to(DEV),
                                                      future_hands=fh.to(DEV), future_valid=fv.to(DEV), is_evaluate=False)
    m2, e2, l2 = r[1], r[3], r[4]
    table = synth.embed_table(D).to(torch.bfloat16).float()
    rm, re_, rl = restate.splice(ids, mask, labels, vis, table, "handsonvlm", future_hands=fh)
    assert e2.shape == (1, 417, D) and e2.dtype == torch.bfloat16
    assert torch.equal(l2.cpu(), rl) and torch.equal(m2.cpu(), rm)
    assert l2[0, 411:415].tolist() == [32100] * 4
    err = relmax(e2[:, 35:35 + 356], re_[:, 35:35 + 356])                 # the 356 visual rows: bf16-GEMM backed
    assert err <= TOL_BF16, err
    assert relmax(e2, re_) <= TOL_BF16
    hidden = synth.gen("cfg2.hidden", tuple(e2.shape), 1.0, 41).to(torch.bfloat16)
    gout, valid = host.gather_hand_traj_states(hidden.to(DEV), l2, future_valid=fv.to(DEV))
    ro, rv, rows = restate.gather_hand_traj(hidden, rl)
    assert torch.equal(gout.cpu(), ro) and torch.equal(valid.cpu(), rv) and rows[0].tolist() == [410, 411, 412, 413]
    assert int(host.last_visual_token_index) == restate.last_visual_token_index(ids, 356) == 35 + 356


def test_config3_sampled_clips_vs_oracle(tower_hf):
    """BASELINE configs[2]: 13B shapes (projector 1024 -> 5120), 16 clips x 100 frames in ONE batch (M = 411 200 token rows),
    collator-padded ragged prompts; oracle numbers for two sampled clips of the 16, ints for all."""
    tw, sd = tower_hf
    D, t, B = 5120, 100, 16
    proj, ps = projector(D)
    emb = torch.nn.Embedding(synth.VOCAB, D)
    emb.weight.data.copy_(synth.embed_table(D))
    host = make_host(tw, proj, emb.to(torch.bfloat16), video_cfg(), B)
    g = torch.Generator(device=DEV)
    g.manual_seed(43)
    px = torch.randn(B, t, 3, 224, 224, device=DEV, generator=g).to(torch.bfloat16)
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=B, seed=43, ragged=True)
    with torch.no_grad():
        r = host.prepare_inputs_labels_for_multimodal(ids.to(DEV), mask.to(DEV), None, labels.to(DEV), px,
                                                      future_hands=fh.to(DEV), future_valid=fv.to(DEV), is_evaluate=False)
    m2, e2, l2 = r[1], r[3], r[4]
    assert e2.shape == (B, ids.shape[1] + 355, D)
    rm, _, rl = restate.splice(ids, mask, labels, torch.zeros(B, 356, 8), torch.zeros(synth.VOCAB, 8), "handsonvlm",
                               future_hands=fh)
    assert torch.equal(l2.cpu(), rl) and torch.equal(m2.cpu(), rm)
    torch.set_num_threads(os.cpu_count() or 1)
    table = synth.embed_table(D).to(torch.bfloat16).float()
    for b in (3, 12):
        vis, _ = restate.pipeline(px[b:b + 1].float().cpu(), sd, ps["mm_projector.weight"], ps["mm_projector.bias"])
        _, re_, _ = restate.splice(ids[b:b + 1], mask[b:b + 1], labels[b:b + 1], vis, table, "handsonvlm",
                                   future_hands=fh[b:b + 1])
        assert relmax(e2[b], re_[0]) <= TOL_BF16, b


@pytest.mark.parametrize("M,N,K", [(25700, 1024, 1024), (25700, 1024, 4096), (25700, 4096, 1024), (25700, 3072, 1024)])
def test_gemm_full_size_schedules(M, N, K):
    """The tower's GEMM shapes at the 100-frame M: out_proj (K = 1024: residual-LOAD epilogue, 404 tiles + half-tile tail),
    fc2 (K = 4096: TMA reduce-add epilogue), fc1 / QKV-sized plain stores.  Same bf16 operands, fp32 math as reference."""
    a = synth.gen("fs.A", (M, K), 1.0, 3).to(torch.bfloat16).to(DEV)
    w = synth.gen("fs.W", (N, K), K ** -0.5, 3).to(torch.bfloat16).to(DEV)
    bias = synth.gen("fs.b", (N,), 0.5, 3).to(DEV)
    ref = a.float() @ w.float().t() + bias
    if N == 1024:
        res = synth.gen("fs.r", (M, N), 1.0, 4).to(DEV)
        r2 = res.clone()
        ops.gemm(a, w, bias, epilogue="residual", resid=r2, out=r2)       # in place, as the tower uses it
        assert relmax(r2, ref + res) <= 2e-5
        r3 = res.clone()
        ops.gemm(a, w, bias, epilogue="residual", resid=r3, out=r3)
        assert torch.equal(r2, r3)                                        # bit-reproducible
    else:
        out = ops.gemm(a, w, bias, epilogue="quick_gelu" if N == 4096 else "bias", out_dtype=torch.bfloat16)
        want = restate.quick_gelu(ref) if N == 4096 else ref
        assert relmax(out, want) <= 6e-3


# ------------------------------------------------------------------ training-shaped variant on two ranks over NCCL
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


TRAIN_D, TRAIN_T, TRAIN_B = 4096, 4, 4       # config-5 gradient shapes (16.8 MB fp32 bucket), 4 clips per rank


def _train_inputs(rank):
    px = synth.pixels((TRAIN_B, TRAIN_T, 3, 224, 224), seed=100 + rank)
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=TRAIN_B, seed=100 + rank)
    Lout = ids.shape[1] + TRAIN_T + 256 - 1
    de = synth.gen("nccl.de", (TRAIN_B, Lout, TRAIN_D), 1.0, 100 + rank)
    dg = synth.gen("nccl.dg", (TRAIN_B, 2, 4, TRAIN_D // 2), 1.0, 100 + rank)
    return px, ids, mask, labels, fh, fv, de, dg


def _nccl_worker(rank, world, port, expect_path, q):
    try:
        os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                          MASTER_PORT=str(port))
        dev = torch.device(f"cuda:{rank}")
        torch.cuda.set_device(dev)
        hd.init_process_group("nccl")
        sd = synth.clip_state_dict(synth.VIT_L14, 0, "hf", n_layers=23)
        tw = CLIPVisionTower("synthetic", types.SimpleNamespace(mm_vision_select_layer=-2), delay_load=True)
        tw.load_model(sd)
        proj, _ = projector(TRAIN_D)
        emb = torch.nn.Embedding(synth.VOCAB, TRAIN_D)
        emb.weight.data.copy_(synth.embed_table(TRAIN_D))
        emb.weight.requires_grad_(False)                  # embed_tokens grads belong to the LLM, not to this path
        host = make_host(tw.to(dev), proj, emb.to(torch.bfloat16), video_cfg(hvlm_static_splice=True), TRAIN_B, dev)
        pm = host.model.mm_projector
        red = hd.ProjectorGradReducer(pm)
        px, ids, mask, labels, fh, fv, de, dg = [x.to(dev) for x in _train_inputs(rank)]
        r = host.prepare_inputs_labels_for_multimodal(ids, mask, None, labels, px.to(torch.bfloat16), future_hands=fh,
                                                      future_valid=fv, is_evaluate=False)
        gout, _ = host.gather_hand_traj_states(r[3], r[4], strict=False)
        torch.autograd.backward([r[3], gout], [de.to(torch.bfloat16), dg.to(torch.bfloat16)])
        local_w = pm.weight.grad.float().clone()
        red.reduce_async(timed=True)
        assert red.pending is not None
        red.wait()
        ms = red.last_allreduce_ms()
        arch.check_deferred_status(host)
        exp = torch.load(expect_path)
        nW = pm.weight.numel()
        relW = relmax(red.bucket[:nW].view(pm.weight.shape), exp["dW"])       # the all-reduced fp32 bucket
        relb = relmax(red.bucket[nW:], exp["db"])
        # .grad carries the same values in the parameters' dtype (bf16 here)
        assert torch.equal(pm.weight.grad, red.bucket[:nW].view(pm.weight.shape).to(pm.weight.dtype))
        assert torch.equal(pm.bias.grad, red.bucket[nW:].to(pm.bias.dtype))
        rel_local = relmax(local_w, exp["dW_rank"][rank])
        # every rank must end with the same bits
        chk = pm.weight.grad.float().sum().reshape(1).double()
        both = [torch.zeros_like(chk) for _ in range(world)]
        torch.distributed.all_gather(both, chk)
        same = bool(both[0] == both[1])
        q.put((rank, relW, relb, rel_local, same, ms, None))
        torch.distributed.destroy_process_group()
    except Exception as e:      # surface the failure in the parent instead of a hang
        import traceback
        q.put((rank, None, None, None, False, None, traceback.format_exc()))


def test_nccl_two_rank_projector_grad_allreduce(tmp_path):
    """SURVEY 8d config 5 / north_star: fwd + bwd through pool / projector / splice / gather on each rank, then ONE NCCL
    all-reduce (mean) of the projector gradients through dist.ProjectorGradReducer (flat fp32 bucket written by the wgrad
    kernels, side stream, AVG).  All-reduced dW / db against the fp32 oracle's autograd formulas summed over both ranks'
    clips: <= 1e-2 / 1e-4 of max|ref|."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    sd = synth.clip_state_dict(synth.VIT_L14, 0, "hf", n_layers=23)
    ps = synth.projector_state(TRAIN_D)
    dWs, dbs = [], []
    torch.set_num_threads(os.cpu_count() or 1)
    for rank in range(2):
        px, ids, mask, labels, fh, fv, de, dg = _train_inputs(rank)
        px, de, dg = px.to(torch.bfloat16).float(), de.to(torch.bfloat16).float(), dg.to(torch.bfloat16).float()
        Lout = de.shape[1]
        _, _, rl = restate.splice(ids, mask, labels, torch.zeros(TRAIN_B, TRAIN_T + 256, 8), torch.zeros(synth.VOCAB, 8),
                                  "handsonvlm", future_hands=fh)
        _, _, rows = restate.gather_hand_traj(torch.zeros(TRAIN_B, Lout, TRAIN_D), rl)
        d_emb = de + restate.gather_hand_traj_backward(dg, rows, Lout)
        d_vis, _ = restate.splice_backward(d_emb, ids, TRAIN_T + 256, TRAIN_B, synth.VOCAB)
        feats = restate.tower_forward(px.reshape(-1, 3, 224, 224), sd, -2).reshape(TRAIN_B, TRAIN_T, 256, 1024)
        dW, db = restate.projector_grads(restate.pool_tokens(feats, "temporal_spatial_pool"), d_vis)
        dWs.append(dW)
        dbs.append(db)
    expect = tmp_path / "expect.pt"
    torch.save({"dW": (dWs[0] + dWs[1]) / 2, "db": (dbs[0] + dbs[1]) / 2, "dW_rank": dWs}, expect)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, str(expect), q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = sorted(q.get(timeout=600) for _ in range(2))
        for p in procs:
            p.join(120)
    finally:
        for p in procs:
            if p.is_alive():
                p.terminate()
    for rank, relW, relb, rel_local, same, ms, err in res:
        assert err is None, err
        assert rel_local <= TOL_BF16, (rank, rel_local)
        assert relW <= TOL_BF16 and relb <= 1e-4, (rank, relW, relb)
        assert same
        assert ms is not None and ms > 0


def test_reducer_sink_single_gpu(tower_hf):
    """World size 1: no collective, but the wgrad / bias-grad kernels write into the reducer's flat bucket and `.grad`
    carries the same values."""
    tw, sd = tower_hf
    D, t, B = 512, 2, 2
    proj, ps = projector(D)
    emb = torch.nn.Embedding(synth.VOCAB, D)
    host = make_host(tw, proj, emb.to(torch.bfloat16), video_cfg(), B)
    pm = host.model.mm_projector
    red = hd.ProjectorGradReducer(pm)
    try:
        px = synth.pixels((B, t, 3, 224, 224), seed=7).to(DEV).to(torch.bfloat16)
        ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=B, seed=7)
        r = host.prepare_inputs_labels_for_multimodal(ids.to(DEV), mask.to(DEV), None, labels.to(DEV), px,
                                                      future_hands=fh.to(DEV), future_valid=fv.to(DEV), is_evaluate=False)
        r[3].float().square().sum().backward()
        g_local = pm.weight.grad.clone()
        bucket_w = red.bucket[: pm.weight.numel()].view(pm.weight.shape)
        assert relmax(bucket_w, g_local) <= 5e-3 and float(bucket_w.abs().max()) > 0      # bf16 rounding of .grad only
        red.reduce_async()
        red.wait()
        assert torch.equal(pm.weight.grad, bucket_w.to(torch.bfloat16))
        assert torch.equal(pm.bias.grad, red.bucket[pm.weight.numel():].to(torch.bfloat16))
    finally:
        red.close()
    assert ((D, 1024), torch.device(DEV)) not in ops._wgrad_sinks


# ------------------------------------------------------------------ frame de-duplication against the oracle
def test_frame_dedup_kernel_vs_oracle():
    base = synth.pixels((6, 3, 224, 224), seed=21).to(torch.bfloat16)
    order = [0, 1, 0, 2, 2, 3, 1, 0, 4, 5, 5, 3]
    clip = base[order].clone()
    clip[7, 2, 200, 17] += 0.5                      # one element differs: frame 7 is NOT a duplicate of frame 0
    fmap_ref, rep_ref = restate.frame_dedup(clip)
    fmap, rep, n_unique = ops.frame_dedup(clip.to(DEV))
    U = int(n_unique.item())
    assert U == rep_ref.numel() == 7
    assert torch.equal(fmap.cpu(), fmap_ref) and torch.equal(rep[:U].cpu(), rep_ref) and rep[U:].abs().sum().item() == 0
    # other element types / frame sizes (uint8 NHWC frames, fp32), all-distinct and all-equal inputs
    u8 = torch.randint(0, 256, (5, 224, 224, 3), dtype=torch.uint8)
    u8 = u8[[0, 1, 1, 0, 4]]
    f2, r2, n2 = ops.frame_dedup(u8.to(DEV))
    fr, rr = restate.frame_dedup(u8)
    assert torch.equal(f2.cpu(), fr) and int(n2.item()) == 3 and torch.equal(r2[:3].cpu(), rr)
    f3, _, n3 = ops.frame_dedup(base.float().to(DEV))
    assert int(n3.item()) == 6 and f3.tolist() == list(range(6))
    same = base[:1].expand(9, -1, -1, -1).contiguous()
    f4, r4, n4 = ops.frame_dedup(same.to(DEV))
    assert int(n4.item()) == 1 and f4.abs().sum().item() == 0 and r4.abs().sum().item() == 0
    # static capacity: the map is clamped, the count still tells
    f5, _, n5 = ops.frame_dedup(clip.to(DEV), capacity=4)
    assert int(n5.item()) == 7 and int(f5.max().item()) == 3
    # 1 600 frames (16 clips x 100): scan across several chunks
    big = base[torch.arange(1600) % 6].contiguous()
    f6, r6, n6 = ops.frame_dedup(big.to(DEV))
    assert int(n6.item()) == 6 and f6.tolist() == [i % 6 for i in range(1600)] and r6[:6].tolist() == list(range(6))
    out = ops.gather_rows(big.to(DEV), r6, 6)
    assert torch.equal(out.cpu(), base)


def test_frame_dedup_tokens_vs_oracle(tower_hf):
    """An EPIC-style clip (distinct frames tiled, handsonvlm/dataset/epic_dataset.py:90-95) with de-duplication: visual
    tokens against the ORACLE's numbers for the full tiled clip, bit-identical to the undeduplicated CUDA path, with ~5x
    fewer tower launches' worth of frames; the static-capacity mode does it without any host sync."""
    tw, sd = tower_hf
    D = 256
    proj, ps = projector(D, torch.float32)
    proj = proj.to(DEV)
    base = synth.pixels((4, 3, 224, 224), seed=21).to(torch.bfloat16)
    clip = base.repeat(5, 1, 1, 1).unsqueeze(0)                            # [1,20,...]: 4 distinct frames tiled x5
    ref, _ = restate.pipeline(clip.float(), sd, ps["mm_projector.weight"], ps["mm_projector.bias"])
    clip = clip.to(DEV)
    d = arch.distinct_frames(clip[0])
    assert d is not None and d[0].shape[0] == 4 and torch.equal(d[0][d[1].long()], clip[0])
    assert arch.distinct_frames(synth.pixels((5, 3, 224, 224), seed=22).to(DEV)) is None
    with torch.no_grad():
        for mode in ("temporal_spatial_pool", "spatial_pool", "temporal", "spatial", "none"):
            plain = arch.video_tokens(tw, proj, clip, mode, dedup=False)
            out = arch.video_tokens(tw, proj, clip, mode, dedup=True)
            assert torch.equal(out, plain), mode
            if mode == "temporal_spatial_pool":
                assert relmax(out, ref) <= TOL_BF16
        clip2 = torch.cat([clip, clip.flip(1)], 0)                         # duplicates across the batch dimension
        assert torch.equal(arch.video_tokens(tw, proj, clip2, "temporal_spatial_pool", True),
                           arch.video_tokens(tw, proj, clip2, "temporal_spatial_pool", False))
        # static capacity through the drop-in host: config.hvlm_dedup_frames = 4 distinct frames per clip, no host sync
        emb = torch.nn.Embedding(synth.VOCAB, D)
        host = make_host(tw, proj, emb, video_cfg(hvlm_dedup_frames=4, hvlm_static_splice=True), 1)
        ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=1, seed=5)
        args = (ids.to(DEV), mask.to(DEV), None, labels.to(DEV), clip)
        kw = dict(future_hands=fh.to(DEV), future_valid=fv.to(DEV), is_evaluate=False)
        n0 = ops.launch_count()
        r = host.prepare_inputs_labels_for_multimodal(*args, **kw)
        n_dedup = ops.launch_count() - n0
        arch.check_deferred_status(host)                                   # contract held: nothing raises
        assert torch.equal(r[3][0, 35:35 + 20 + 256].float(), plain_tsp(tw, proj, clip)[0].float())
        # contract broken: 4 distinct frames, capacity 2 per clip -> reported by the deferred check, not silently wrong
        # (by the first poll that finds the copied-out count: the same call's splice, or check_deferred_status)
        host.config.hvlm_dedup_frames = 2
        with pytest.raises(RuntimeError, match="distinct frames"):
            host.prepare_inputs_labels_for_multimodal(*args, **kw)
            arch.check_deferred_status(host)
        host.__dict__.pop("_hvlm_deferred", None)
        assert n_dedup < 7 * 23 + 40


def plain_tsp(tw, proj, clip):
    with torch.no_grad():
        return arch.video_tokens(tw, proj, clip, "temporal_spatial_pool", dedup=False)


# ------------------------------------------------------------------ uint8 resize + centre crop (CLIPImageProcessor)
@pytest.mark.parametrize("H,W", [(256, 456), (224, 224), (480, 640), (300, 200), (225, 230), (1080, 1920)])
def test_resize_center_crop_u8_bit_exact(H, W):
    """hvlm_resize_crop_u8 against the oracle (pinned bit-exactly to PIL.Image.resize(BICUBIC) + transformers' centre crop
    by tests/test_oracle_golden.py): bit-exact, for down- and up-scaling, portrait / landscape, and the identity size."""
    rng = np.random.RandomState(H + W)
    frames = rng.randint(0, 256, (3, H, W, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    frames[1] = np.stack([(xx * 255 // max(W - 1, 1)), (yy * 255 // max(H - 1, 1)), ((xx + yy) % 256)], -1).astype(np.uint8)
    frames[2, : H // 2] = 255                                              # saturated edge: exercises the clipping
    frames[2, H // 2:] = 0
    ref = restate.clip_resize_center_crop_u8(frames)
    out = ops.resize_center_crop_u8(torch.from_numpy(frames).to(DEV))
    assert out.shape == (3, 224, 224, 3) and out.dtype == torch.uint8
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("H,W", [(256, 456), (300, 200), (128, 228), (224, 224)])
def test_pad_square_and_resize_bit_exact(tower_hf, H, W):
    """The `image_aspect_ratio == 'pad'` branch of load_image (video_utils.py:30-31): expand2square on the device, then the
    resize -- against the oracle (pinned to the reference's own expand2square + PIL by the preprocess_* fixtures)."""
    tw, _ = tower_hf
    frames = np.random.RandomState(H * 7 + W).randint(0, 256, (2, H, W, 3), dtype=np.uint8)
    ref_sq = restate.expand2square_u8(frames)
    out_sq = ops.pad_square_u8(torch.from_numpy(frames).to(DEV))
    assert np.array_equal(out_sq.cpu().numpy(), ref_sq)
    out = tw.preprocess_u8(torch.from_numpy(frames).to(DEV), image_aspect_ratio="pad")
    assert np.array_equal(out.cpu().numpy(), restate.clip_resize_center_crop_u8(ref_sq))


def test_decoded_frames_to_features(tower_hf):
    """Decoded EPIC-KITCHENS-sized frames (256 x 456 uint8) -> preprocess_u8 (resize + crop kernel) -> tower (rescale +
    normalise fused into the patch extraction) against the oracle's processor + ViT."""
    tw, sd = tower_hf
    frames = np.random.RandomState(9).randint(0, 256, (2, 256, 456, 3), dtype=np.uint8)
    px_ref = restate.clip_normalize_u8(restate.clip_resize_center_crop_u8(frames))
    ref = restate.vit_hidden(px_ref, sd, 23)
    u8 = tw.preprocess_u8(torch.from_numpy(frames).to(DEV))
    assert u8.shape == (2, 224, 224, 3)
    hid = tw.forward_hidden(u8)
    assert relmax(hid, ref) <= TOL_BF16
    with pytest.raises(ValueError):
        ops.resize_center_crop_u8(torch.zeros(1, 3, 224, 224, dtype=torch.uint8, device=DEV))


# ------------------------------------------------------------------ boundary details
def test_tower_dtype_follows_module_and_outputs_follow_images(tower_hf):
    tw, sd = tower_hf
    assert tw.dtype == torch.float32 and tw.device.type == "cuda"
    px = synth.pixels((1, 3, 224, 224), seed=6)
    assert tw(px.to(DEV)).dtype == torch.float32 and tw(px.to(DEV).half()).dtype == torch.float16
    t2 = CLIPVisionTower("synthetic", types.SimpleNamespace(mm_vision_select_layer=1), delay_load=True)
    t2.load_model(synth.clip_state_dict(synth.VIT_L14, 0, "hf", n_layers=1))
    t2.to(device=DEV, dtype=torch.float16)                                # handsonvlm/model/builder.py:108
    assert t2.dtype == torch.float16 and t2.dummy_feature.dtype == torch.float16 and t2.dummy_feature.is_cuda
    assert t2.image_processor is not None


@pytest.mark.parametrize("case", ["two_images", "ragged", "train_b3"])
def test_last_visual_token_index_matches_reference(case):
    """handsonvlm.py:288 side effect as a 0-d int64 device tensor written by the plan kernel: reference fixtures for two
    image tokens in one sample (relative offset of the SECOND), and a last sample WITHOUT image token (earlier sample's
    value survives)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"splice_hvlm_{case}.npz"))
    ids = torch.from_numpy(g["ids"])
    B, Nv, D = ids.shape[0], int(g["t"]) + 256, 64
    vis = synth.gen("lv.vis", (B, Nv, D), 1.0, 1)
    table = synth.gen("lv.tab", (synth.VOCAB, D), 1.0, 1)
    host = types.SimpleNamespace(
        get_model=lambda: types.SimpleNamespace(embed_tokens=types.SimpleNamespace(weight=table.to(DEV))),
        config=types.SimpleNamespace())
    lve = torch.full((), -7, dtype=torch.int64, device=DEV)
    arch.splice_tokens(host, L.SPLICE_HANDSONVLM, ids.to(DEV), torch.from_numpy(g["in_mask"]).to(DEV),
                       torch.from_numpy(g["in_labels"]).to(DEV), vis.to(DEV), None,
                       torch.from_numpy(g["future_hands"]).to(DEV), False, last_visual_end=lve)
    assert lve.dim() == 0 and int(lve) == int(g["last_visual_token_index"]) == restate.last_visual_token_index(ids, Nv)
    # no sample has an image token: the previous value is left alone
    none = ids.clone()
    none[none == -200] = 5
    lve.fill_(-7)
    arch.splice_tokens(host, L.SPLICE_HANDSONVLM, none.to(DEV), None, None, vis.to(DEV), None, None, True,
                       last_visual_end=lve)
    assert int(lve) == -7


def test_visual_token_cache_is_not_fooled_by_recycled_storage(tower_hf):
    """the next sample's freshly allocated clip can get the freed clip's address, shape and version; the cache
    must not serve the previous clip's tokens for it."""
    tw, sd = tower_hf
    D, t = 256, 2
    proj, ps = projector(D, torch.float32)
    emb = torch.nn.Embedding(synth.VOCAB, D)
    host = make_host(tw, proj, emb, video_cfg(hvlm_cache_visual_tokens=True), 1)
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=1, seed=5)
    outs, ptrs = [], []
    with torch.no_grad():
        for seed in (1, 2, 3):
            px = synth.pixels((1, t, 3, 224, 224), seed=seed).to(DEV)      # a new tensor every "sample"
            ptrs.append(px.data_ptr())
            r = host.prepare_inputs_labels_for_multimodal(ids.to(DEV), mask.to(DEV), None, None, px, is_evaluate=True)
            r_again = host.prepare_inputs_labels_for_multimodal(ids.to(DEV), mask.to(DEV), None, None, px, is_evaluate=True)
            assert torch.equal(r[3], r_again[3])
            outs.append(r[3].clone())
            del px, r, r_again
    cache = host.__dict__["_hvlm_visual_cache"]
    assert cache.hits == 3
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])


def test_static_splice_contract_violation_is_reported():
    """with hvlm_static_splice the plan status used to be stashed and never looked at.  Now a broken collator
    contract (a sample with two image tokens / none) raises at the next call's poll or at check_deferred_status."""
    D = 128
    table = synth.embed_table(D)
    vis = synth.gen("vis", (2, 356, D), 1.0, 5)
    ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=2, seed=3)
    host = types.SimpleNamespace(
        get_model=lambda: types.SimpleNamespace(embed_tokens=types.SimpleNamespace(weight=table.to(DEV))),
        config=types.SimpleNamespace(hvlm_static_splice=True))
    dv = lambda x: x.to(DEV)
    run = lambda i: arch.splice_tokens(host, L.SPLICE_HANDSONVLM, dv(i), dv(mask), dv(labels), dv(vis), None, dv(fh), False)
    run(ids)
    arch.check_deferred_status(host)                                       # contract held
    bad = ids.clone()
    bad[1, 35] = 77                                                        # sample 1 lost its image token
    run(bad)
    with pytest.raises(RuntimeError, match="static_splice"):
        arch.check_deferred_status(host)
    two = ids.clone()
    two[0, 20] = -200                                                      # sample 0 has two image tokens
    run(two)
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="static_splice"):
        run(ids)                                                           # the NEXT call's poll reports it
    arch.check_deferred_status(host)


def test_splice_many_image_tokens_per_sample():
    """Round 1 capped a sample at 64 image tokens; the plan kernel now uses closed-form row offsets: 150 image tokens in one
    sample, spread over two 256-token chunks of the ids, against the oracle (both variants, ints + rows bit-exact)."""
    D, Nv, T, B = 64, 3, 600, 2
    g = torch.Generator().manual_seed(77)
    ids = torch.randint(1, 32000, (B, T), generator=g, dtype=torch.int64)
    pos = torch.randperm(T - 2, generator=g)[:150].sort().values + 1
    ids[0, pos] = -200
    ids[1, 300] = -200
    labels = ids.clone()
    mask = torch.ones(B, T, dtype=torch.bool)
    vis = synth.gen("many.vis", (151, Nv, D), 1.0, 3)
    table = synth.gen("many.tab", (synth.VOCAB, D), 1.0, 3)
    host = types.SimpleNamespace(
        get_model=lambda: types.SimpleNamespace(embed_tokens=types.SimpleNamespace(weight=table.to(DEV))),
        config=types.SimpleNamespace())
    for variant, name in ((L.SPLICE_LLAVA, "llava"), (L.SPLICE_HANDSONVLM, "handsonvlm")):
        lve = torch.full((), -1, dtype=torch.int64, device=DEV)
        m2, e2, l2 = arch.splice_tokens(host, variant, ids.to(DEV), mask.to(DEV), labels.to(DEV), vis.to(DEV), None, None,
                                        True, last_visual_end=lve)
        rm, re_, rl = restate.splice(ids, mask, labels, vis, table, name, is_evaluate=True)
        assert e2.shape == re_.shape == (B, T + 150 * (Nv - 1), D)
        assert torch.equal(e2.cpu(), re_) and torch.equal(l2.cpu(), rl) and torch.equal(m2.cpu(), rm), name
        assert int(lve) == restate.last_visual_token_index(ids, Nv)


_OPTIN_SCRIPT = r'''
import sys, torch
sys.path.insert(0, sys.argv[1])
import hvlm_b200
from hvlm_b200 import ops
from oracle import synth
dev = "cuda:0"
def relmax(a, b):
    return float((a.float().cpu() - b.float().cpu()).abs().max() / b.float().cpu().abs().max())
for (M, N, K) in [(300, 1024, 1024), (2571, 1024, 1024), (25700, 1024, 1024), (777, 256, 72), (1025, 512, 4096)]:
    a = synth.gen("o.A", (M, K), 1.0, 1).to(torch.bfloat16).to(dev)
    w = synth.gen("o.W", (N, K), K ** -0.5, 1).to(torch.bfloat16).to(dev)
    bias = synth.gen("o.b", (N,), 0.5, 1).to(dev)
    res = synth.gen("o.r", (M, N), 1.0, 2).to(dev)
    ref = a.float() @ w.float().t() + bias + res
    r2 = res.clone()
    ops.gemm(a, w, bias, epilogue="residual", resid=r2, out=r2)
    assert relmax(r2, ref) <= 2e-5, (M, N, K, relmax(r2, ref))
    r3 = res.clone()
    ops.gemm(a, w, bias, epilogue="residual", resid=r3, out=r3)
    assert torch.equal(r2, r3)
y = synth.gen("o.y", (3 * 257, 1024), 1.0, 3).to(torch.bfloat16).to(dev)
wq = synth.gen("o.wq", (3072, 1024), 1024 ** -0.5, 3).to(torch.bfloat16).to(dev)
bq = synth.gen("o.bq", (3072,), 0.1, 3).to(dev)
qkv = ops.vit_qkv(y, wq, bq, 3)                                  # [48, M, 64] column-block-major
ref = (y.float() @ wq.float().t() + bq).reshape(3 * 257, 48, 64).permute(1, 0, 2)
assert relmax(qkv, ref) <= 5e-3
# attention (HVLM_ATTN_PINGPONG=1: one CTA per SM, two programs taking turns on the exp pass), 37 frames = 592 items:
# uneven item lists, program B's padding rounds
F = 37
qb = torch.randn(48, F * 257, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(3)).to(torch.bfloat16)
qb[:16] *= 0.3
out = ops.vit_attention(qb)
t = qb.reshape(3, 16, F, 257, 64).float()
ref = torch.nn.functional.scaled_dot_product_attention(t[0], t[1], t[2], scale=1.0)
assert relmax(out.reshape(F, 257, 16, 64).permute(2, 0, 1, 3), ref) <= 8e-3
assert torch.equal(out, ops.vit_attention(qb))
for F2 in (1, 2, 19):
    q2 = qb[:, : F2 * 257].contiguous()
    t2 = q2.reshape(3, 16, F2, 257, 64).float()
    r2_ = torch.nn.functional.scaled_dot_product_attention(t2[0], t2[1], t2[2], scale=1.0)
    assert relmax(ops.vit_attention(q2).reshape(F2, 257, 16, 64).permute(2, 0, 1, 3), r2_) <= 8e-3, F2
print("optin ok")
'''


def test_opt_in_kernel_variants_stay_correct(tmp_path):
    """The experiment switches that are off by default (residual-load epilogue, second epilogue group for the QKV GEMM,
    ping-pong attention) are read once per process, so they are exercised in a child process: same tolerances as the
    default kernels."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "optin.py"
    script.write_text(_OPTIN_SCRIPT)
    env = dict(os.environ, HVLM_RESID_LOAD_MAXK="4096", HVLM_QKV_EPI_GROUPS="2", HVLM_ATTN_PINGPONG="1")
    r = subprocess.run([sys.executable, str(script), root], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "optin ok" in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]


def test_frame_dedup_randomised():
    """Seeded fuzz: random duplicate patterns, frame sizes (16-byte multiples), element types; frame map / rep / count against
    the oracle, and gather_rows(rep) reproduces every frame through the map."""
    rng = np.random.RandomState(11)
    for case in range(20):
        n = int(rng.choice([2, 3, 17, 64, 100, 257, 1000]))
        u = int(rng.randint(1, n + 1))
        words = int(rng.choice([4, 12, 100, 1029, 37632]))            # 16-byte vectors per frame: x4 int32 words
        base = torch.from_numpy(rng.randint(-2 ** 31, 2 ** 31 - 1, (u, words * 4), dtype=np.int64).astype(np.int32))
        if case % 3 == 0 and u > 1:
            base[1] = base[0]
            base[1, -1] ^= 1                                           # differs from frame 0 in its very last bit only
        pick = torch.from_numpy(rng.randint(0, u, n))
        frames = base[pick].contiguous()
        if case % 2:
            frames = frames.view(torch.uint8)
        fr, rr = restate.frame_dedup(frames)
        fm, rep, nu = ops.frame_dedup(frames.to(DEV))
        U = int(nu.item())
        assert U == rr.numel(), (case, U, rr.numel())
        assert torch.equal(fm.cpu(), fr) and torch.equal(rep[:U].cpu(), rr), case
        back = ops.gather_rows(frames.to(DEV), rep, U)[fm.long()]
        assert torch.equal(back.cpu(), frames), case


def test_resize_center_crop_randomised_geometries():
    """Seeded fuzz over source geometries (portrait / landscape, mild to strong down-scaling, up-scaling, odd sizes whose
    rows are not 4- or 16-byte multiples): bit-exact against the oracle (pinned to PIL)."""
    rng = np.random.RandomState(12)
    for case in range(14):
        H = int(rng.randint(120, 760))
        W = int(rng.randint(120, 1100))
        frames = rng.randint(0, 256, (2, H, W, 3), dtype=np.uint8)
        ref = restate.clip_resize_center_crop_u8(frames)
        out = ops.resize_center_crop_u8(torch.from_numpy(frames).to(DEV))
        assert np.array_equal(out.cpu().numpy(), ref), (case, H, W)


def test_training_shaped_backward_config5_size(tower_hf):
    """BASELINE configs[4] at full per-GPU size: 4 clips x 100 frames, D = 4096, fwd + bwd through pool / projector / splice /
    gather with upstream gradients; projector dW / db (wgrad GEMM K = 4 x 356 = 1424 pooled rows, through the reducer's
    flat fp32 bucket) against the fp32 oracle's autograd formulas."""
    tw, sd = tower_hf
    D, t, B = 4096, 100, 4
    proj, ps = projector(D)
    emb = torch.nn.Embedding(synth.VOCAB, D)
    emb.weight.data.copy_(synth.embed_table(D))
    emb.weight.requires_grad_(False)
    host = make_host(tw, proj, emb.to(torch.bfloat16), video_cfg(hvlm_static_splice=True), B)
    pm = host.model.mm_projector
    red = hd.ProjectorGradReducer(pm)
    try:
        g = torch.Generator(device=DEV)
        g.manual_seed(55)
        px = torch.randn(B, t, 3, 224, 224, device=DEV, generator=g).to(torch.bfloat16)
        ids, mask, labels, fh, fv = synth.prompt_handsonvlm(B=B, seed=55)
        Lout = ids.shape[1] + 355
        de = synth.gen("c5.de", (B, Lout, D), 1.0, 55).to(torch.bfloat16)
        dg = synth.gen("c5.dg", (B, 2, 4, D // 2), 1.0, 55).to(torch.bfloat16)
        r = host.prepare_inputs_labels_for_multimodal(ids.to(DEV), mask.to(DEV), None, labels.to(DEV), px,
                                                      future_hands=fh.to(DEV), future_valid=fv.to(DEV), is_evaluate=False)
        gout, _ = host.gather_hand_traj_states(r[3], r[4], strict=False)
        torch.autograd.backward([r[3], gout], [de.to(DEV), dg.to(DEV)])
        arch.check_deferred_status(host)
        nW = pm.weight.numel()
        dW_gpu, db_gpu = red.bucket[:nW].view(pm.weight.shape).clone(), red.bucket[nW:].clone()
    finally:
        red.close()
    # oracle
    torch.set_num_threads(os.cpu_count() or 1)
    _, _, rl = restate.splice(ids, mask, labels, torch.zeros(B, 356, 8), torch.zeros(synth.VOCAB, 8), "handsonvlm",
                              future_hands=fh)
    _, _, rows = restate.gather_hand_traj(torch.zeros(B, Lout, D), rl)
    d_emb = de.float() + restate.gather_hand_traj_backward(dg.float(), rows, Lout)
    d_vis, _ = restate.splice_backward(d_emb, ids, 356, B, synth.VOCAB)
    pooled = []
    for b in range(B):                                          # one clip at a time: bounded host memory
        feats = restate.tower_forward(px[b].float().cpu(), sd, -2).reshape(1, t, 256, 1024)
        pooled.append(restate.pool_tokens(feats, "temporal_spatial_pool"))
    dW, db = restate.projector_grads(torch.cat(pooled, 0), d_vis)
    assert relmax(dW_gpu, dW) <= TOL_BF16 and relmax(db_gpu, db) <= 1e-4
    assert relmax(pm.weight.grad, dW) <= TOL_BF16


def test_bench_line_has_the_contract_keys(tmp_path):
    """`python bench.py` (forward-only mode, 3 steps) on this GPU: one JSON line with the driver's keys, a roofline object
    whose fraction is achieved / peak, an e2e object with declared copy sizes, a non-zero launch count and a clock record."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--mode", "forward", "--steps", "3", "--warmup", "3",
                        "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
        assert k in line, k
    assert line["unit"] == "frames/s" and line["n_gpus"] == 1 and line["steps"] == 3 and line["warmup"] >= 3
    assert line["value"] > 1000 and abs(line["value"] - 100 / (line["ms_per_step"] / 1e3)) / line["value"] < 1e-3
    # per step: im2col + patch GEMM + pre-LayerNorm + 23 x (QKV, attention, out_proj, fc1, fc2) - the pooled last fc2
    # + 2 pools + 2 GEMMs + 3 splice + gather = 125 (171 with HVLM_LN_FOLD=0: two LayerNorm launches more per layer)
    assert line["gpu_launches"] == 3 * 125
    rf = line["roofline"]
    assert rf["bound"] == "tensor" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3 and 0.3 < rf["frac"] < 1.2
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] == 100 * 3 * 224 * 224 * 4 + 62 * (8 + 1 + 8) + 2 * 4 * 2 * 4 + 2 and e["d2h_bytes_per_step"] > 0
    assert e["value"] > 1000 and "workload" in line["config"] and line["config"]["frames_per_clip"] == 100
