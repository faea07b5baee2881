"""Pins oracle/restate.py against fixtures frozen from the REAL reference (oracle/make_golden.py).

CPU only.  The fixtures are outputs of the unmodified reference functions run in the dev
container; inputs/weights are regenerated here from oracle.synth seeds.
"""
import types

import numpy as np
import pytest
import torch

from oracle import restate, synth
from oracle.make_golden import SMALL, SMALL_D, small_parts


def T(a):
    return torch.from_numpy(np.asarray(a))


def relmax(a, b):
    return float((a - b).abs().max() / b.abs().max())


# ------------------------------------------------------------------ pooling (a3/a4)
@pytest.mark.parametrize("t,d", [(100, 32), (10, 16), (2, 8), (5, 8)])
@pytest.mark.parametrize("mode", ["temporal_spatial_pool", "spatial_pool", "none"])
def test_pool_matches_reference(golden, t, d, mode):
    g = golden(f"pool_{mode}_t{t}")
    tok = synth.gen(f"pooltok{t}", (2, t, 256, d), 1.0, seed=5)
    out = restate.pool_tokens(tok, mode)
    assert list(out.shape[:2]) == list(g["mask_shape"])
    if mode == "none":
        out = out[:, ::97]
    assert relmax(out, T(g["out"])) <= 1e-6
    assert bool(g["mask_all_true"])


def test_selected_frames_known_answers():
    assert restate.selected_frames(100).tolist() == [0, 33, 66, 99]
    assert restate.selected_frames(10).tolist() == [0, 3, 6, 9]
    assert restate.selected_frames(2).tolist() == [0, 0, 1, 1]


def test_pool_backward_matches_reference_autograd(golden):
    g = golden("pool_bwd_t10")
    dout = synth.gen("pooldout_bwd", (1, 266, 8), 1.0, seed=6)
    d = restate.pool_tokens_backward(dout, 10)
    assert relmax(d, T(g["dtok"])) <= 1e-6


# ------------------------------------------------------------------ trajectory head, generation side (8f item 4)
def test_traj_inference_matches_reference(golden):
    """oracle vs the reference's CVAETrajDecoder.inference (fixture froze its output and the noise it drew)."""
    g = golden("traj_infer")
    sd = synth.traj_cvae_state(32, seed=3)
    emb = synth.gen("traj_emb", (3, 2, 4, 32), 1.0, seed=51)
    out = restate.traj_decoder_inference(emb, T(g["z"]), sd)
    assert out.shape == (3, 2, 4, 2)
    assert float((out - T(g["out"])).abs().max()) <= 1e-6


def test_traj_step_matches_reference(golden):
    g = golden("traj_step")
    sd = synth.traj_cvae_state(32, seed=3)
    hidden_last = synth.gen("traj_hidden_last", (1, 64), 1.0, seed=52)
    out = restate.traj_decode_step(hidden_last, T(g["z"]), sd)
    assert out.shape == (1, 2, 2)
    assert float((out - T(g["out"])).abs().max()) <= 1e-6


def test_traj_mlp_inference_matches_reference(golden):
    g = golden("traj_mlp_infer")
    sd = synth.traj_mlp_state(32, seed=4)
    emb = synth.gen("trajmlp_emb", (3, 2, 4, 32), 1.0, seed=53)
    out = restate.traj_mlp_inference(emb, sd)
    assert out.shape == (3, 2, 4, 2) and float((out - T(g["out"])).abs().max()) <= 1e-6


# ------------------------------------------------------------------ gather (a7)
@pytest.mark.parametrize("name,shape,seedname,seed", [
    ("gather_toy", None, None, None),
    ("gather_rand", (4, 40, 64), "gather_hidden", 41),
    ("gather_pos0", (1, 12, 16), "gather_hidden0", 42)])
def test_gather_matches_reference(golden, name, shape, seedname, seed):
    g = golden(name)
    labels = T(g["labels"])
    if shape is None:
        hidden = torch.arange(2 * 10 * 8, dtype=torch.float32).reshape(2, 10, 8)
    else:
        hidden = synth.gen(seedname, shape, 1.0, seed=seed)
    out, valid, rows = restate.gather_hand_traj(hidden, labels)
    assert torch.equal(out, T(g["out"]))
    # reference zeroes future_valid[i] in place for samples without hand tokens
    assert torch.equal(valid, T(g["future_valid"]).all(dim=1))


def test_gather_toy_known_answer():
    hidden = torch.arange(2 * 10 * 8, dtype=torch.float32).reshape(2, 10, 8)
    labels = torch.full((2, 10), -100, dtype=torch.int64)
    labels[0, 5:9] = 32100
    out, valid, rows = restate.gather_hand_traj(hidden, labels)
    assert rows[0].tolist() == [4, 5, 6, 7] and rows[1].tolist() == [-1] * 4
    assert out[0, 0].tolist() == [[32, 34, 36, 38], [40, 42, 44, 46], [48, 50, 52, 54], [56, 58, 60, 62]]
    assert out[0, 1].tolist() == [[33, 35, 37, 39], [41, 43, 45, 47], [49, 51, 53, 55], [57, 59, 61, 63]]
    assert valid.tolist() == [True, False] and float(out[1].abs().sum()) == 0.0


def test_gather_wrong_count_raises():
    labels = torch.full((1, 10), -100, dtype=torch.int64)
    labels[0, 3:6] = 32100
    with pytest.raises(RuntimeError):
        restate.gather_hand_traj(torch.zeros(1, 10, 8), labels)


# ------------------------------------------------------------------ small tower helpers
@pytest.fixture(scope="module")
def small():
    sd, proj, emb = small_parts()
    return sd, proj.weight.data, proj.bias.data, emb.weight.data


def small_visual(small, px, mode="temporal_spatial_pool"):
    sd, pw, pb, _ = small
    return restate.pipeline(px, sd, pw, pb, mode, select_layer=-2, cfg=SMALL)[0]


@pytest.mark.parametrize("arch", ["all", "temporal", "spatial", "temporal_spatial", "temporal_spatial_pool",
                                  "spatial_pool"])
def test_lita_videos_to_tokens(golden, small, arch):
    g = golden(f"lita_{arch}")
    px = synth.pixels((1, 6, 3, 224, 224), seed=7)
    out = small_visual(small, px, arch)
    assert list(out.shape) == list(g["shape"])
    if arch == "all":
        out = out[:, ::7]
    assert relmax(out, T(g["out"])) <= 2e-5


def test_v2t_pipeline(golden, small):
    g = golden("v2t_pipeline")
    px = synth.pixels((1, 6, 3, 224, 224), seed=7)
    sd, pw, pb, _ = small
    out, mask = restate.pipeline(px, sd, pw, pb, cfg=SMALL)
    assert relmax(out, T(g["out"])) <= 2e-5
    assert torch.equal(mask, T(g["mask"]))


# ------------------------------------------------------------------ splice (a5/a6)
def _check_splice(g, mask2, e2, l2, tol=2e-5):
    ge = T(g["embeds"])
    assert e2.shape == ge.shape
    assert relmax(e2, ge) <= tol
    if "labels" in g.files:
        assert torch.equal(l2, T(g["labels"]))
    else:
        assert l2 is None
    if "mask" in g.files:
        gm = T(g["mask"])
        assert str(mask2.dtype) == str(g["mask_dtype"]) if "mask_dtype" in g.files else True
        assert torch.equal(mask2, gm)
    else:
        assert mask2 is None


@pytest.mark.parametrize("name,is_eval", [
    ("splice_hvlm_train_b3", False), ("splice_hvlm_train_padded", False), ("splice_hvlm_2hand", False),
    ("splice_hvlm_0hand", False), ("splice_hvlm_ragged", False), ("splice_hvlm_eval_hands", True),
    ("splice_hvlm_eval_nohands", True), ("splice_hvlm_empty_tail", False), ("splice_hvlm_two_images", False)])
def test_splice_handsonvlm(golden, small, name, is_eval):
    g = golden(name)
    ids = T(g["ids"])
    B = ids.shape[0]
    px = synth.pixels((B, int(g["t"]), 3, 224, 224), seed=int(g["px_seed"]))
    vis = small_visual(small, px)
    mask = T(g["in_mask"]) if "in_mask" in g.files else None
    labels = T(g["in_labels"]) if "in_labels" in g.files else None
    fh = T(g["future_hands"]) if "future_hands" in g.files else None
    m2, e2, l2 = restate.splice(ids, mask, labels, vis, small[3], "handsonvlm", future_hands=fh,
                                is_evaluate=is_eval)
    _check_splice(g, m2, e2, l2)
    # the reference's write-only side effect (handsonvlm.py:288), as the reference left it after this call
    assert restate.last_visual_token_index(ids, vis.shape[1]) == int(g["last_visual_token_index"])
    # text rows are exact copies of embedding rows, visual rows exact copies of pipeline() output
    plan, _ = restate.splice_plan(ids[0], vis.shape[1])
    for r, (kind, idx) in enumerate(plan):
        if kind == 1:
            assert torch.equal(e2[0, r], vis.reshape(-1, vis.shape[-1])[idx])   # sample 0 owns visual slots 0..k-1


@pytest.mark.parametrize("name,pxshape,pxseed,cfg", [
    ("splice_llava_cfg1", (1, 3, 224, 224), 12, "image"),
    ("splice_llava_ragged", (2, 3, 224, 224), 13, "image"),
    ("splice_llava_two_images", (2, 3, 224, 224), 14, "image"),
    ("splice_llava_video", (1, 4, 3, 224, 224), 15, "video")])
def test_splice_llava(golden, small, name, pxshape, pxseed, cfg):
    g = golden(name)
    sd, pw, pb, ew = small
    px = synth.pixels(pxshape, seed=pxseed)
    if cfg == "image":
        feats = restate.tower_forward(px, sd, -2, SMALL)
        vis = restate.project(feats, pw, pb)
    else:
        vis = small_visual(small, px)
    ids = T(g["ids"])
    m2, e2, l2 = restate.splice(ids, T(g["in_mask"]), T(g["in_labels"]), vis, ew, "llava")
    assert bool(g["returned_ids_is_none"])
    _check_splice(g, m2, e2, l2)


def test_splice_llava_im_start_end(golden, small):
    """The tune_mm_mlp_adapter + mm_use_im_start_end branch (llava_arch.py:146-161): outputs and which embedding rows
    receive gradient, against the reference run with both flags set."""
    g = golden("splice_llava_im_start_end")
    sd, pw, pb, ew = small
    feats = restate.tower_forward(synth.pixels((2, 3, 224, 224), seed=16), sd, -2, SMALL)
    vis = restate.project(feats, pw, pb)
    ids = T(g["ids"])
    m2, e2, l2 = restate.splice(ids, T(g["in_mask"]), T(g["in_labels"]), vis, ew, "llava", im_start_end=True)
    assert torch.equal(l2, T(g["labels"])) and torch.equal(m2, T(g["mask"]))
    assert relmax(e2, T(g["embeds"])) <= 2e-5
    # the label rule differs from the plain branch exactly at the <im_end> slot
    _, _, l_plain = restate.splice(ids, T(g["in_mask"]), T(g["in_labels"]), vis, ew, "llava")
    assert int((l_plain != l2).sum()) == ids.shape[0]
    dout = synth.gen("ise.dout", tuple(e2.shape), 1.0, seed=16)
    _, d_table = restate.splice_backward(dout, ids, vis.shape[1], vis.shape[0], ew.shape[0], im_start_end=True)
    rows = torch.nonzero(d_table.abs().sum(1) > 0).flatten()
    assert torch.equal(rows, T(g["grad_rows"]))
    assert relmax(d_table[rows], T(g["grad_vals"])) <= 1e-6


def test_splice_hvlm_im_start_end(golden, small):
    """HandsOnVLM's own splice with both flags set (handsonvlm.py:263-286,343-344): position-spliced mask follows the
    <im_end> rule too, and the tail gets no hand embeddings."""
    g = golden("splice_hvlm_im_start_end")
    _, _, _, ew = small
    vis = small_visual(small, synth.pixels((2, int(g["t"]), 3, 224, 224), seed=17))
    m2, e2, l2 = restate.splice(T(g["ids"]), T(g["in_mask"]), T(g["in_labels"]), vis, ew, "handsonvlm",
                                im_start_end=True)
    assert torch.equal(l2, T(g["labels"])) and torch.equal(m2, T(g["mask"])) and m2.dtype == torch.bool
    assert relmax(e2, T(g["embeds"])) <= 2e-5
    assert not bool(g["has_last_visual_token_index"])


def _list_visual(small, sizes, seeds=(18, 19, 20)):
    sd, pw, pb = small[0], small[1], small[2]
    blocks = []
    for n, sd_ in zip(sizes, seeds):
        feats = restate.tower_forward(synth.pixels((int(n), 3, 224, 224), seed=sd_), sd, -2, SMALL)
        blocks.append(restate.project(feats, pw, pb).reshape(-1, pw.shape[0]))       # [n*256, D]
    return blocks


def test_splice_llava_list_of_image_groups(golden, small):
    """images_to_tokens list path (llava_arch.py:95-106): groups of 1 / 3 / 2 images -> token blocks of 256 / 768 / 512
    rows, one image token per sample expands to its whole block; the third sample has no image token."""
    g = golden("splice_llava_list_ragged")
    blocks = _list_visual(small, g["group_sizes"])
    m2, e2, l2 = restate.splice(T(g["ids"]), T(g["in_mask"]), T(g["in_labels"]), blocks, small[3], "llava")
    assert e2.shape == tuple(g["embeds"].shape)
    assert torch.equal(l2, T(g["labels"])) and torch.equal(m2, T(g["mask"]))
    assert relmax(e2, T(g["embeds"])) <= 2e-5


def test_splice_cfg1_shapes(golden):
    g = golden("splice_llava_cfg1")
    assert g["embeds"].shape == (1, 311, SMALL_D)
    assert (g["labels"][0, 35:291] == -100).all() and g["labels"][0, 34] != -100


# ------------------------------------------------------------------ ViT-L/14 (a1; third-party HF CLIP)
@pytest.mark.parametrize("profile,tol", [("hf", 2e-5), ("strong", 5e-5)])
def test_vit_l14_restatement_matches_hf(golden, profile, tol):
    g = golden(f"vit_l14_{profile}")
    sd = synth.clip_state_dict(synth.VIT_L14, seed=0, profile=profile, n_layers=23)
    px = synth.pixels((2, 3, 224, 224), seed=3)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    feats = restate.tower_forward(px, sd, -2)
    assert feats.shape == (2, 256, 1024)
    assert relmax(feats[:, ::8, ::4], T(g["sub"])) <= tol
    assert relmax(feats.norm(dim=-1), T(g["norms"])) <= tol


def test_pool_first_equals_project_first():
    """mean o linear commute (SURVEY 8a notes): the CUDA path pools 1024-d features first."""
    x = synth.gen("commute", (1, 10, 256, 64), 1.0, 3)
    w = synth.gen("commute.w", (48, 64), 0.1, 3)
    b = synth.gen("commute.b", (48,), 0.1, 3)
    a = restate.pool_tokens(restate.project(x, w, b), "temporal_spatial_pool")
    c = restate.project(restate.pool_tokens(x, "temporal_spatial_pool"), w, b)
    assert relmax(c, a) <= 1e-5


# ------------------------------------------------------------------ f3: CLIPImageProcessor resize + centre crop
@pytest.mark.parametrize("H,W", [(256, 456), (224, 224), (480, 640), (300, 200), (225, 230), (720, 1280)])
def test_resize_center_crop_restatement_is_pil_exact(H, W):
    """The integer restatement of Pillow's 8-bit bicubic resampler + transformers' centre crop against PIL itself
    (what transformers==4.31.0's CLIPImageProcessor runs, hoi_forecast/dataset/video_utils.py:47-52): bit-exact."""
    from PIL import Image
    rng = np.random.RandomState(H * 1000 + W)
    frames = rng.randint(0, 256, (2, H, W, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    frames[1] = np.stack([(xx * 255 // max(W - 1, 1)), (yy * 255 // max(H - 1, 1)), ((xx + yy) % 256)], -1).astype(np.uint8)
    out = restate.clip_resize_center_crop_u8(frames)
    nh, nw = restate.clip_resize_output_size(H, W)
    top, left = (nh - 224) // 2, (nw - 224) // 2
    for i in range(2):
        ref = np.asarray(Image.fromarray(frames[i]).resize((nw, nh), resample=Image.BICUBIC))[top:top + 224, left:left + 224]
        assert np.array_equal(out[i], ref)


def test_resize_restatement_matches_hf_pil_processor():
    """End of the preprocessing chain against HuggingFace's PIL-backed CLIPImageProcessor (the 4.31-era code path):
    resize + crop + rescale + normalise -> float pixels."""
    import transformers
    from PIL import Image
    cls = getattr(transformers.models.clip, "CLIPImageProcessorPil", None)
    if cls is None:
        pytest.skip("this transformers has no PIL-backed CLIPImageProcessor")
    frames = np.random.RandomState(5).randint(0, 256, (1, 256, 456, 3), dtype=np.uint8)     # EPIC-KITCHENS frame size
    ref = cls().preprocess(Image.fromarray(frames[0]), return_tensors="pt")["pixel_values"][0]
    mine = restate.clip_normalize_u8(restate.clip_resize_center_crop_u8(frames))[0]
    assert float((ref - mine).abs().max()) <= 1e-6


def test_resize_host_tables_match_restatement():
    """hvlm_resize_table_host (pure host code of the C ABI, no GPU needed) == the oracle's coefficient tables."""
    import ctypes as C
    import hvlm_b200
    lib = hvlm_b200._lib.lib()
    for i, o in [(456, 399), (256, 224), (224, 224), (1280, 398), (200, 224)]:
        k = lib.hvlm_resize_table_host(i, o, 0, o, None, None)
        b = (C.c_int32 * (2 * o))()
        c = (C.c_int32 * (o * k))()
        assert lib.hvlm_resize_table_host(i, o, 0, o, b, c) == k
        rb, rc = restate.resample_table(i, o)
        assert np.array_equal(np.array(b).reshape(o, 2), rb) and np.array_equal(np.array(c).reshape(o, k), rc)
    # a crop window is the same table, sliced
    k = lib.hvlm_resize_table_host(456, 399, 87, 224, None, None)
    b = (C.c_int32 * (2 * 224))()
    c = (C.c_int32 * (224 * k))()
    assert lib.hvlm_resize_table_host(456, 399, 87, 224, b, c) == k
    rb, rc = restate.resample_table(456, 399)
    assert np.array_equal(np.array(b).reshape(224, 2), rb[87:311]) and np.array_equal(np.array(c).reshape(224, k), rc[87:311])
    assert lib.hvlm_resize_table_host(456, 399, 300, 224, b, c) < 0


def test_frame_dedup_restatement():
    x = synth.pixels((3, 3, 8, 8), seed=1)
    clip = x[[0, 1, 0, 2, 1, 1]]
    fmap, rep = restate.frame_dedup(clip)
    assert fmap.tolist() == [0, 1, 0, 2, 1, 1] and rep.tolist() == [0, 1, 3]
    neg0 = torch.zeros(2, 4)
    neg0[1] = -0.0                       # byte-wise comparison: +0 and -0 are different frames
    assert restate.frame_dedup(neg0)[1].tolist() == [0, 1]


@pytest.mark.parametrize("name", ["landscape", "portrait"])
def test_preprocess_matches_reference_fixture(golden, name):
    """load_image of hoi_forecast/dataset/video_utils.py:28-53, both `image_aspect_ratio` branches, frozen from the
    reference's own expand2square + the PIL-backed CLIPImageProcessor: the oracle's integer restatement reproduces the
    float pixels (rescale + normalise rounding only)."""
    g = golden(f"preprocess_{name}")
    frame = g["frame"][None]
    sq = restate.clip_normalize_u8(restate.clip_resize_center_crop_u8(frame))[0]
    assert float((sq[:, ::2, ::2] - T(g["square"])).abs().max()) <= 1e-6
    padded = restate.expand2square_u8(frame)
    assert tuple(padded.shape[1:3][::-1]) == tuple(int(v) for v in g["padded_size"])
    pd = restate.clip_normalize_u8(restate.clip_resize_center_crop_u8(padded))[0]
    assert float((pd[:, ::2, ::2] - T(g["pad"])).abs().max()) <= 1e-6


def test_resample_restatement_vs_pil_random_sizes():
    """Seeded sweep over (in, out) size pairs -- strong down-scaling, up-scaling, identity, tiny images -- of the oracle's
    Pillow restatement against PIL.Image.resize(BICUBIC) itself, both passes, bit-exact."""
    from PIL import Image
    rng = np.random.RandomState(2024)
    for case in range(40):
        H, W = int(rng.randint(3, 260)), int(rng.randint(3, 260))
        nh, nw = int(rng.randint(2, 260)), int(rng.randint(2, 260))
        img = rng.randint(0, 256, (H, W, 3), dtype=np.uint8)
        xb, xc = restate.resample_table(W, nw)
        yb, yc = restate.resample_table(H, nh)
        tmp = restate._resample_axis_last(np.ascontiguousarray(img.transpose(0, 2, 1)), xb, xc)       # [H,3,nw]
        out = restate._resample_axis_last(np.ascontiguousarray(tmp.transpose(1, 2, 0)), yb, yc)      # [3,nw,nh]
        out = out.transpose(2, 1, 0)
        ref = np.asarray(Image.fromarray(img).resize((nw, nh), resample=Image.BICUBIC))
        assert np.array_equal(out, ref), (case, H, W, nh, nw)


def test_resize_plan_host_matches_restatement_for_random_geometries():
    """hvlm_resize_plan_host / hvlm_resize_tables_host (host side of the C ABI): geometry as transformers derives it,
    tables == the oracle's sliced to the crop window, spans (x_lo, x_cols, rows_cap) consistent with those tables."""
    import ctypes as C
    import hvlm_b200
    from hvlm_b200 import _lib as L
    lib = L.lib()
    rng = np.random.RandomState(7)
    for case in range(25):
        H, W = int(rng.randint(224, 1300)), int(rng.randint(224, 2000))
        plan = L.ResizePlan()
        assert lib.hvlm_resize_plan_host(H, W, 224, 224, C.byref(plan)) == 0
        nh, nw = restate.clip_resize_output_size(H, W)
        assert (plan.new_h, plan.new_w) == (nh, nw) and (plan.top, plan.left) == ((nh - 224) // 2, (nw - 224) // 2)
        tab = (C.c_int32 * plan.table_ints)()
        assert lib.hvlm_resize_tables_host(C.byref(plan), tab) == 0
        t = np.array(tab)
        xb_r, xc_r = restate.resample_table(W, nw)
        yb_r, yc_r = restate.resample_table(H, nh)
        xb, rest = t[:448].reshape(224, 2), t[448:]
        xc, rest = rest[:224 * plan.xk].reshape(224, plan.xk), rest[224 * plan.xk:]
        yb, yc = rest[:448].reshape(224, 2), rest[448:].reshape(224, plan.yk)
        assert np.array_equal(xb, xb_r[plan.left:plan.left + 224]) and np.array_equal(xc, xc_r[plan.left:plan.left + 224])
        assert np.array_equal(yb, yb_r[plan.top:plan.top + 224]) and np.array_equal(yc, yc_r[plan.top:plan.top + 224])
        assert plan.x_lo == xb[0, 0] and plan.x_lo + plan.x_cols == xb[-1, 0] + xb[-1, 1] <= W
        need = max(yb[min(y0 + 8, 224) - 1].sum() - yb[y0, 0] for y0 in range(0, 224, 8))
        assert plan.rows_cap == need
    assert lib.hvlm_resize_plan_host(100, 100, 224, 300, C.byref(L.ResizePlan())) < 0      # crop larger than the resize
