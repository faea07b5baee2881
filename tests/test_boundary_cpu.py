"""CPU-only: the drop-in vision tower keeps the host-side contract of the reference's CLIPVisionTower / build_vision_tower
(llava/model/multimodal_encoder/{clip_encoder,builder}.py) that the reference's loader and trainer rely on BEFORE the first
forward: handsonvlm/model/builder.py:104-109 and handsonvlm/train/train.py:310-318.  Only host code runs here (weight
packing uses the library's host-side layout function); no kernel is launched."""
import os
import types

import numpy as np
import pytest
import torch

import hvlm_b200
from hvlm_b200.builder import build_vision_tower
from hvlm_b200.tower import CLIPVisionTower
from oracle import ref_shim, synth


@pytest.fixture(scope="module")
def one_layer_sd(tmp_path_factory):
    sd = synth.clip_state_dict(synth.VIT_L14, 0, "hf", n_layers=1)
    path = tmp_path_factory.mktemp("clip") / "clip-vit-l14-1layer.pt"
    torch.save(sd, path)
    return sd, str(path)


def _args(path=None, layer=1):
    return types.SimpleNamespace(vision_tower=path, mm_vision_tower=path, mm_vision_select_layer=layer,
                                 mm_vision_select_feature="patch", pretrain_mm_mlp_adapter=None)


def test_tower_contract_like_reference(one_layer_sd):
    sd, path = one_layer_sd
    tw = CLIPVisionTower(path, _args(path), delay_load=True)
    assert not tw.is_loaded and tw.config.hidden_size == 1024           # cfg_only before loading (clip_encoder.py:69-73)
    tw.load_model()                                                     # clip_encoder.py:22-27
    assert tw.is_loaded and tw.hidden_size == 1024 and tw.num_patches == 256
    assert all(not p.requires_grad for p in tw.parameters())
    # image_processor: what builder.py:109 / train.py:318 read; CLIP defaults (shortest edge 224, crop 224, CLIP mean/std)
    ip = tw.image_processor
    assert ip.crop_size["height"] == 224 and ip.size["shortest_edge"] == 224
    assert np.allclose(ip.image_mean, [0.48145466, 0.4578275, 0.40821073])
    from PIL import Image
    img = Image.fromarray(np.random.RandomState(0).randint(0, 256, (256, 456, 3), dtype=np.uint8))
    px = ip.preprocess(img, return_tensors="pt")["pixel_values"][0]     # hoi_forecast/dataset/video_utils.py:47
    assert tuple(px.shape) == (3, 224, 224) and px.dtype == torch.float32
    # dtype follows the module like the reference's `self.vision_tower.dtype` (clip_encoder.py:57-59): fp32 after loading,
    # whatever `.to(dtype=...)` / `.half()` set afterwards (builder.py:108: vision_tower.to(device=..., dtype=float16))
    assert tw.dtype == torch.float32
    tw.to(device="cpu", dtype=torch.float16)
    assert tw.dtype == torch.float16 and tw.dummy_feature.dtype == torch.float16
    assert tuple(tw.dummy_feature.shape) == (1, 1024)
    tw.bfloat16()
    assert tw.dtype == torch.bfloat16
    assert tw.weight_blob.dtype == torch.uint8                          # the packed weights are bytes: never cast


def test_not_delay_load_loads_immediately(one_layer_sd):
    sd, path = one_layer_sd
    tw = CLIPVisionTower(path, _args(path))                             # clip_encoder.py:17-18
    assert tw.is_loaded and hasattr(tw, "image_processor")
    with pytest.raises(ValueError):                                     # select_layer=-2 needs 23 layers, the file has 1
        CLIPVisionTower(path, _args(path, layer=-2))


def test_build_vision_tower_name_rules(one_layer_sd):
    sd, path = one_layer_sd
    tw = build_vision_tower(_args(path), delay_load=True)               # absolute path
    assert isinstance(tw, CLIPVisionTower) and tw.vision_tower_name == path
    tw = build_vision_tower(_args("clip-vit-large-patch14"), delay_load=True)      # bare clip* name -> openai/...
    assert tw.vision_tower_name == os.path.join("openai", "clip-vit-large-patch14")
    tw = build_vision_tower(types.SimpleNamespace(vision_tower="openai/clip-vit-large-patch14", mm_vision_select_layer=-2),
                            delay_load=True)
    assert tw.select_layer == -2 and tw.select_feature == "patch"
    with pytest.raises(ValueError):
        build_vision_tower(_args("some/other-encoder"), delay_load=True)


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present (GPU box)")
def test_reference_loader_and_trainer_walk(one_layer_sd):
    """The reference's own LlavaMetaModel (llava/model/llava_arch.py:25-65), unmodified, with the drop-in builder bound in
    place of its `build_vision_tower`: __init__ with delay_load, then the loader lines (handsonvlm/model/builder.py:104-109)
    and the trainer lines (handsonvlm/train/train.py:310-318)."""
    sd, path = one_layer_sd
    ns = ref_shim.load()
    ref_arch = ns.llava_arch
    orig = ref_arch.build_vision_tower
    ref_arch.build_vision_tower = build_vision_tower
    try:
        class Base(torch.nn.Module):
            def __init__(self, config):
                super().__init__()
                self.config = config

        class Model(ref_arch.LlavaMetaModel, Base):
            pass

        cfg = types.SimpleNamespace(mm_vision_tower=path, mm_hidden_size=1024, hidden_size=64, mm_vision_select_layer=1,
                                    mm_vision_select_feature="patch")
        model = Model(cfg)                                              # llava_arch.py:27-32 (delay_load=True)
        # --- handsonvlm/model/builder.py:104-109
        vision_tower = model.get_vision_tower()
        assert isinstance(vision_tower, CLIPVisionTower) and not vision_tower.is_loaded
        if not vision_tower.is_loaded:
            vision_tower.load_model()
        vision_tower.to(device="cpu", dtype=torch.float16)              # 'cuda' in the reference
        image_processor = vision_tower.image_processor
        assert image_processor is not None and vision_tower.dtype == torch.float16
        # --- handsonvlm/train/train.py:310-318
        model_args = _args(path)
        model.initialize_vision_modules(model_args=model_args, fsdp=None)
        vision_tower = model.get_vision_tower()
        vision_tower.to(dtype=torch.float16, device="cpu")
        data_args = types.SimpleNamespace()
        data_args.image_processor = vision_tower.image_processor
        assert vision_tower.is_loaded and model.config.mm_hidden_size == 1024
        assert isinstance(model.mm_projector, torch.nn.Linear) and model.mm_projector.in_features == 1024
    finally:
        ref_arch.build_vision_tower = orig


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present (GPU box)")
def test_patch_reference_swaps_the_hot_path_only(one_layer_sd):
    """hvlm_b200.patch_reference(): the reference's own classes, imported unmodified, dispatch their hot-path methods to the
    drop-ins afterwards (and nothing else changes); unpatch restores them.  A CPU call through the patched reference mixin
    reaches our ops and fails loudly there (no CPU fallback) -- proof that the reference code path is no longer taken."""
    from hvlm_b200 import arch, integrate
    sd, path = one_layer_sd
    ns = ref_shim.load()
    if ns.handsonvlm is None:
        pytest.skip("reference handsonvlm module does not import here")
    ref_prep = ns.handsonvlm.HandsOnVLMForCausalLM.prepare_inputs_labels_for_multimodal
    ref_v2t = ns.lita_arch.LitaMetaForCausalLM.videos_to_tokens
    ref_fwd = ns.handsonvlm.HandsOnVLMForCausalLM.forward
    done = integrate.patch_reference()
    try:
        assert "llava.model.llava_arch.LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal" in done
        assert "handsonvlm.model.language_model.handsonvlm.HandsOnVLMForCausalLM.prepare_inputs_labels_for_multimodal" in done
        H = ns.handsonvlm.HandsOnVLMForCausalLM
        assert H.prepare_inputs_labels_for_multimodal is arch.HandsOnVLMMetaForCausalLM.prepare_inputs_labels_for_multimodal
        assert ns.lita_arch.LitaMetaForCausalLM.videos_to_tokens is arch.LitaMetaForCausalLM.videos_to_tokens
        assert ns.llava_arch.build_vision_tower is build_vision_tower
        assert ns.clip_encoder.CLIPVisionTower is CLIPVisionTower and ns.v2t.VisualToTokenHelper is arch.VisualToTokenHelper
        assert ns.handsonvlm.VisualToTokenHelper is arch.VisualToTokenHelper and hasattr(H, "gather_hand_traj_states")
        assert H.forward is ref_fwd                                     # everything off the path is untouched
        # a reference-side host that mixes in the (patched) reference mixin, like LitaLlamaForCausalLM does
        tower = CLIPVisionTower(path, _args(path), delay_load=True)
        tower.load_model()
        proj = torch.nn.Linear(1024, 64)

        class Host(ns.lita_arch.LitaMetaForCausalLM):
            config = types.SimpleNamespace(input_type="video", video_arch="temporal_spatial_pool")

            def get_model(self):
                return types.SimpleNamespace(get_vision_tower=lambda: tower, mm_projector=proj, vision_tower=tower)

        with pytest.raises(RuntimeError, match="no CPU fallback"):
            Host().visual_to_tokens(torch.zeros(1, 2, 3, 224, 224))
    finally:
        integrate.unpatch_reference()
    assert ns.handsonvlm.HandsOnVLMForCausalLM.prepare_inputs_labels_for_multimodal is ref_prep
    assert ns.lita_arch.LitaMetaForCausalLM.videos_to_tokens is ref_v2t
    assert not hasattr(ns.handsonvlm.HandsOnVLMForCausalLM, "gather_hand_traj_states")
