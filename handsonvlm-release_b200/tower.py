"""Drop-in for the reference's CLIP vision tower wrapper.

Mirrors ``CLIPVisionTower`` (llava/model/multimodal_encoder/clip_encoder.py:7-78): same constructor arguments,
``forward(images)`` (tensor or list), ``feature_select`` semantics (``select_layer``, ``select_feature`` in
{'patch','cls_patch'}), and the ``dummy_feature / dtype / device / config / hidden_size / num_patches`` properties.
The HF ``CLIPVisionModel`` underneath is replaced by the sm_100a kernels behind ``hvlm_vit_l14_fwd``.
"""
from __future__ import annotations

import types

import torch
import torch.nn as nn

from . import ops
from .weights import pack_vit_weights

_CFG = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16, image_size=224,
            patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5, projection_dim=768)


def _load_clip_state_dict(src: str) -> dict:
    import os
    if os.path.isfile(src):
        if src.endswith(".safetensors"):
            from safetensors.torch import load_file
            return load_file(src)
        return torch.load(src, map_location="cpu")
    from transformers import CLIPVisionModel                      # directory or hub id, like the reference
    return CLIPVisionModel.from_pretrained(src).state_dict()


def _default_image_processor(src=None):
    """The reference reads ``vision_tower.image_processor`` (handsonvlm/model/builder.py:109, handsonvlm/train/train.py:318),
    set by ``CLIPImageProcessor.from_pretrained(name)`` in load_model (clip_encoder.py:23).  Prefer the PIL-backed processor
    (what transformers==4.31.0, the reference's pin, runs); fall back to the CLIP defaults (shortest edge 224, bicubic,
    centre crop 224, OPENAI_CLIP mean / std) when ``src`` has no preprocessor_config.json or there is no network."""
    import transformers
    cls = getattr(transformers.models.clip, "CLIPImageProcessorPil", None) or transformers.CLIPImageProcessor
    if isinstance(src, str):
        try:
            return cls.from_pretrained(src)
        except Exception:
            pass
    return cls()


class CLIPVisionTower(nn.Module):
    def __init__(self, vision_tower="openai/clip-vit-large-patch14", args=None, delay_load=False):
        super().__init__()
        self.is_loaded = False
        self.vision_tower_name = vision_tower
        self.select_layer = getattr(args, "mm_vision_select_layer", -2)
        self.select_feature = getattr(args, "mm_vision_select_feature", "patch")
        self.cfg_only = types.SimpleNamespace(**_CFG)
        # what `.dtype` reports: the reference returns the HF module's parameter dtype (clip_encoder.py:57-59), fp32 after
        # from_pretrained and whatever the caller's `.to(dtype=...)` / `.half()` set afterwards.  The kernels' numeric
        # regime does not depend on it (bf16 operands, fp32 accumulation / residual stream).
        # It is tracked by an empty floating-point buffer that nn.Module's own machinery casts along with the module.
        self.register_buffer("_dtype_probe", torch.zeros(0, dtype=torch.float32), persistent=False)
        self.register_buffer("weight_blob", torch.zeros(0, dtype=torch.uint8), persistent=False)
        if not delay_load:
            self.load_model(vision_tower if isinstance(vision_tower, dict) else None)

    # -- loading -------------------------------------------------------------------------------
    def load_model(self, state_dict=None):
        """``state_dict``: HF-named CLIPVisionModel tensors, or where to find them.  With no argument this does what
        the reference does (clip_encoder.py:22-26): ``CLIPVisionModel.from_pretrained(self.vision_tower_name)`` -- a
        hub id (needs the HF cache or a network) or a local ``save_pretrained`` directory.  A path to a ``.pt`` /
        ``.bin`` / ``.safetensors`` state dict is accepted as well."""
        if state_dict is None:
            state_dict = self.vision_tower_name
        self.image_processor = _default_image_processor(state_dict if isinstance(state_dict, str) else None)
        if isinstance(state_dict, str):
            state_dict = _load_clip_state_dict(state_dict)
        dev = self.weight_blob.device
        blob = pack_vit_weights(state_dict, n_layers=self.n_layers_needed)
        self._blob_layers = blob.hvlm_n_layers
        if self._blob_layers < self.n_layers_needed:
            raise ValueError(f"state dict has {self._blob_layers} encoder layers, select_layer={self.select_layer} "
                             f"needs {self.n_layers_needed}")
        self.weight_blob = blob.to(dev)
        self.requires_grad_(False)
        self.is_loaded = True

    @property
    def n_layers_needed(self) -> int:
        L = _CFG["num_hidden_layers"]
        return self.select_layer if self.select_layer >= 0 else L + 1 + self.select_layer

    # -- forward -------------------------------------------------------------------------------
    def feature_select(self, hidden: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
        if self.select_feature == "patch":
            return ops.feature_select(hidden, out_dtype, keep_cls=False)
        elif self.select_feature == "cls_patch":
            return ops.feature_select(hidden, out_dtype, keep_cls=True)
        raise ValueError(f"Unexpected select feature: {self.select_feature}")

    def forward_hidden(self, images: torch.Tensor) -> torch.Tensor:
        """fp32 residual stream [N,257,1024] == HF hidden_states[select_layer] (no copy of the patch rows)."""
        if not self.is_loaded:
            raise RuntimeError("CLIPVisionTower.load_model() has not been called")
        if images.dtype == torch.uint8:
            # raw decoded frames [N,224,224,3]: rescale + CLIP normalisation are fused into the patch extraction
            return ops.vit_l14_hidden_u8(self.weight_blob, images.to(self.device), self.n_layers_needed)
        if images.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            images = images.float()
        return ops.vit_l14_hidden(self.weight_blob, images.to(self.device), self.n_layers_needed)

    def forward_hidden_open(self, images: torch.Tensor):
        """(hidden after the last attention block, f1 = gelu(fc1(LN2(hidden))) bf16 [N,257,4096]): the tower with the last
        layer's fc2 left to the caller -- see ``last_fc2`` and ``arch.video_tokens``."""
        if not self.is_loaded:
            raise RuntimeError("CLIPVisionTower.load_model() has not been called")
        if images.dtype not in (torch.uint8, torch.float32, torch.bfloat16, torch.float16):
            images = images.float()
        return ops.vit_l14_hidden_open(self.weight_blob, images.to(self.device), self.n_layers_needed)

    def last_fc2(self):
        """Views into the weight blob: (W2 bf16 [1024,4096], b2 f32 [1024]) of the last layer that runs."""
        from .weights import vit_layout
        y = vit_layout(self._blob_layers).layer[self.n_layers_needed - 1]
        w = self.weight_blob[int(y.w_fc2): int(y.w_fc2) + 1024 * 4096 * 2].view(torch.bfloat16).view(1024, 4096)
        b = self.weight_blob[int(y.b_fc2): int(y.b_fc2) + 1024 * 4].view(torch.float32)
        return w, b

    def preprocess_u8(self, frames: torch.Tensor, image_aspect_ratio: str = "square") -> torch.Tensor:
        """Decoded frames uint8 [N,H,W,3] on the device -> uint8 [N,224,224,3]: ``load_image`` of
        hoi_forecast/dataset/video_utils.py:28-53 without the CPU -- the optional ``image_aspect_ratio == 'pad'`` canvas
        (expand2square with int(255 * image_mean), :30-31) and this tower's ``image_processor`` resize + centre crop as
        kernels, bit-identical to PIL.  The result goes straight into ``forward`` / ``forward_hidden`` (rescale + normalise
        are fused there)."""
        frames = frames.to(self.device)
        if image_aspect_ratio == "pad":
            mean = getattr(getattr(self, "image_processor", None), "image_mean", None) or ops.CLIP_MEAN
            frames = ops.pad_square_u8(frames, tuple(int(x * 255) for x in mean))
        return ops.resize_center_crop_u8(frames, _CFG["image_size"])

    @torch.no_grad()
    def forward(self, images):
        if type(images) is list:
            image_features = []
            for image in images:
                hidden = self.forward_hidden(image.unsqueeze(0))
                image_features.append(self.feature_select(hidden, image.dtype))
        else:
            hidden = self.forward_hidden(images)
            image_features = self.feature_select(hidden, images.dtype if images.is_floating_point() else torch.bfloat16)
        return image_features

    # -- properties ----------------------------------------------------------------------------
    @property
    def dummy_feature(self):
        return torch.zeros(1, self.hidden_size, device=self.device, dtype=self.dtype)

    @property
    def dtype(self):
        return self._dtype_probe.dtype

    @property
    def device(self):
        return self.weight_blob.device

    @property
    def config(self):
        return self.cfg_only

    @property
    def hidden_size(self):
        return self.config.hidden_size

    @property
    def num_patches(self):
        return (self.config.image_size // self.config.patch_size) ** 2
